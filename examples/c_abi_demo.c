/* libdustyb200 from plain C: no Python, no torch. Exercises the boundary exactly as a host program of the
 * reference's own extensions would (raw device pointers, caller-owned scratch, int status):
 *   scans -> dusty_scan_preprocess -> dusty_fps (+ fused gather) -> dusty_chamfer_matrix -> dusty_cov_mmd_1nna_finalize
 * and checks a few Chamfer entries against a brute-force loop on the host.
 *
 *   gcc -std=c99 -O2 -I include -I /usr/local/cuda/include examples/c_abi_demo.c \
 *       -L dusty-gan_b200/lib -ldustyb200 -L /usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,$PWD/dusty-gan_b200/lib -o c_abi_demo
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "dusty_b200.h"

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) { fprintf(stderr, "CUDA: %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 2; } \
  } while (0)
#define DK(call)                                                                                   \
  do {                                                                                             \
    int rc_ = (call);                                                                              \
    if (rc_ != 0) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, dusty_last_error_string()); return 3; } \
  } while (0)

static float urand(unsigned* s) { *s = *s * 1664525u + 1013904223u; return (float)(*s >> 8) / 16777216.0f; }

/* mean_p min_q + mean_q min_p of squared distances, in double: the host check */
static double chamfer_host(const float* a, const float* b, int p) {
  double s1 = 0, s2 = 0;
  for (int i = 0; i < p; ++i) {
    double best = 1e30, best2 = 1e30;
    for (int j = 0; j < p; ++j) {
      double dx = a[3 * i] - b[3 * j], dy = a[3 * i + 1] - b[3 * j + 1], dz = a[3 * i + 2] - b[3 * j + 2];
      double d = dx * dx + dy * dy + dz * dz;
      if (d < best) best = d;
      dx = b[3 * i] - a[3 * j]; dy = b[3 * i + 1] - a[3 * j + 1]; dz = b[3 * i + 2] - a[3 * j + 2];
      d = dx * dx + dy * dy + dz * dz;
      if (d < best2) best2 = d;
    }
    s1 += best; s2 += best2;
  }
  return s1 / p + s2 / p;
}

int main(void) {
  enum { NS = 12, HS = 16, WS = 512, H = 16, W = 128, P = 256 };   /* 12 scans -> 6 "ref" + 6 "gen" clouds */
  const int npix = H * W;
  if (dusty_abi_version() != DUSTY_B200_ABI_VERSION) { fprintf(stderr, "ABI mismatch\n"); return 1; }

  /* synthetic organised scans (x, y, z, reflectance), some pixels empty */
  size_t scan_floats = (size_t)NS * HS * WS * 4;
  float* h_scans = (float*)malloc(scan_floats * sizeof(float));
  unsigned seed = 12345u;
  for (int s = 0; s < NS; ++s)
    for (int r = 0; r < HS; ++r)
      for (int c = 0; c < WS; ++c) {
        float* q = h_scans + (((size_t)s * HS + r) * WS + c) * 4;
        const float el = (2.0f - 26.8f * r / (HS - 1)) * 0.017453292f, az = 3.14159265f - 6.2831853f * c / WS;
        float range = 5.0f + 60.0f * urand(&seed);
        if (urand(&seed) < 0.25f) range = 0.0f;
        q[0] = range * cosf(el) * cosf(az); q[1] = range * cosf(el) * sinf(az); q[2] = range * sinf(el); q[3] = urand(&seed);
      }

  float *d_scans, *d_mask, *d_inv, *d_points, *d_sampled, *d_M, *d_out;
  int32_t* d_idx;
  CK(cudaMalloc((void**)&d_scans, scan_floats * sizeof(float)));
  CK(cudaMalloc((void**)&d_mask, (size_t)NS * npix * sizeof(float)));
  CK(cudaMalloc((void**)&d_inv, (size_t)NS * npix * sizeof(float)));
  CK(cudaMalloc((void**)&d_points, (size_t)NS * npix * 3 * sizeof(float)));
  CK(cudaMalloc((void**)&d_sampled, (size_t)NS * P * 3 * sizeof(float)));
  CK(cudaMalloc((void**)&d_idx, (size_t)NS * P * sizeof(int32_t)));
  CK(cudaMalloc((void**)&d_M, (size_t)NS * NS * sizeof(float)));
  CK(cudaMalloc((void**)&d_out, 7 * sizeof(float)));
  CK(cudaMemcpy(d_scans, h_scans, scan_floats * sizeof(float), cudaMemcpyHostToDevice));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));

  /* 1. raw scans -> inverse-depth image, mask, unit-space points */
  const double min_depth = 0.9, max_depth = 120.0;
  dusty_scan_params sp;
  memset(&sp, 0, sizeof sp);
  sp.b = NS; sp.hs = HS; sp.ws = WS; sp.channels = 4; sp.h = H; sp.w = W;
  sp.scale_h = (float)HS / (float)H; sp.scale_w = (float)WS / (float)W;
  sp.min_depth = (float)min_depth; sp.max_depth = (float)max_depth; sp.range = (float)(max_depth - min_depth);
  sp.disp_lo = (float)(1.0 / max_depth); sp.inv_disp_range = (float)(1.0 / (1.0 / min_depth - 1.0 / max_depth));
  sp.drop_const = -1.0f;
  DK(dusty_scan_preprocess(&sp, d_scans, NULL, d_mask, d_inv, d_points, NULL, st));

  /* 2. farthest-point sampling with the fused gather */
  size_t fps_bytes = dusty_fps_workspace_bytes(NS, npix, P);
  void* d_fps_ws;
  CK(cudaMalloc(&d_fps_ws, fps_bytes));
  DK(dusty_fps(d_points, NS, npix, P, d_idx, d_sampled, d_fps_ws, fps_bytes, st));

  /* 3. the stacked symmetric Chamfer matrix in one launch, 4. MMD / COV / 1-NNA on the device */
  size_t mat_bytes = dusty_chamfer_matrix_workspace_bytes(NS, P, 0, P);
  void* d_mat_ws;
  CK(cudaMalloc(&d_mat_ws, mat_bytes));
  DK(dusty_chamfer_matrix(d_sampled, NS, P, d_sampled, NS, P, 0, NS, 1, DUSTY_MATRIX_SYMMETRIC | DUSTY_MATRIX_MIRROR,
                          d_M, NS, d_mat_ws, mat_bytes, st));
  const int nr = NS / 2, ng = NS / 2;
  float *d_rr, *d_rg, *d_gg;
  CK(cudaMalloc((void**)&d_rr, (size_t)nr * nr * sizeof(float)));
  CK(cudaMalloc((void**)&d_rg, (size_t)nr * ng * sizeof(float)));
  CK(cudaMalloc((void**)&d_gg, (size_t)ng * ng * sizeof(float)));
  CK(cudaMemcpy2DAsync(d_rr, nr * sizeof(float), d_M, NS * sizeof(float), nr * sizeof(float), nr, cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpy2DAsync(d_rg, ng * sizeof(float), d_M + nr, NS * sizeof(float), ng * sizeof(float), nr, cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpy2DAsync(d_gg, ng * sizeof(float), d_M + (size_t)nr * NS + nr, NS * sizeof(float), ng * sizeof(float), ng, cudaMemcpyDeviceToDevice, st));
  size_t fin_bytes = dusty_cov_mmd_1nna_workspace_bytes(nr, ng);
  void* d_fin_ws;
  CK(cudaMalloc(&d_fin_ws, fin_bytes));
  DK(dusty_cov_mmd_1nna_finalize(d_rr, d_rg, d_gg, nr, ng, d_out, d_fin_ws, fin_bytes, st));
  CK(cudaStreamSynchronize(st));

  float h_M[NS * NS], h_out[7];
  float* h_sampled = (float*)malloc((size_t)NS * P * 3 * sizeof(float));
  CK(cudaMemcpy(h_M, d_M, sizeof h_M, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(h_out, d_out, sizeof h_out, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(h_sampled, d_sampled, (size_t)NS * P * 3 * sizeof(float), cudaMemcpyDeviceToHost));

  double worst = 0;
  for (int i = 0; i < NS; ++i)
    for (int j = 0; j < NS; ++j) {
      const double want = chamfer_host(h_sampled + (size_t)i * P * 3, h_sampled + (size_t)j * P * 3, P);
      const double err = fabs(h_M[i * NS + j] - want) / (want > 1e-30 ? want : 1.0);
      if (err > worst) worst = err;
      if (h_M[i * NS + j] != h_M[j * NS + i]) { fprintf(stderr, "matrix not symmetric at %d,%d\n", i, j); return 4; }
    }
  if (worst > 1e-5) { fprintf(stderr, "Chamfer entry off by %.3g relative\n", worst); return 5; }
  if (h_out[3] + h_out[4] + h_out[5] + h_out[6] != (float)NS) { fprintf(stderr, "1-NN confusion counts do not add up\n"); return 6; }
  printf("c_abi_demo ok: %d kernels launched, worst relative Chamfer error %.2e, mmd %.6f cov %.0f/%d 1-NN tp/fp/fn/tn %.0f/%.0f/%.0f/%.0f\n",
         (int)dusty_launch_count(), worst, h_out[0], h_out[2], nr, h_out[3], h_out[4], h_out[5], h_out[6]);
  return 0;
}
