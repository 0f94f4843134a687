// Farthest-point sampling + gather for sm_100a.
//
// Replaces furthest_point_sampling_kernel<512> and gather_points_kernel (reference
// utils/sampling/fps/furthest_point_sampling.cu:99-207, :38-50) and their host wrappers
// (furthest_point_sampling.cpp:27-50, :79-100), bit for bit.
//
// What the reference computes (and this kernel reproduces exactly):
//   idx[0] = 0; temp[k] = 1e10; then m-1 times: centre = point idx[j-1]; for every point k with
//   (double)fma(z,z,fma(x,x,y*y)) > 1e-3:  temp[k] = fminf(temp[k], fma(dz,dz,fma(dx,dx,dy*dy)));
//   idx[j] = arg-max temp over those points. The reference's arg-max is a strided per-thread scan
//   (strict >) followed by a shared-memory tree that keeps the left operand on ties, so among
//   bit-equal maxima it returns the point minimising (bitreverse_L(k mod T), k div T), where
//   T = min(2^floor(log2 n), 512) is its block size and L = log2 T. With no eligible point every
//   pick is index 0.
//
// Two kernels, one CTA per cloud each:
//
// fps_pruned_kernel (n <= 32768, the path every config takes)
//   prologue  eligible points are counting-sorted by a 4096-cell grid (x,y Morton-interleaved, z low
//             bits) into a SoA copy -- first SMEM_CAP points in shared memory, the rest in the
//             L2-resident workspace -- so that the 128 points owned by one (warp, slot) pair form a
//             spatially tight *bucket*; each bucket keeps its bounding box and (max distance, tie
//             key, position, original index) of its current best point in shared memory;
//   loop      running distances never leave registers (64 per thread). Per iteration a lane tests
//             one bucket: LB = fma(gz,gz,fma(gx,gx,gy*gy)) on the per-axis gaps between the new
//             centre and the box. Every operation of the reference's distance is monotone in
//             |dx|,|dy|,|dz| and rounding is monotone, so LB <= d(k) for every point k of the bucket
//             *in floating point*; when LB >= the bucket's max distance no fminf can change
//             anything and the bucket is skipped -- results stay bit-identical while late
//             iterations touch only the buckets near the new centre. Touched buckets are updated
//             with packed FP32 math (sub/mul/fma.f32x2) and re-elect their best point with warp
//             REDUX; the block arg-max is a REDUX over bucket records, one smem hop across the 16
//             warps and a single __syncthreads (slots double buffered by iteration parity);
//   ties      resolved explicitly on the reference's key (bitreverse_L(k mod T), k div T).
//
// fps_flat_kernel (n > 32768, or DUSTY_FPS_FLAT=1): no pruning; points compacted in tie-break
//   order, distances in registers (or, beyond 32768 points, in the workspace).
#include <algorithm>

#include "common.cuh"

namespace dusty {
namespace fps {

constexpr int TPB = 512;
constexpr int NW = TPB / 32;
constexpr int GROUPS_MAX = 16;                   // 4 points per group per thread
constexpr int REG_CAP = TPB * 4 * GROUPS_MAX;    // 32768 points with register-resident distances
constexpr int SMEM_CAP = 18432;                  // points whose coordinates are staged in smem (216 KB)
constexpr int SMEM_BYTES = SMEM_CAP * 12;

__device__ __forceinline__ unsigned bitrev(unsigned v, int bits) { return bits ? (__brev(v) >> (32 - bits)) : 0u; }

struct WinSlot { unsigned val; unsigned e; };

// Block-wide exclusive scan of one int per thread (TPB threads); also returns the total.
__device__ __forceinline__ int block_exscan(int v, int* total, int* wsum) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
  #pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  int base = 0, tot = 0;
  #pragma unroll
  for (int w = 0; w < NW; ++w) { const int s = wsum[w]; if (w < warp) base += s; tot += s; }
  __syncthreads();
  *total = tot;
  return base + inc - v;
}

// comp layout per cloud (float, n_pad = n rounded up to 4): x[n_pad] y[n_pad] z[n_pad] oidx[n_pad]
template <bool FITS_REG>
__global__ void __launch_bounds__(TPB, 1) fps_kernel(const float* __restrict__ xyz_all, int n, int m,
                                                     int* __restrict__ idx_all, float* __restrict__ out_all,
                                                     float* __restrict__ comp_all, float* __restrict__ temp_all) {
  extern __shared__ __align__(16) float sm[];
  float* const sx = sm;
  float* const sy = sm + SMEM_CAP;
  float* const sz = sm + 2 * SMEM_CAP;
  __shared__ int wsum[NW];
  __shared__ WinSlot slots[2][NW];
  __shared__ int s_total;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long cloud = blockIdx.x;
  const float* const xyz = xyz_all + cloud * n * 3;
  int* const idx = idx_all + cloud * m;
  const int n_pad = (n + 3) & ~3;
  float* const cx = comp_all + cloud * 4 * n_pad;
  float* const cy = cx + n_pad;
  float* const cz = cy + n_pad;
  int* const co = reinterpret_cast<int*>(cz + n_pad);

  // ---- prologue: compaction in the reference's tie-break order ----
  int L = 0;
  while ((2 << L) <= n && L < 9) ++L;      // T = 2^L = min(2^floor(log2 n), 512)
  const int T = 1 << L;
  const int vt = tid < T ? (int)bitrev((unsigned)tid, L) : 0;   // the reference thread this rank stands for
  int cnt = 0;
  if (tid < T) {
    for (int k = vt; k < n; k += T) {
      const float x = xyz[3 * k], y = xyz[3 * k + 1], z = xyz[3 * k + 2];
      const float mag = fmaf(z, z, fmaf(x, x, y * y));
      cnt += !((double)mag <= 1e-3);
    }
  }
  int E;
  int off = block_exscan(cnt, &E, wsum);
  if (tid < T) {
    for (int k = vt; k < n; k += T) {
      const float x = xyz[3 * k], y = xyz[3 * k + 1], z = xyz[3 * k + 2];
      const float mag = fmaf(z, z, fmaf(x, x, y * y));
      if (!((double)mag <= 1e-3)) {
        cx[off] = x; cy[off] = y; cz[off] = z; co[off] = k;
        if (off < SMEM_CAP) { sx[off] = x; sy[off] = y; sz[off] = z; }
        ++off;
      }
    }
  }
  // pad the compact copy to a multiple of 4 so that whole groups can be loaded
  const int E_pad = (E + 3) & ~3;
  if (tid < E_pad - E) {
    const int e = E + tid;
    cx[e] = 0.f; cy[e] = 0.f; cz[e] = 0.f; co[e] = 0;
    if (e < SMEM_CAP) { sx[e] = 0.f; sy[e] = 0.f; sz[e] = 0.f; }
  }
  __threadfence_block();
  __syncthreads();

  if (tid == 0) idx[0] = 0;
  if (E == 0) {                          // nothing eligible: the reference returns index 0 throughout
    for (int j = 1 + tid; j < m; j += TPB) idx[j] = 0;
  } else {
    const int ngroups = (E_pad / 4 + TPB - 1) / TPB;      // groups per thread actually in use
    float temp[FITS_REG ? GROUPS_MAX : 1][4];
    float* const tg = FITS_REG ? nullptr : temp_all + cloud * n_pad;   // spill path for n > REG_CAP
    if (FITS_REG) {
      #pragma unroll
      for (int i = 0; i < GROUPS_MAX; ++i) {
        const int e0 = 4 * (i * TPB + tid);
        #pragma unroll
        for (int q = 0; q < 4; ++q) temp[i][q] = (e0 + q < E) ? 1e10f : -1.0f;
      }
    } else {
      for (int e = tid; e < E_pad; e += TPB) tg[e] = e < E ? 1e10f : -1.0f;
      __syncthreads();
    }

    float ccx = xyz[0], ccy = xyz[1], ccz = xyz[2];       // centre of the first iteration: point 0 as given
    for (int j = 1; j < m; ++j) {
      const f32x2 c2x = pack2(ccx, ccx), c2y = pack2(ccy, ccy), c2z = pack2(ccz, ccz);
      float best = -1.0f;
      int be = 0;

      auto visit = [&](int i, float (&tq)[4]) {
        const int g = i * TPB + tid;            // group index; points 4g..4g+3
        const int e0 = 4 * g;
        float4 px, py, pz;
        if (e0 < SMEM_CAP) {
          px = *reinterpret_cast<const float4*>(sx + e0);
          py = *reinterpret_cast<const float4*>(sy + e0);
          pz = *reinterpret_cast<const float4*>(sz + e0);
        } else {
          px = *reinterpret_cast<const float4*>(cx + e0);
          py = *reinterpret_cast<const float4*>(cy + e0);
          pz = *reinterpret_cast<const float4*>(cz + e0);
        }
        const f32x2 dx0 = sub2(pack2(px.x, px.y), c2x), dx1 = sub2(pack2(px.z, px.w), c2x);
        const f32x2 dy0 = sub2(pack2(py.x, py.y), c2y), dy1 = sub2(pack2(py.z, py.w), c2y);
        const f32x2 dz0 = sub2(pack2(pz.x, pz.y), c2z), dz1 = sub2(pack2(pz.z, pz.w), c2z);
        const f32x2 d0 = fma2(dz0, dz0, fma2(dx0, dx0, mul2(dy0, dy0)));
        const f32x2 d1 = fma2(dz1, dz1, fma2(dx1, dx1, mul2(dy1, dy1)));
        float d[4];
        unpack2(d0, d[0], d[1]);
        unpack2(d1, d[2], d[3]);
        #pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float t = fminf(d[q], tq[q]);
          tq[q] = t;
          const bool gt = t > best;             // strict: ascending compact index keeps the tie rule
          best = gt ? t : best;
          be = gt ? e0 + q : be;
        }
      };

      if (FITS_REG) {
        #pragma unroll
        for (int i = 0; i < GROUPS_MAX; ++i) {
          if (i < ngroups) visit(i, temp[i]);
        }
      } else {
        for (int i = 0; i < ngroups; ++i) {
          const int e0 = 4 * (i * TPB + tid);
          if (e0 < E_pad) {
            float4 t4 = *reinterpret_cast<float4*>(tg + e0);
            float tq[4] = {t4.x, t4.y, t4.z, t4.w};
            visit(i, tq);
            *reinterpret_cast<float4*>(tg + e0) = make_float4(tq[0], tq[1], tq[2], tq[3]);
          }
        }
      }

      // warp arg-max: value first (non-negative floats order like their bit patterns), then lowest e
      const unsigned vb = best < 0.f ? 0u : __float_as_uint(best);
      const unsigned vmax = __reduce_max_sync(0xffffffffu, vb);
      const unsigned ecand = (vb == vmax && !(best < 0.f)) ? (unsigned)be : 0xffffffffu;
      const unsigned emin = __reduce_min_sync(0xffffffffu, ecand);
      const int par = j & 1;
      if (lane == 0) { slots[par][warp].val = vmax; slots[par][warp].e = emin; }
      __syncthreads();
      const WinSlot s = slots[par][lane & (NW - 1)];
      const unsigned V = __reduce_max_sync(0xffffffffu, s.val);
      const unsigned ew = __reduce_min_sync(0xffffffffu, s.val == V ? s.e : 0xffffffffu);
      // ew is always a real point here because E >= 1 (some temp >= 0 exists)
      if (ew < (unsigned)SMEM_CAP) { ccx = sx[ew]; ccy = sy[ew]; ccz = sz[ew]; }
      else { ccx = cx[ew]; ccy = cy[ew]; ccz = cz[ew]; }
      if (tid == 0) idx[j] = (int)ew;          // compact index for now; translated below
    }
    __threadfence_block();
    __syncthreads();
    for (int j = 1 + tid; j < m; j += TPB) idx[j] = co[idx[j]];
  }

  if (out_all != nullptr) {                   // fused gather of the sampled points, (m,3) per cloud
    __threadfence_block();
    __syncthreads();
    float* const out = out_all + cloud * m * 3;
    for (int j = tid; j < m; j += TPB) {
      const int k = idx[j];
      out[3 * j] = xyz[3 * k]; out[3 * j + 1] = xyz[3 * k + 1]; out[3 * j + 2] = xyz[3 * k + 2];
    }
  }
}

__global__ void __launch_bounds__(256) gather_kernel(const float* __restrict__ points, const int* __restrict__ idx,
                                                     int c, int n, int m, float* __restrict__ out) {
  const int i = blockIdx.z, l = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < m) out[((long long)i * c + l) * m + j] = points[((long long)i * c + l) * n + idx[(long long)i * m + j]];
}

__global__ void __launch_bounds__(256) gather_grad_kernel(const float* __restrict__ grad_out, const int* __restrict__ idx,
                                                          int c, int n, int m, float* __restrict__ grad_points) {
  const int i = blockIdx.z, l = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < m)
    atomicAdd(grad_points + ((long long)i * c + l) * n + idx[(long long)i * m + j], grad_out[((long long)i * c + l) * m + j]);
}

}  // namespace fps
}  // namespace dusty

using namespace dusty;
using namespace dusty::fps;

static size_t fps_comp_bytes(int b, int n) { return align_up((size_t)b * 4 * ((n + 3) & ~3) * sizeof(float), 256); }
static size_t fps_temp_bytes(int b, int n) { return n > REG_CAP ? align_up((size_t)b * ((n + 3) & ~3) * sizeof(float), 256) : 0; }

extern "C" size_t dusty_fps_workspace_bytes(int b, int n, int m) {
  (void)m;
  if (b <= 0 || n <= 0) return 0;
  return fps_comp_bytes(b, n) + fps_temp_bytes(b, n);
}

extern "C" int dusty_fps(const float* xyz, int b, int n, int m, int32_t* idx, float* out_xyz, void* workspace,
                         size_t workspace_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (b < 0 || n <= 0 || m < 0) return fail_arg(DUSTY_EINVAL, "fps: bad sizes b=%d n=%d m=%d", b, n, m);
  if (b == 0 || m == 0) return 0;
  if (int rc = check_device()) return rc;
  if (!xyz || !idx || !workspace) return fail_arg(DUSTY_EINVAL, "fps: null pointer");
  if (!aligned16(workspace)) return fail_arg(DUSTY_EALIGN, "fps: workspace must be 16-byte aligned");
  if (workspace_bytes < dusty_fps_workspace_bytes(b, n, m))
    return fail_arg(DUSTY_ENOSPACE, "fps: workspace %zu < %zu", workspace_bytes, dusty_fps_workspace_bytes(b, n, m));
  float* comp = static_cast<float*>(workspace);
  float* temp = reinterpret_cast<float*>(static_cast<char*>(workspace) + fps_comp_bytes(b, n));
  static bool configured = false;
  if (!configured) {
    DUSTY_CUDA(cudaFuncSetAttribute(fps_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    DUSTY_CUDA(cudaFuncSetAttribute(fps_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  if (n <= REG_CAP) fps_kernel<true><<<b, TPB, SMEM_BYTES, st>>>(xyz, n, m, idx, out_xyz, comp, temp);
  else fps_kernel<false><<<b, TPB, SMEM_BYTES, st>>>(xyz, n, m, idx, out_xyz, comp, temp);
  DUSTY_AFTER_LAUNCH("fps_kernel");
  return 0;
}

extern "C" int dusty_gather_points(const float* points, const int32_t* idx, int b, int c, int n, int m, float* out,
                                   void* stream) {
  if (b < 0 || c < 0 || n <= 0 || m < 0) return fail_arg(DUSTY_EINVAL, "gather_points: bad sizes");
  if (b == 0 || c == 0 || m == 0) return 0;
  if (b > 65535 || c > 65535) return fail_arg(DUSTY_EINVAL, "gather_points: b and c must be <= 65535");
  if (int rc = check_device()) return rc;
  if (!points || !idx || !out) return fail_arg(DUSTY_EINVAL, "gather_points: null pointer");
  gather_kernel<<<dim3((m + 255) / 256, c, b), 256, 0, static_cast<cudaStream_t>(stream)>>>(points, idx, c, n, m, out);
  DUSTY_AFTER_LAUNCH("gather_kernel");
  return 0;
}

extern "C" int dusty_gather_points_grad(const float* grad_out, const int32_t* idx, int b, int c, int n, int m,
                                        float* grad_points, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (b < 0 || c < 0 || n <= 0 || m < 0) return fail_arg(DUSTY_EINVAL, "gather_points_grad: bad sizes");
  if (b == 0 || c == 0) return 0;
  if (b > 65535 || c > 65535) return fail_arg(DUSTY_EINVAL, "gather_points_grad: b and c must be <= 65535");
  if (int rc = check_device()) return rc;
  if (!grad_points) return fail_arg(DUSTY_EINVAL, "gather_points_grad: null pointer");
  DUSTY_CUDA(cudaMemsetAsync(grad_points, 0, sizeof(float) * (size_t)b * c * n, st));
  if (m == 0) return 0;
  if (!grad_out || !idx) return fail_arg(DUSTY_EINVAL, "gather_points_grad: null pointer");
  gather_grad_kernel<<<dim3((m + 255) / 256, c, b), 256, 0, st>>>(grad_out, idx, c, n, m, grad_points);
  DUSTY_AFTER_LAUNCH("gather_grad_kernel");
  return 0;
}
