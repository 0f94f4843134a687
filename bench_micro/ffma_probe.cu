// Micro-probe for the Chamfer inner loop on sm_100a: what do FFMA, FFMA2 (fma.rn.f32x2),
// FMNMX and FMNMX3 (3-input min.f32) sustain per SM when mixed the way the scan loop mixes them?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o ffma_probe ffma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

typedef unsigned long long u64;

__device__ __forceinline__ u64 pack2(float lo, float hi) {
  u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ float min3(float a, float b, float c) {
  float d; asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d;
}

constexpr int TPB = 256;
constexpr int NPTS = 2048;          // scanned points per pass (smem resident)

// smem layout per pair of scanned points: {x0,x1,y0,y1} {z0,z1,n0,n1}
template <int R, int MODE>
__global__ void __launch_bounds__(TPB) probe(const float4* __restrict__ scan, float* out, int passes, long long* cyc) {
  __shared__ float4 s[NPTS];         // NPTS/2 pairs * 2 float4
  for (int i = threadIdx.x; i < NPTS; i += TPB) s[i] = scan[i];
  __syncthreads();
  float ax[R], ay[R], az[R], best[R];
  #pragma unroll
  for (int r = 0; r < R; ++r) {
    ax[r] = 0.001f * (threadIdx.x * R + r); ay[r] = ax[r] * 0.5f; az[r] = ax[r] * 0.25f; best[r] = 1e30f;
  }
  long long t0 = clock64();
  for (int p = 0; p < passes; ++p) {
    if (MODE == 0) {            // scalar: 3 FFMA + 1 FMNMX per pair
      #pragma unroll 4
      for (int k = 0; k < NPTS / 2; ++k) {
        float4 q0 = s[2 * k], q1 = s[2 * k + 1];
        #pragma unroll
        for (int r = 0; r < R; ++r) {
          float t0_ = fmaf(az[r], q1.x, q1.z); t0_ = fmaf(ay[r], q0.z, t0_); t0_ = fmaf(ax[r], q0.x, t0_);
          float t1_ = fmaf(az[r], q1.y, q1.w); t1_ = fmaf(ay[r], q0.w, t1_); t1_ = fmaf(ax[r], q0.y, t1_);
          best[r] = fminf(best[r], t0_); best[r] = fminf(best[r], t1_);
        }
      }
    } else if (MODE == 1) {     // scalar FFMA + FMNMX3
      #pragma unroll 4
      for (int k = 0; k < NPTS / 2; ++k) {
        float4 q0 = s[2 * k], q1 = s[2 * k + 1];
        #pragma unroll
        for (int r = 0; r < R; ++r) {
          float t0_ = fmaf(az[r], q1.x, q1.z); t0_ = fmaf(ay[r], q0.z, t0_); t0_ = fmaf(ax[r], q0.x, t0_);
          float t1_ = fmaf(az[r], q1.y, q1.w); t1_ = fmaf(ay[r], q0.w, t1_); t1_ = fmaf(ax[r], q0.y, t1_);
          best[r] = min3(best[r], t0_, t1_);
        }
      }
    } else if (MODE == 2 || MODE == 3) {   // FFMA2 + (FMNMX3 | 2x FMNMX)
      u64 ax2[R], ay2[R], az2[R];
      #pragma unroll
      for (int r = 0; r < R; ++r) { ax2[r] = pack2(ax[r], ax[r]); ay2[r] = pack2(ay[r], ay[r]); az2[r] = pack2(az[r], az[r]); }
      #pragma unroll 4
      for (int k = 0; k < NPTS / 2; ++k) {
        float4 q0 = s[2 * k], q1 = s[2 * k + 1];
        u64 bx = pack2(q0.x, q0.y), by = pack2(q0.z, q0.w), bz = pack2(q1.x, q1.y), bn = pack2(q1.z, q1.w);
        #pragma unroll
        for (int r = 0; r < R; ++r) {
          u64 t = fma2(az2[r], bz, bn); t = fma2(ay2[r], by, t); t = fma2(ax2[r], bx, t);
          float lo, hi; unpack2(t, lo, hi);
          if (MODE == 2) best[r] = min3(best[r], lo, hi);
          else { best[r] = fminf(best[r], lo); best[r] = fminf(best[r], hi); }
        }
      }
    } else if (MODE == 4) {     // FFMA2 only (no min): pure FMA-pipe ceiling with this operand pattern
      u64 ax2[R], ay2[R], az2[R], acc[R];
      #pragma unroll
      for (int r = 0; r < R; ++r) { ax2[r] = pack2(ax[r], ax[r]); ay2[r] = pack2(ay[r], ay[r]); az2[r] = pack2(az[r], az[r]); acc[r] = pack2(0.f, 0.f); }
      #pragma unroll 4
      for (int k = 0; k < NPTS / 2; ++k) {
        float4 q0 = s[2 * k], q1 = s[2 * k + 1];
        u64 bx = pack2(q0.x, q0.y), by = pack2(q0.z, q0.w), bz = pack2(q1.x, q1.y);
        #pragma unroll
        for (int r = 0; r < R; ++r) { acc[r] = fma2(az2[r], bz, acc[r]); acc[r] = fma2(ay2[r], by, acc[r]); acc[r] = fma2(ax2[r], bx, acc[r]); }
      }
      #pragma unroll
      for (int r = 0; r < R; ++r) { float lo, hi; unpack2(acc[r], lo, hi); best[r] = lo + hi; }
    } else if (MODE == 5) {     // scalar FFMA only
      float acc0[R], acc1[R];
      #pragma unroll
      for (int r = 0; r < R; ++r) { acc0[r] = 0.f; acc1[r] = 0.f; }
      #pragma unroll 4
      for (int k = 0; k < NPTS / 2; ++k) {
        float4 q0 = s[2 * k], q1 = s[2 * k + 1];
        #pragma unroll
        for (int r = 0; r < R; ++r) {
          acc0[r] = fmaf(az[r], q1.x, acc0[r]); acc0[r] = fmaf(ay[r], q0.z, acc0[r]); acc0[r] = fmaf(ax[r], q0.x, acc0[r]);
          acc1[r] = fmaf(az[r], q1.y, acc1[r]); acc1[r] = fmaf(ay[r], q0.w, acc1[r]); acc1[r] = fmaf(ax[r], q0.y, acc1[r]);
        }
      }
      #pragma unroll
      for (int r = 0; r < R; ++r) best[r] = acc0[r] + acc1[r];
    }
  }
  long long t1 = clock64();
  float acc = 0.f;
  #pragma unroll
  for (int r = 0; r < R; ++r) acc += best[r];
  out[blockIdx.x * TPB + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int R, int MODE>
void run(const char* name, int ctas_per_sm, const float4* d_scan, float* d_out, long long* d_cyc) {
  int passes = 64;
  int grid = 148 * ctas_per_sm;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  probe<R, MODE><<<grid, TPB>>>(d_scan, d_out, 2, d_cyc);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  probe<R, MODE><<<grid, TPB>>>(d_scan, d_out, passes, d_cyc);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  long long cyc; CK(cudaMemcpy(&cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost));
  double fma_lane = (double)grid * TPB * R * NPTS * 3.0 * passes;  // scalar-equivalent FMAs
  double per_clk_sm = fma_lane / ctas_per_sm / 148.0 / (double)cyc * ctas_per_sm; // per SM per clk using CTA-0 cycles
  printf("%-34s R=%2d ctas/SM=%d  %.3f ms  %7.2f TFLOP/s  cta0 cycles=%lld  FMA/clk/SM=%.1f  eff clk=%.0f MHz\n",
         name, R, ctas_per_sm, ms, 2.0 * fma_lane / (ms * 1e-3) / 1e12, cyc, per_clk_sm, (double)cyc / (ms * 1e3));
}

int main() {
  float4* d_scan; float* d_out; long long* d_cyc;
  CK(cudaMalloc(&d_scan, NPTS * sizeof(float4)));
  CK(cudaMalloc(&d_out, 148 * 8 * TPB * sizeof(float)));
  CK(cudaMalloc(&d_cyc, 148 * 8 * sizeof(long long)));
  float4* h = (float4*)malloc(NPTS * sizeof(float4));
  for (int i = 0; i < NPTS; ++i) h[i] = make_float4(0.01f * (i % 97), 0.02f * (i % 89), -0.01f * (i % 83), 0.5f + 0.001f * i);
  CK(cudaMemcpy(d_scan, h, NPTS * sizeof(float4), cudaMemcpyHostToDevice));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s SMs=%d clock=%d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  for (int c = 1; c <= 4; c *= 2) {
    run<8, 5>("scalar FFMA only", c, d_scan, d_out, d_cyc);
    run<8, 4>("FFMA2 only", c, d_scan, d_out, d_cyc);
    run<8, 0>("scalar 3FFMA+FMNMX /pair", c, d_scan, d_out, d_cyc);
    run<8, 1>("scalar 6FFMA+FMNMX3 /2pair", c, d_scan, d_out, d_cyc);
    run<8, 2>("3FFMA2+FMNMX3 /2pair", c, d_scan, d_out, d_cyc);
    run<8, 3>("3FFMA2+2FMNMX /2pair", c, d_scan, d_out, d_cyc);
  }
  run<4, 2>("3FFMA2+FMNMX3 /2pair", 2, d_scan, d_out, d_cyc);
  run<4, 2>("3FFMA2+FMNMX3 /2pair", 4, d_scan, d_out, d_cyc);
  run<16, 2>("3FFMA2+FMNMX3 /2pair", 1, d_scan, d_out, d_cyc);
  run<16, 2>("3FFMA2+FMNMX3 /2pair", 2, d_scan, d_out, d_cyc);
  run<16, 4>("FFMA2 only", 2, d_scan, d_out, d_cyc);
  run<16, 5>("scalar FFMA only", 2, d_scan, d_out, d_cyc);
  return 0;
}
