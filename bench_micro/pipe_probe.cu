// Register-only pipe probe for the Chamfer search loop on sm_100a: how much of the FMA pipe survives when
// the three FFMA2 of a (row, candidate pair) are followed by the ALU work that consumes them?
// Eight independent rows per thread exactly as in nn_kernel<8,...>; the candidate pairs come from a 2048-point
// shared-memory tile in scan format (two LDS.128 per 24 FFMA2), nothing else touches memory in the timed loop.
// One line per (mode, CTAs per SM).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o pipe_probe pipe_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

typedef unsigned long long u64;

__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 vfma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float vfma(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float vmin3(float a, float b, float c) { float d; asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float vmin2(float a, float b) { float d; asm volatile("min.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ int vimin2(int a, int b) { int d; asm volatile("min.s32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ unsigned vlop3(unsigned a, unsigned b, unsigned c) { unsigned d; asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ int viadd3(int a, int b, int c) { int d; asm volatile("{ .reg .s32 t; add.s32 t, %1, %2; add.s32 %0, t, %3; }" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ unsigned vhmin2(unsigned a, unsigned b) { unsigned d; asm volatile("min.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }

constexpr int TPB = 256;
constexpr int R = 8;
constexpr int INNER = 16;       // candidate pairs per "chunk"

enum Mode { FFMA2_ONLY = 0, FMNMX3, FMNMX_X2, IMIN_FUSED, LOP3, IADD3, FMNMX3_INDEP, RATIO_6_1, WITH_BOOKKEEPING, SCALAR_FMNMX3,
            FFMA2_CHAIN, HMNMX2, BOOKKEEPING_TOP2, ROWPAIR, ROWPAIR_TOP2, PIPELINED, MIN_AFTER_GROUP1, MIN_LO_ONLY, MIN_OF_CHAIN_ACC, MIN_LO_PLUS_INDEP, CROSS, CARRY, CROSS_TOP2, MIN_HI_ONLY, LO_LO, HI_HI, N_MODES };
const char* kNames[] = {"3 FFMA2 chained + 1 FMNMX3 per 96", "3 FFMA2 + FMNMX3(best,lo,hi)", "3 FFMA2 + 2 FMNMX", "3 FFMA2 + int min(min()) [VIMNMX3?]",
                        "3 FFMA2 + LOP3", "3 FFMA2 + 2 IADD", "3 FFMA2 + FMNMX3 on other registers", "6 FFMA2 + FMNMX3 (two candidates pairs summed)",
                        "3 FFMA2 + FMNMX3 + 3 ops / row / 16 pairs", "6 FFMA + FMNMX3 (scalar)", "3 FFMA2 chained through the accumulator",
                        "3 FFMA2 + bf16x2 min", "3 FFMA2 + FMNMX3 + 5 ops / row / 16 pairs (top-2)",
                        "row pairs packed, candidate broadcast: 6 FFMA2 + 2 FMNMX3", "row pairs packed + 5 ops / row / 16 pairs (top-2)",
                        "3 FFMA2 + FMNMX3 of the previous pair (software pipelined)", "FMNMX3 of the previous pair issued inside group 1",
                        "3 FFMA2 + FMNMX3(best,lo,lo)", "3 FFMA2 chained in place + FMNMX3(aux,acc.lo,acc.hi)", "3 FFMA2 + FMNMX3(best,lo,other reg)",
                        "two pairs per row: FMNMX3(best,loA,hiB), FMNMX3(best,loB,hiA)", "FMNMX3(best,lo,hi of the previous pair)",
                        "crossed halves + 5 ops / row / 16 pairs (top-2)",
                        "3 FFMA2 + FMNMX3(best,hi,hi)", "two pairs per row: one FMNMX3(best,loA,loB)", "two pairs per row: one FMNMX3(best,hiA,hiB)"};

template <int MODE>
__global__ void __launch_bounds__(TPB, 2) probe(float* out, int chunks, long long* cyc, float seed) {
  __shared__ float4 tile[2048];
  for (int i = threadIdx.x; i < 2048; i += TPB) tile[i] = make_float4(seed * (i % 97), seed * (i % 89), -seed * (i % 83), 0.5f + seed * i);
  __syncthreads();
  u64 ax[R], ay[R], az[R];
  float best[R], second[R], aux[R];
  int cid[R];
  #pragma unroll
  for (int r = 0; r < R; ++r) {
    const float v = seed * (threadIdx.x * R + r + 1);
    const float w = (MODE == ROWPAIR || MODE == ROWPAIR_TOP2) ? 1.5f * v : v;       // distinct halves: a real row pair
    ax[r] = pack2(v, w); ay[r] = pack2(0.5f * v, 0.5f * w); az[r] = pack2(0.25f * v, 0.25f * w);
    best[r] = 1e30f; second[r] = 1e30f; aux[r] = 1e30f; cid[r] = 0;
  }
  u64 acc[R];
  #pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = ax[r];
  const long long t0 = clock64();
  for (int c = 0; c < chunks; ++c) {
    float cm[R];
    #pragma unroll
    for (int r = 0; r < R; ++r) cm[r] = best[r];
    const float4* cp = tile + (c & 63) * (2 * INNER);
    #pragma unroll
    for (int k = 0; k < INNER; ++k) {
      const float4 q0 = cp[2 * k], q1 = cp[2 * k + 1];
      const u64 bx = pack2(q0.x, q0.y), by = pack2(q0.z, q0.w), bz = pack2(q1.x, q1.y), bn = pack2(q1.z, q1.w);
      if (MODE == ROWPAIR || MODE == ROWPAIR_TOP2) {
        // rows 2p, 2p+1 share one FFMA2; the candidate is the scalar (broadcast) operand
        const u64 n0 = pack2(q1.z, q1.z), n1 = pack2(q1.w, q1.w);
        #pragma unroll
        for (int p = 0; p < R / 2; ++p) {
          u64 s0 = vfma2(az[p], pack2(q1.x, q1.x), n0);
          u64 s1 = vfma2(az[p], pack2(q1.y, q1.y), n1);
          s0 = vfma2(ay[p], pack2(q0.z, q0.z), s0);
          s1 = vfma2(ay[p], pack2(q0.w, q0.w), s1);
          s0 = vfma2(ax[p], pack2(q0.x, q0.x), s0);
          s1 = vfma2(ax[p], pack2(q0.y, q0.y), s1);
          float l0, h0, l1, h1;
          unpack2(s0, l0, h0); unpack2(s1, l1, h1);
          cm[2 * p] = vmin3(cm[2 * p], l0, l1);
          cm[2 * p + 1] = vmin3(cm[2 * p + 1], h0, h1);
        }
        continue;
      }
      if (MODE == CROSS || MODE == CROSS_TOP2 || MODE == LO_LO || MODE == HI_HI) {
        if (k & 1) continue;                       // pairs k and k + 1 together
        const float4 p0 = cp[2 * k + 2], p1 = cp[2 * k + 3];
        const u64 cx = pack2(p0.x, p0.y), cy = pack2(p0.z, p0.w), cz = pack2(p1.x, p1.y), cn = pack2(p1.z, p1.w);
        #pragma unroll
        for (int r = 0; r < R; ++r) {
          u64 sa = vfma2(az[r], bz, bn); sa = vfma2(ay[r], by, sa); sa = vfma2(ax[r], bx, sa);
          u64 sb = vfma2(az[r], cz, cn); sb = vfma2(ay[r], cy, sb); sb = vfma2(ax[r], cx, sb);
          float la, ha, lb, hb;
          unpack2(sa, la, ha); unpack2(sb, lb, hb);
          if (MODE == LO_LO) { cm[r] = vmin3(cm[r], la, lb); aux[r] = vmin3(aux[r], second[r], q1.w); }
          else if (MODE == HI_HI) { cm[r] = vmin3(cm[r], ha, hb); aux[r] = vmin3(aux[r], second[r], q1.w); }
          else { cm[r] = vmin3(cm[r], la, hb); cm[r] = vmin3(cm[r], lb, ha); }
        }
        continue;
      }
      if (MODE == MIN_AFTER_GROUP1) {
        u64 s[R];
        #pragma unroll
        for (int r = 0; r < R; ++r) {
          s[r] = vfma2(az[r], bz, bn);
          float lo, hi; unpack2(acc[r], lo, hi);
          cm[r] = vmin3(cm[r], lo, hi);
        }
        #pragma unroll
        for (int r = 0; r < R; ++r) s[r] = vfma2(ay[r], by, s[r]);
        #pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = vfma2(ax[r], bx, s[r]);
        continue;
      }
      #pragma unroll
      for (int r = 0; r < R; ++r) {
        if (MODE == CARRY) {
          u64 s = vfma2(az[r], bz, bn); s = vfma2(ay[r], by, s); s = vfma2(ax[r], bx, s);
          float lo, hi; unpack2(s, lo, hi);
          cm[r] = vmin3(cm[r], lo, aux[r]);
          aux[r] = hi;
          continue;
        }
        if (MODE == PIPELINED) {
          float lo, hi; unpack2(acc[r], lo, hi);
          cm[r] = vmin3(cm[r], lo, hi);
          u64 s = vfma2(az[r], bz, bn); s = vfma2(ay[r], by, s); acc[r] = vfma2(ax[r], bx, s);
          continue;
        }
        if (MODE == SCALAR_FMNMX3) {
          float ax_, d0, ay_, az_, b0, b1, b2, b3, b4, b5, b6, b7;
          unpack2(ax[r], ax_, d0); unpack2(ay[r], ay_, d0); unpack2(az[r], az_, d0);
          unpack2(bx, b0, b1); unpack2(by, b2, b3); unpack2(bz, b4, b5); unpack2(bn, b6, b7);
          float lo = vfma(az_, b4, b6); lo = vfma(ay_, b2, lo); lo = vfma(ax_, b0, lo);
          float hi = vfma(az_, b5, b7); hi = vfma(ay_, b3, hi); hi = vfma(ax_, b1, hi);
          cm[r] = vmin3(cm[r], lo, hi);
          continue;
        }
        if (MODE == FFMA2_CHAIN || MODE == FMNMX3_INDEP || MODE == FFMA2_ONLY || MODE == MIN_OF_CHAIN_ACC) {
          acc[r] = vfma2(az[r], bz, acc[r]); acc[r] = vfma2(ay[r], by, acc[r]); acc[r] = vfma2(ax[r], bx, acc[r]);
          if (MODE == FMNMX3_INDEP) aux[r] = vmin3(aux[r], second[r], q1.w);
          if (MODE == MIN_OF_CHAIN_ACC) { float lo, hi; unpack2(acc[r], lo, hi); aux[r] = vmin3(aux[r], lo, hi); }
          if (MODE == FFMA2_ONLY && r == 0 && (k & 3) == 0) aux[0] = vmin3(aux[0], q1.z, q1.w);     // one in 96
          continue;
        }
        u64 s = vfma2(az[r], bz, bn);
        s = vfma2(ay[r], by, s);
        s = vfma2(ax[r], bx, s);
        if (MODE == RATIO_6_1) {
          u64 s2 = vfma2(az[r], bx, bn);
          s2 = vfma2(ay[r], bz, s2);
          s = vfma2(ax[r], by, s2 ^ (s & 1ull));     // keeps both chains alive with one cheap op... (LOP on the pair)
        }
        float lo, hi;
        unpack2(s, lo, hi);
        if (MODE == FMNMX3 || MODE == RATIO_6_1 || MODE == WITH_BOOKKEEPING || MODE == BOOKKEEPING_TOP2) cm[r] = vmin3(cm[r], lo, hi);
        else if (MODE == MIN_LO_ONLY) cm[r] = vmin3(cm[r], lo, lo);
        else if (MODE == MIN_HI_ONLY) cm[r] = vmin3(cm[r], hi, hi);
        else if (MODE == MIN_LO_PLUS_INDEP) cm[r] = vmin3(cm[r], lo, second[r]);
        else if (MODE == FMNMX_X2) { cm[r] = vmin2(cm[r], lo); cm[r] = vmin2(cm[r], hi); }
        else if (MODE == IMIN_FUSED) cm[r] = __int_as_float(min(__float_as_int(cm[r]), min(__float_as_int(lo), __float_as_int(hi))));
        else if (MODE == LOP3) cm[r] = __uint_as_float(vlop3(__float_as_uint(cm[r]), __float_as_uint(lo), __float_as_uint(hi)));
        else if (MODE == IADD3) cm[r] = __int_as_float(viadd3(__float_as_int(cm[r]), __float_as_int(lo), __float_as_int(hi)));
        else if (MODE == HMNMX2) cm[r] = __uint_as_float(vhmin2(__float_as_uint(cm[r]), vhmin2(__float_as_uint(lo), __float_as_uint(hi))));
      }
    }
    #pragma unroll
    for (int r = 0; r < R; ++r) {
      if (MODE == WITH_BOOKKEEPING) {
        const bool better = cm[r] < best[r];
        best[r] = better ? cm[r] : best[r];
        cid[r] = better ? c : cid[r];
      } else if (MODE == BOOKKEEPING_TOP2 || MODE == ROWPAIR_TOP2 || MODE == CROSS_TOP2) {
        const bool better = cm[r] < best[r];
        const float loser = better ? best[r] : cm[r];
        second[r] = fminf(second[r], loser);
        best[r] = better ? cm[r] : best[r];
        cid[r] = better ? c : cid[r];
      } else {
        best[r] = cm[r];
      }
    }
    if (MODE == WITH_BOOKKEEPING || MODE == BOOKKEEPING_TOP2 || MODE == ROWPAIR_TOP2 || MODE == CROSS_TOP2) {       // new chunk: the running minimum restarts
      #pragma unroll
      for (int r = 0; r < R; ++r) best[r] += 0.0f;
    }
  }
  const long long t1 = clock64();
  float a = 0.f;
  #pragma unroll
  for (int r = 0; r < R; ++r) { float lo, hi; unpack2(acc[r], lo, hi); a += best[r] + second[r] + aux[r] + (float)cid[r] + lo + hi; }
  out[blockIdx.x * TPB + threadIdx.x] = a;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(int ctas_per_sm, float* d_out, long long* d_cyc) {
  const int chunks = 4096;
  const int grid = 148 * ctas_per_sm;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  probe<MODE><<<grid, TPB>>>(d_out, 64, d_cyc, 1e-3f);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  probe<MODE><<<grid, TPB>>>(d_out, chunks, d_cyc, 1e-3f);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  long long cyc; CK(cudaMemcpy(&cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost));
  const double per_row_pair = (MODE == RATIO_6_1) ? 6.0 : 3.0;       // FFMA2 (or scalar-pair equivalents) per row and iteration
  const double fma_lane = (double)grid * TPB * R * INNER * (double)chunks * per_row_pair * 2.0;   // scalar FMAs
  printf("%-52s ctas/SM=%d  %8.3f ms  %6.2f TFLOP/s  FMA/clk/SM=%6.1f (of 128)  cta0 clk=%.0f MHz\n", kNames[MODE], ctas_per_sm, ms,
         2.0 * fma_lane / (ms * 1e-3) / 1e12, fma_lane / 148.0 / ((double)cyc), (double)cyc / (ms * 1e3));
}

static int g_only[64]; static int g_nonly = 0;      // modes named on the command line (default: all)

template <int MODE>
void sweep(float* d_out, long long* d_cyc) {
  if constexpr (MODE < N_MODES) {
    bool want = g_nonly == 0;
    for (int i = 0; i < g_nonly; ++i) want |= g_only[i] == MODE;
    if (want) {
      if (g_nonly == 0) run<MODE>(1, d_out, d_cyc);
      run<MODE>(2, d_out, d_cyc);
    }
    sweep<MODE + 1>(d_out, d_cyc);
  }
}

int main(int argc, char** argv) {
  for (int i = 1; i < argc && g_nonly < 64; ++i) g_only[g_nonly++] = atoi(argv[i]);
  float* d_out; long long* d_cyc;
  CK(cudaMalloc(&d_out, 148 * 4 * TPB * sizeof(float)));
  CK(cudaMalloc(&d_cyc, 148 * 4 * sizeof(long long)));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s SMs=%d clock=%d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  sweep<0>(d_out, d_cyc);
  return 0;
}
