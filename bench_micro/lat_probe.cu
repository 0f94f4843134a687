// Latency probe for the per-iteration chain of the FPS kernels on sm_100a: dependent chains of
// redux.sync, shfl.bfly, ballot, bar.sync, LDS (timed with clock64, 512-thread CTA, 1 CTA per SM).
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
constexpr int N = 2048;
__global__ void __launch_bounds__(512) probe(long long* out, unsigned* sink) {
  __shared__ unsigned sm[1024];
  const int tid = threadIdx.x, lane = tid & 31;
  unsigned v = tid * 2654435761u;
  sm[tid] = v; sm[tid + 512] = v ^ 0x5555u;
  __syncthreads();
  long long t0, t1;
  // 1. redux max chain
  t0 = clock64();
  for (int i = 0; i < N; ++i) v = __reduce_max_sync(0xffffffffu, v ^ (unsigned)i) + lane;
  t1 = clock64(); if (tid == 0) out[0] = (t1 - t0) / N;
  // 2. shfl xor butterfly max (5 steps)
  t0 = clock64();
  for (int i = 0; i < N; ++i) { unsigned w = v ^ (unsigned)i;
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) w = max(w, __shfl_xor_sync(0xffffffffu, w, o));
    v = w + lane; }
  t1 = clock64(); if (tid == 0) out[1] = (t1 - t0) / N;
  // 3. ballot chain
  t0 = clock64();
  for (int i = 0; i < N; ++i) v = __ballot_sync(0xffffffffu, (v + i) & 1) + lane;
  t1 = clock64(); if (tid == 0) out[2] = (t1 - t0) / N;
  // 4. __syncthreads chain (16 warps)
  t0 = clock64();
  for (int i = 0; i < N; ++i) { __syncthreads(); }
  t1 = clock64(); if (tid == 0) out[3] = (t1 - t0) / N;
  // 5. dependent LDS chain
  unsigned a = tid;
  t0 = clock64();
  for (int i = 0; i < N; ++i) a = sm[a & 1023];
  t1 = clock64(); if (tid == 0) out[4] = (t1 - t0) / N;
  v += a;
  // 6. STS -> bar -> LDS round trip (the slot exchange)
  t0 = clock64();
  for (int i = 0; i < N; ++i) { if (lane == 0) sm[(i & 1) * 16 + (tid >> 5)] = v; __syncthreads(); v += sm[(i & 1) * 16 + (lane & 15)]; }
  t1 = clock64(); if (tid == 0) out[5] = (t1 - t0) / N;
  // 7. full arg-max exchange as in the kernel: 2 redux, STS, bar, LDS, 2 redux, ballot, 2 shfl, LDS
  t0 = clock64();
  for (int i = 0; i < N; ++i) {
    unsigned vm = __reduce_max_sync(0xffffffffu, v);
    unsigned km = __reduce_min_sync(0xffffffffu, v == vm ? (unsigned)lane : 0xffffffffu);
    if (lane == 0) { sm[(i & 1) * 32 + (tid >> 5) * 2] = vm; sm[(i & 1) * 32 + (tid >> 5) * 2 + 1] = km; }
    __syncthreads();
    unsigned sv = sm[(i & 1) * 32 + (lane & 15) * 2], sk = sm[(i & 1) * 32 + (lane & 15) * 2 + 1];
    unsigned V = __reduce_max_sync(0xffffffffu, sv);
    unsigned K = __reduce_min_sync(0xffffffffu, sv == V ? sk : 0xffffffffu);
    int src = __ffs(__ballot_sync(0xffffffffu, sk == K)) - 1;
    unsigned e = __shfl_sync(0xffffffffu, sk, src);
    v = sm[(e + V) & 1023] + tid;
  }
  t1 = clock64(); if (tid == 0) out[6] = (t1 - t0) / N;
  // 8. global store per iteration by thread 0 + the exchange of 6
  t0 = clock64();
  for (int i = 0; i < N; ++i) { if (tid == 0) sink[i] = v; if (lane == 0) sm[(i & 1) * 16 + (tid >> 5)] = v; __syncthreads(); v += sm[(i & 1) * 16 + (lane & 15)]; }
  t1 = clock64(); if (tid == 0) out[7] = (t1 - t0) / N;
  sink[N + tid] = v;
}
int main() {
  long long* d; unsigned* s; CK(cudaMalloc(&d, 64)); CK(cudaMalloc(&s, 4 * (N + 512) * 148));
  probe<<<1, 512>>>(d, s); CK(cudaDeviceSynchronize());
  long long h[8]; CK(cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost));
  const char* names[8] = {"redux.max chain", "shfl butterfly max (5 steps)", "ballot chain", "__syncthreads (16 warps)", "dependent LDS", "STS+bar+LDS exchange", "full arg-max exchange", "exchange + global store by tid 0"};
  for (int i = 0; i < 8; ++i) printf("%-36s %lld cycles/iter\n", names[i], h[i]);
  return 0;
}
