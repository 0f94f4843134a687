// On-device MMD / COV / 1-NNA from the three Chamfer matrices.
//
// Replaces _compute_cov_mmd and _compute_nna(k=1) (reference utils/metrics/cov_mmd_1nna.py:54-65,
// 68-106) without materialising the stacked (Nr+Ng)^2 matrix, its +inf diagonal or the topk:
//   column kernel  one CTA per stacked column c: arg-min over the stacked rows with the diagonal
//                  skipped (leave-one-out 1-NN vote) and, for generated columns, the arg-min over
//                  the reference rows alone (MMD-sample / COV);
//   row kernel     one CTA per reference row: min over generated columns (MMD);
//   final kernel   one CTA: means, number of distinct covered references, confusion counts.
// Ties resolve to the lowest index.
#include "common.cuh"

namespace dusty {
namespace metrics {

constexpr int TPB = 256;

struct MinIdx { float v; int i; };

__device__ __forceinline__ MinIdx better(MinIdx a, MinIdx b) {
  return (b.v < a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}

__device__ __forceinline__ MinIdx block_argmin(MinIdx m, MinIdx* sm) {
  #pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    MinIdx t;
    t.v = __shfl_down_sync(0xffffffffu, m.v, o);
    t.i = __shfl_down_sync(0xffffffffu, m.i, o);
    m = better(m, t);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sm[warp] = m;
  __syncthreads();
  MinIdx r = sm[0];
  #pragma unroll
  for (int w = 1; w < TPB / 32; ++w) r = better(r, sm[w]);
  return r;
}

// scratch layout (ints/floats): nn_label[nr+ng], colmin[ng], colarg[ng], rowmin[nr], covered[nr]
__global__ void __launch_bounds__(TPB) column_kernel(const float* __restrict__ Mrr, const float* __restrict__ Mrg,
                                                     const float* __restrict__ Mgg, int nr, int ng,
                                                     int* __restrict__ nn_label, float* __restrict__ colmin,
                                                     int* __restrict__ covered) {
  __shared__ MinIdx sm[TPB / 32];
  const int c = blockIdx.x;
  const float inf = __int_as_float(0x7f800000);
  MinIdx all{inf, 0x7fffffff}, refonly{inf, 0x7fffffff};
  for (int r = threadIdx.x; r < nr + ng; r += TPB) {
    float v;
    if (c < nr) v = r < nr ? Mrr[(long long)r * nr + c] : Mrg[(long long)c * ng + (r - nr)];
    else v = r < nr ? Mrg[(long long)r * ng + (c - nr)] : Mgg[(long long)(r - nr) * ng + (c - nr)];
    if (c >= nr && r < nr) refonly = better(refonly, MinIdx{v, r});
    if (r != c) all = better(all, MinIdx{v, r});
  }
  all = block_argmin(all, sm);
  if (c >= nr) {
    refonly = block_argmin(refonly, sm);
    if (threadIdx.x == 0) {
      colmin[c - nr] = refonly.v;
      if (refonly.i < nr) covered[refonly.i] = 1;
    }
  }
  if (threadIdx.x == 0) nn_label[c] = (all.i < nr) ? 1 : 0;   // label of the nearest neighbour (ref = 1)
}

// Leave-one-out k-NN vote for k > 1 (_compute_nna, reference cov_mmd_1nna.py:84-90: topk(k, dim=0, largest=False),
// count of reference labels among the k nearest, pred = count / k >= 0.5). One CTA per stacked column; k rounds of a
// block arg-min that skips the diagonal and the rows already taken (ties: lowest index). Values are compared as
// stored: the reference's optional sqrt is monotone, so it cannot change which rows are the k nearest.
constexpr int KMAX = 64;
__global__ void __launch_bounds__(TPB) knn_label_kernel(const float* __restrict__ Mrr, const float* __restrict__ Mrg,
                                                        const float* __restrict__ Mgg, int nr, int ng, int k,
                                                        int* __restrict__ nn_label) {
  __shared__ MinIdx sm[TPB / 32];
  __shared__ int taken[KMAX];
  const int c = blockIdx.x;
  const float inf = __int_as_float(0x7f800000);
  int votes = 0;
  for (int round = 0; round < k; ++round) {
    MinIdx best{inf, 0x7fffffff};
    for (int r = threadIdx.x; r < nr + ng; r += TPB) {
      if (r == c) continue;
      bool used = false;
      for (int t = 0; t < round; ++t) used |= taken[t] == r;
      if (used) continue;
      float v;
      if (c < nr) v = r < nr ? Mrr[(long long)r * nr + c] : Mrg[(long long)c * ng + (r - nr)];
      else v = r < nr ? Mrg[(long long)r * ng + (c - nr)] : Mgg[(long long)(r - nr) * ng + (c - nr)];
      best = better(best, MinIdx{v, r});
    }
    best = block_argmin(best, sm);
    if (best.i == 0x7fffffff) break;              // fewer than k other rows
    votes += best.i < nr ? 1 : 0;
    __syncthreads();
    if (threadIdx.x == 0) taken[round] = best.i;
    __syncthreads();
  }
  if (threadIdx.x == 0) nn_label[c] = ((float)votes / (float)k >= 0.5f) ? 1 : 0;
}

__global__ void __launch_bounds__(TPB) row_kernel(const float* __restrict__ Mrg, int nr, int ng, float* __restrict__ rowmin) {
  __shared__ MinIdx sm[TPB / 32];
  const int r = blockIdx.x;
  MinIdx m{__int_as_float(0x7f800000), 0x7fffffff};
  for (int j = threadIdx.x; j < ng; j += TPB) m = better(m, MinIdx{Mrg[(long long)r * ng + j], j});
  m = block_argmin(m, sm);
  if (threadIdx.x == 0) rowmin[r] = m.v;
}

__device__ __forceinline__ double block_sum_d(double v, double* sm) {
  #pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  double s = 0.0;
  #pragma unroll
  for (int w = 0; w < TPB / 32; ++w) s += sm[w];
  return s;
}

__global__ void __launch_bounds__(TPB) final_kernel(int nr, int ng, const int* __restrict__ nn_label,
                                                    const float* __restrict__ colmin, const float* __restrict__ rowmin,
                                                    const int* __restrict__ covered, float* __restrict__ out7) {
  __shared__ double sm[TPB / 32];
  double s_row = 0, s_col = 0, cov = 0, tp = 0, fp = 0, fn = 0, tn = 0;
  for (int i = threadIdx.x; i < nr; i += TPB) { s_row += rowmin[i]; cov += covered[i]; }
  for (int j = threadIdx.x; j < ng; j += TPB) s_col += colmin[j];
  for (int c = threadIdx.x; c < nr + ng; c += TPB) {
    const int pred = nn_label[c], label = c < nr ? 1 : 0;
    tp += pred & label; fp += pred & (1 - label); fn += (1 - pred) & label; tn += (1 - pred) & (1 - label);
  }
  s_row = block_sum_d(s_row, sm); s_col = block_sum_d(s_col, sm); cov = block_sum_d(cov, sm);
  tp = block_sum_d(tp, sm); fp = block_sum_d(fp, sm); fn = block_sum_d(fn, sm); tn = block_sum_d(tn, sm);
  if (threadIdx.x == 0) {
    out7[0] = (float)(s_row / nr);
    out7[1] = (float)(s_col / ng);
    out7[2] = (float)cov;                 // number of distinct covered references (exact integer)
    out7[3] = (float)tp; out7[4] = (float)fp; out7[5] = (float)fn; out7[6] = (float)tn;
  }
}

// Fused path: the matrix kernel's epilogue left packed (value bits << 32 | stacked index) minima, one set of
// 3 (nr+ng) keys per row shard (after the all-gather: `shards` sets). Reduce over the shards and fill the same
// scratch vectors the column / row kernels produce, so that final_kernel -- and hence every score -- is shared.
__global__ void __launch_bounds__(TPB) keys_kernel(const unsigned long long* __restrict__ keys, int shards, int nr, int ng,
                                                   int* __restrict__ nn_label, float* __restrict__ colmin,
                                                   float* __restrict__ rowmin, int* __restrict__ covered) {
  const int n = nr + ng;
  const int c = blockIdx.x * TPB + threadIdx.x;
  if (c >= n) return;
  unsigned long long all = ~0ull, other = ~0ull;
  const int which = c < nr ? 2 : 1;            // reference clouds: nearest generated one; generated: nearest reference
  for (int s = 0; s < shards; ++s) {
    const unsigned long long* k = keys + (size_t)s * 3 * n;
    all = min(all, k[c]);
    other = min(other, k[(size_t)which * n + c]);
  }
  nn_label[c] = ((int)(unsigned)all < nr) ? 1 : 0;
  const float v = __uint_as_float((unsigned)(other >> 32));
  if (c < nr) {
    rowmin[c] = v;
  } else {
    colmin[c - nr] = v;
    const int ref = (int)(unsigned)other;
    if (ref >= 0 && ref < nr) covered[ref] = 1;
  }
}

// (G, cap, n) gathered compact row blocks of the cyclic deal (block g, row r = global row g + r G, entries j >= i
// valid) -> full symmetric (n, n) matrix, every entry read from the shard that computed it.
__global__ void __launch_bounds__(TPB) symmetric_kernel(const float* __restrict__ blocks, int G, int cap, int n,
                                                        float* __restrict__ out, long long ldo) {
  const int j = blockIdx.x * TPB + threadIdx.x, i = blockIdx.y;
  if (j >= n) return;
  const int r = i <= j ? i : j, c = i <= j ? j : i;
  out[(long long)i * ldo + j] = blocks[((size_t)(r % G) * cap + r / G) * n + c];
}

}  // namespace metrics
}  // namespace dusty

using namespace dusty;
using namespace dusty::metrics;

extern "C" size_t dusty_cov_mmd_1nna_workspace_bytes(int nr, int ng) {
  if (nr <= 0 || ng <= 0) return 0;
  return align_up(sizeof(int) * ((size_t)nr + ng) + sizeof(float) * ng + sizeof(float) * nr + sizeof(int) * nr, 256);
}

extern "C" int dusty_cov_mmd_1nna_finalize(const float* Mrr, const float* Mrg, const float* Mgg, int nr, int ng,
                                           float* out7, void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (nr <= 0 || ng <= 0) return fail_arg(DUSTY_EINVAL, "cov_mmd_1nna: nr=%d ng=%d must be positive", nr, ng);
  if (int rc = check_device()) return rc;
  if (!Mrr || !Mrg || !Mgg || !out7 || !workspace) return fail_arg(DUSTY_EINVAL, "cov_mmd_1nna: null pointer");
  if (workspace_bytes < dusty_cov_mmd_1nna_workspace_bytes(nr, ng)) return fail_arg(DUSTY_ENOSPACE, "cov_mmd_1nna: workspace too small");
  int* nn_label = static_cast<int*>(workspace);
  float* colmin = reinterpret_cast<float*>(nn_label + nr + ng);
  float* rowmin = colmin + ng;
  int* covered = reinterpret_cast<int*>(rowmin + nr);
  DUSTY_CUDA(cudaMemsetAsync(covered, 0, sizeof(int) * nr, st));
  column_kernel<<<nr + ng, TPB, 0, st>>>(Mrr, Mrg, Mgg, nr, ng, nn_label, colmin, covered);
  DUSTY_AFTER_LAUNCH("metrics column_kernel");
  row_kernel<<<nr, TPB, 0, st>>>(Mrg, nr, ng, rowmin);
  DUSTY_AFTER_LAUNCH("metrics row_kernel");
  final_kernel<<<1, TPB, 0, st>>>(nr, ng, nn_label, colmin, rowmin, covered, out7);
  DUSTY_AFTER_LAUNCH("metrics final_kernel");
  return 0;
}

extern "C" int dusty_cov_mmd_knna_finalize(const float* Mrr, const float* Mrg, const float* Mgg, int nr, int ng, int k,
                                          float* out7, void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (k < 1 || k > KMAX) return fail_arg(DUSTY_EINVAL, "cov_mmd_knna: k=%d must be in [1, %d]", k, KMAX);
  if (int rc = dusty_cov_mmd_1nna_finalize(Mrr, Mrg, Mgg, nr, ng, out7, workspace, workspace_bytes, stream)) return rc;
  if (k == 1) return 0;
  int* nn_label = static_cast<int*>(workspace);          // same scratch layout as the k = 1 call that just ran
  float* colmin = reinterpret_cast<float*>(nn_label + nr + ng);
  float* rowmin = colmin + ng;
  int* covered = reinterpret_cast<int*>(rowmin + nr);
  knn_label_kernel<<<nr + ng, TPB, 0, st>>>(Mrr, Mrg, Mgg, nr, ng, k, nn_label);
  DUSTY_AFTER_LAUNCH("metrics knn_label_kernel");
  final_kernel<<<1, TPB, 0, st>>>(nr, ng, nn_label, colmin, rowmin, covered, out7);
  DUSTY_AFTER_LAUNCH("metrics final_kernel");
  return 0;
}

extern "C" int dusty_cov_mmd_1nna_from_keys(const uint64_t* keys, int shards, int nr, int ng, float* out7, void* workspace,
                                            size_t workspace_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (nr <= 0 || ng <= 0 || shards <= 0) return fail_arg(DUSTY_EINVAL, "cov_mmd_1nna_from_keys: nr=%d ng=%d shards=%d must be positive", nr, ng, shards);
  if (int rc = check_device()) return rc;
  if (!keys || !out7 || !workspace) return fail_arg(DUSTY_EINVAL, "cov_mmd_1nna_from_keys: null pointer");
  if (workspace_bytes < dusty_cov_mmd_1nna_workspace_bytes(nr, ng)) return fail_arg(DUSTY_ENOSPACE, "cov_mmd_1nna_from_keys: workspace too small");
  int* nn_label = static_cast<int*>(workspace);
  float* colmin = reinterpret_cast<float*>(nn_label + nr + ng);
  float* rowmin = colmin + ng;
  int* covered = reinterpret_cast<int*>(rowmin + nr);
  DUSTY_CUDA(cudaMemsetAsync(covered, 0, sizeof(int) * nr, st));
  keys_kernel<<<(nr + ng + TPB - 1) / TPB, TPB, 0, st>>>(reinterpret_cast<const unsigned long long*>(keys), shards, nr, ng,
                                                          nn_label, colmin, rowmin, covered);
  DUSTY_AFTER_LAUNCH("metrics keys_kernel");
  final_kernel<<<1, TPB, 0, st>>>(nr, ng, nn_label, colmin, rowmin, covered, out7);
  DUSTY_AFTER_LAUNCH("metrics final_kernel");
  return 0;
}

extern "C" int dusty_symmetric_from_shards(const float* blocks, int shards, int cap, int n, float* out, long long ldo,
                                           void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n < 0 || shards <= 0 || cap < 0 || (long long)shards * cap < n || ldo < n || n > 65535)
    return fail_arg(DUSTY_EINVAL, "symmetric_from_shards: shards=%d cap=%d n=%d ldo=%lld", shards, cap, n, ldo);
  if (n == 0) return 0;
  if (int rc = check_device()) return rc;
  if (!blocks || !out) return fail_arg(DUSTY_EINVAL, "symmetric_from_shards: null pointer");
  symmetric_kernel<<<dim3((n + TPB - 1) / TPB, n), TPB, 0, st>>>(blocks, shards, cap, n, out, ldo);
  DUSTY_AFTER_LAUNCH("metrics symmetric_kernel");
  return 0;
}
