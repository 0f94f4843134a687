// Real-data side of the generate-and-evaluate path for sm_100a: organised LiDAR scans -> range
// images + clouds, one streaming pass.
//
// Replaces, per batch of scans (the reference does the first half per scan on DataLoader workers in
// numpy and the second half as ~12 element-wise ATen kernels):
//   KITTIOdometry.preprocess                     datasets/kitti.py:54-67
//     depth = ||xyz||_2, mask = depth>0 & depth>min & depth<max, depth = (depth-min)/(max-min),
//     xyz /= max, everything zeroed where masked
//   KITTIOdometry.transform (to_tensor + nearest resize)   datasets/kitti.py:69-78
//   preprocess_reals                             evaluate_synthesis.py:49-57
//     lidar.invert_depth (utils/lidar.py:31-36), sigmoid_to_tanh (utils/__init__.py:70-73),
//     inv = mask*inv + (1-mask)*drop_const, xyz.flatten(2).transpose(1,2)
//
// Rounding: the numpy half is IEEE f32 (np.linalg.norm over the last axis is
// sqrt((x*x + y*y) + z*z); `-=`, `/=` with Python scalars are f32 ops, true division); the torch
// half follows ATen's CUDA element-wise kernels (division by a Python scalar = multiplication by
// the f32 reciprocal). Every operation is written with an explicit _rn intrinsic so that nothing is
// contracted. HBM traffic per output pixel: one source point (16 B for (…,4) scans) in, 4+4+12 B
// (inv, mask, points) out, +4 (depth) +12 (planar xyz) when requested.
#include "common.cuh"

namespace dusty {
namespace scan {

constexpr int TPB = 256;

struct Args {
  dusty_scan_params p;
  const float* scans;
  float* out_depth;
  float* out_mask;
  float* out_inv;
  float* out_points;
  float* out_xyz;
  int npix;
  int vec4;    // channels == 4 and scans 16-byte aligned: one LDG.128 per source point
};

__device__ __forceinline__ int nearest_src(int dst, float scale, int size) {
  const int s = (int)floorf(__fmul_rn((float)dst, scale));
  return s < size - 1 ? s : size - 1;
}

__global__ void __launch_bounds__(TPB, 6) scan_preprocess_kernel(const Args a) {
  __shared__ __align__(16) float xyz_stage[TPB / 32][384];
  const dusty_scan_params& p = a.p;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long img = blockIdx.y;
  const int npix = a.npix;
  const int pix = (blockIdx.x * TPB + tid) * 4;
  const int wpix = blockIdx.x * TPB * 4 + warp * 128;
  const bool live = pix < npix;

  float X[4] = {0, 0, 0, 0}, Y[4] = {0, 0, 0, 0}, Z[4] = {0, 0, 0, 0};
  if (live) {
    const int row = pix / p.w, col = pix - row * p.w;
    const int srow = nearest_src(row, p.scale_h, p.hs);
    const float* src_row = a.scans + ((img * p.hs + srow) * (long long)p.ws) * p.channels;
    float sx[4], sy[4], sz[4];
    #pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int scol = nearest_src(col + q, p.scale_w, p.ws);
      const float* s = src_row + (long long)scol * p.channels;
      if (a.vec4) {
        const float4 v = ldg_stream(reinterpret_cast<const float4*>(s));
        sx[q] = v.x; sy[q] = v.y; sz[q] = v.z;
      } else {
        sx[q] = __ldg(s); sy[q] = __ldg(s + 1); sz[q] = __ldg(s + 2);
      }
    }
    float dn[4], mk[4], inv[4];
    #pragma unroll
    for (int q = 0; q < 4; ++q) {
      // np.linalg.norm(xyz, ord=2, axis=2): sqrt(add.reduce(x*x)) over three contiguous values
      const float r = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(sx[q], sx[q]), __fmul_rn(sy[q], sy[q])), __fmul_rn(sz[q], sz[q])));
      const bool m = (r > 0.0f) && (r > p.min_depth) && (r < p.max_depth);
      const float d = __fdiv_rn(__fsub_rn(r, p.min_depth), p.range);
      dn[q] = m ? d : 0.0f;
      mk[q] = m ? 1.0f : 0.0f;
      X[q] = m ? __fdiv_rn(sx[q], p.max_depth) : 0.0f;
      Y[q] = m ? __fdiv_rn(sy[q], p.max_depth) : 0.0f;
      Z[q] = m ? __fdiv_rn(sz[q], p.max_depth) : 0.0f;
      // invert_depth: denormalize_minmax -> 1/depth -> normalize_minmax over [1/max, 1/min]
      const float depth = __fadd_rn(__fmul_rn(dn[q], p.range), p.min_depth);
      const float disp = __frcp_rn(depth);          // == 1.0f / depth, correctly rounded
      const float nd = __fmul_rn(__fsub_rn(disp, p.disp_lo), p.inv_disp_range);
      const float t = __fsub_rn(__fmul_rn(nd, 2.0f), 1.0f);                       // sigmoid_to_tanh
      inv[q] = __fadd_rn(__fmul_rn(mk[q], t), __fmul_rn(__fsub_rn(1.0f, mk[q]), p.drop_const));
    }
    const long long o1 = img * npix + pix;
    if (a.out_depth) stg_stream(reinterpret_cast<float4*>(a.out_depth + o1), make_float4(dn[0], dn[1], dn[2], dn[3]));
    stg_stream(reinterpret_cast<float4*>(a.out_mask + o1), make_float4(mk[0], mk[1], mk[2], mk[3]));
    stg_stream(reinterpret_cast<float4*>(a.out_inv + o1), make_float4(inv[0], inv[1], inv[2], inv[3]));
    if (a.out_xyz) {
      float* o = a.out_xyz + img * 3 * npix + pix;
      stg_stream(reinterpret_cast<float4*>(o), make_float4(X[0], X[1], X[2], X[3]));
      stg_stream(reinterpret_cast<float4*>(o + npix), make_float4(Y[0], Y[1], Y[2], Y[3]));
      stg_stream(reinterpret_cast<float4*>(o + 2 * npix), make_float4(Z[0], Z[1], Z[2], Z[3]));
    }
  }
  if (a.out_points == nullptr) return;
  if (wpix + 128 <= npix) {
    // whole warp in range (uniform): transpose through shared memory so that each STG.128 of the
    // warp covers 512 contiguous bytes of the interleaved (n,3) layout
    float4* stage = reinterpret_cast<float4*>(xyz_stage[warp]);
    stage[lane * 3] = make_float4(X[0], Y[0], Z[0], X[1]);
    stage[lane * 3 + 1] = make_float4(Y[1], Z[1], X[2], Y[2]);
    stage[lane * 3 + 2] = make_float4(Z[2], X[3], Y[3], Z[3]);
    __syncwarp();
    float4* o = reinterpret_cast<float4*>(a.out_points + (img * npix + wpix) * 3);
    stg_stream(o + lane, stage[lane]);
    stg_stream(o + lane + 32, stage[lane + 32]);
    stg_stream(o + lane + 64, stage[lane + 64]);
  } else if (live) {
    float4* o = reinterpret_cast<float4*>(a.out_points + (img * npix + pix) * 3);
    stg_stream(o, make_float4(X[0], Y[0], Z[0], X[1]));
    stg_stream(o + 1, make_float4(Y[1], Z[1], X[2], Y[2]));
    stg_stream(o + 2, make_float4(Z[2], X[3], Y[3], Z[3]));
  }
}

}  // namespace scan
}  // namespace dusty

using namespace dusty;
using namespace dusty::scan;

extern "C" int dusty_scan_preprocess(const dusty_scan_params* p, const float* scans, float* out_depth, float* out_mask,
                                     float* out_inv, float* out_points, float* out_xyz, void* stream) {
  if (!p) return fail_arg(DUSTY_EINVAL, "scan_preprocess: null params");
  if (p->b < 0 || p->hs <= 0 || p->ws <= 0 || p->h <= 0 || p->w <= 0 || p->channels < 3)
    return fail_arg(DUSTY_EINVAL, "scan_preprocess: bad shape b=%d scans=(%d,%d,%d) out=(%d,%d)", p->b, p->hs, p->ws,
                    p->channels, p->h, p->w);
  if (p->w % 4 != 0) return fail_arg(DUSTY_EINVAL, "scan_preprocess: w=%d must be a multiple of 4", p->w);
  if ((long long)p->h * p->w > 0x7fffffffLL / 4) return fail_arg(DUSTY_EINVAL, "scan_preprocess: image too large");
  if (p->b == 0) return 0;
  if (p->b > 65535) return fail_arg(DUSTY_EINVAL, "scan_preprocess: batch %d exceeds 65535", p->b);
  if (int rc = check_device()) return rc;
  if (!scans || !out_mask || !out_inv) return fail_arg(DUSTY_EINVAL, "scan_preprocess: null pointer");
  if (!aligned16(out_mask) || !aligned16(out_inv) || (out_depth && !aligned16(out_depth)) ||
      (out_points && !aligned16(out_points)) || (out_xyz && !aligned16(out_xyz)))
    return fail_arg(DUSTY_EALIGN, "scan_preprocess: outputs must be 16-byte aligned");
  if (reinterpret_cast<uintptr_t>(scans) & 3u) return fail_arg(DUSTY_EALIGN, "scan_preprocess: scans must be 4-byte aligned");
  Args a{};
  a.p = *p;
  a.scans = scans;
  a.out_depth = out_depth; a.out_mask = out_mask; a.out_inv = out_inv; a.out_points = out_points; a.out_xyz = out_xyz;
  a.npix = p->h * p->w;
  a.vec4 = (p->channels == 4 && aligned16(scans)) ? 1 : 0;
  const dim3 grid((a.npix / 4 + TPB - 1) / TPB, p->b);
  scan_preprocess_kernel<<<grid, TPB, 0, static_cast<cudaStream_t>(stream)>>>(a);
  DUSTY_AFTER_LAUNCH("scan_preprocess_kernel");
  return 0;
}
