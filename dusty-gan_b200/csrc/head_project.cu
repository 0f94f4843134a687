// Fused point-drop head + inverse spherical projection (+ valid-point compaction) for sm_100a.
//
// One streaming pass replaces ~30 element-wise ATen kernels of the reference:
//   GumbelSigmoid.logistic_noise / forward      models/dusty.py:30-59
//   DUSty1.maskout / DUSty2.maskout (eval)      models/dusty.py:77-91, 107-127
//   tanh_to_sigmoid(...).clamp_(0, 1)           utils/__init__.py:76-79, evaluate_synthesis.py:60
//   Coordinate.inv_to_xyz (revert_depth,        utils/lidar.py:23-29, 38-47, 49-56, 61-68
//     normalize/denormalize_minmax, pol_to_xyz)
//   xyz.flatten(2).transpose(1,2) + .contiguous() copies   evaluate_synthesis.py:62, fps/...py:88
//
// Every intermediate is rounded exactly where the reference's separate kernels round it
// (__fadd_rn/__fmul_rn/__fdiv_rn block FMA contraction; expf/logf are the same libdevice
// functions ATen's CUDA kernels call), so masks and points are bit-identical to the reference
// run on the same GPU. HBM traffic per pixel: read depth + confidence, write mask + depth + xyz
// = 28 B (DUSty-I) / 36 B (DUSty-II); the noise map and the trig table are small and L2 resident.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace dusty {
namespace head {

constexpr int TPB = 256;
constexpr int GROUP = TPB * 4;          // pixels per CTA and iteration (1024)
constexpr int MAX_ITERS = 4;

struct GateDev {
  int mode;
  const float* a;
  const float* b;
  long long bstride;
  int pstride;
};

struct Args {
  dusty_head_params p;
  GateDev gp, gi;
  const float* depth;
  const float* conf;
  const float* trig;
  float* out_mask;
  float* out_depth;
  float* out_points;
  int* out_count;
  int* out_index;
  float* out_compact;
  unsigned* seg_state;      // one word per CTA: 0x80000000 | count once published
  int npix;                 // h*w
  int segs_per_image;
};

__device__ __forceinline__ float logistic_from_uniform(float u1, float u2, float eps) {
  const float l1 = logf(__fadd_rn(u1, eps));
  const float l2 = logf(__fadd_rn(u2, eps));
  return -logf(__fadd_rn(__fdiv_rn(l1, l2), eps));
}

// GumbelSigmoid.forward for one logit; returns the straight-through forward value.
__device__ __forceinline__ float gate_value(float logit, int mode, float na, float nb, const dusty_head_params& p) {
  if (mode == DUSTY_NOISE_NONE) return logit > 0.0f ? 1.0f : 0.0f;
  const float l = mode == DUSTY_NOISE_UNIFORM ? logistic_from_uniform(na, nb, p.eps) : na;
  const float x = __fmul_rn(__fadd_rn(logit, l), p.inv_tau);
  const float soft = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));
  const float hard = soft > p.threshold ? 1.0f : 0.0f;
  return __fadd_rn(__fsub_rn(hard, soft), soft);
}

__device__ __forceinline__ float4 load_noise(const GateDev& g, const float* base, long long img, int pix) {
  // base = g.a or g.b
  const float* q = base + img * g.bstride;
  if (g.pstride == 0) { const float v = q[0]; return make_float4(v, v, v, v); }
  return *reinterpret_cast<const float4*>(q + pix);
}

// inv in [0,1] (NaN passes through) -> normalised range d and validity
__device__ __forceinline__ float range_from_inv(float inv, const dusty_head_params& p, bool& valid) {
  valid = fabsf(inv) > p.tol;
  const float disp = __fadd_rn(__fmul_rn(inv, p.disp_scale), p.disp_shift);
  const float depth = __fdiv_rn(1.0f, disp);
  const float nrm = __fmul_rn(__fsub_rn(depth, p.min_depth), p.inv_range);
  float d = __fadd_rn(__fmul_rn(nrm, p.range), p.min_depth);
  d = __fmul_rn(d, p.inv_max_depth);
  return __fmul_rn(d, valid ? 1.0f : 0.0f);
}

// ITERS 4-pixel groups per thread: more independent 128-bit loads in flight per thread against
// finer CTAs, fewer registers and more resident warps; see pick_iters for the measurement.
template <int C, bool COMPACT, int ITERS>
__global__ void __launch_bounds__(TPB) head_project_kernel(const Args a) {
  constexpr int SEG = GROUP * ITERS;
  __shared__ int wsum[2][TPB / 32];
  __shared__ int s_base;
  __shared__ __align__(16) float xyz_stage[TPB / 32][384];     // per-warp transpose buffer for interleaved points
  const dusty_head_params& p = a.p;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long img = blockIdx.x / a.segs_per_image;
  const int seg = blockIdx.x - (int)img * a.segs_per_image;
  const int npix = a.npix;

  const float* depth = a.depth + img * npix;
  const float* conf0 = a.conf + img * C * npix;
  float* omask0 = a.out_mask + img * C * npix;
  float* odepth = a.out_depth + img * npix;

  float4 xs[ITERS], ys[ITERS], zs[ITERS];
  unsigned vbits[ITERS];
  int tcount = 0;

  // all streaming loads of this thread first: 2-3 independent 128-bit requests per iteration in flight
  float4 dvs[ITERS], cvs[ITERS], c1s[C == 2 ? ITERS : 1];
  #pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int pix = seg * SEG + it * (TPB * 4) + tid * 4;
    if (pix < npix) {
      dvs[it] = ldg_stream(reinterpret_cast<const float4*>(depth + pix));
      cvs[it] = ldg_stream(reinterpret_cast<const float4*>(conf0 + pix));
      if (C == 2) c1s[it] = ldg_stream(reinterpret_cast<const float4*>(conf0 + npix + pix));
    }
  }

  #pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int pix = seg * SEG + it * (TPB * 4) + tid * 4;
    vbits[it] = 0;
    if (pix < npix) {
      const float4 dv = dvs[it];
      const float4 cv = cvs[it];
      float4 na = make_float4(0, 0, 0, 0), nb = na;
      if (a.gp.mode != DUSTY_NOISE_NONE) na = load_noise(a.gp, a.gp.a, img, pix);
      if (a.gp.mode == DUSTY_NOISE_UNIFORM) nb = load_noise(a.gp, a.gp.b, img, pix);
      float mp[4] = {gate_value(cv.x, a.gp.mode, na.x, nb.x, p), gate_value(cv.y, a.gp.mode, na.y, nb.y, p),
                     gate_value(cv.z, a.gp.mode, na.z, nb.z, p), gate_value(cv.w, a.gp.mode, na.w, nb.w, p)};
      float mk[4] = {mp[0], mp[1], mp[2], mp[3]};
      stg_stream(reinterpret_cast<float4*>(omask0 + pix), make_float4(mp[0], mp[1], mp[2], mp[3]));
      if (C == 2) {
        const float4 c1 = c1s[C == 2 ? it : 0];
        float4 ia = make_float4(0, 0, 0, 0), ib = ia;
        if (a.gi.mode != DUSTY_NOISE_NONE) ia = load_noise(a.gi, a.gi.a, img, pix);
        if (a.gi.mode == DUSTY_NOISE_UNIFORM) ib = load_noise(a.gi, a.gi.b, img, pix);
        const float mi[4] = {gate_value(c1.x, a.gi.mode, ia.x, ib.x, p), gate_value(c1.y, a.gi.mode, ia.y, ib.y, p),
                             gate_value(c1.z, a.gi.mode, ia.z, ib.z, p), gate_value(c1.w, a.gi.mode, ia.w, ib.w, p)};
        stg_stream(reinterpret_cast<float4*>(omask0 + npix + pix), make_float4(mi[0], mi[1], mi[2], mi[3]));
        #pragma unroll
        for (int q = 0; q < 4; ++q) mk[q] = __fmul_rn(mp[q], mi[q]);
      }
      const float din[4] = {dv.x, dv.y, dv.z, dv.w};
      float dout[4], rng[4];
      #pragma unroll
      for (int q = 0; q < 4; ++q) {
        // mask * depth + (1 - mask) * drop_const
        dout[q] = __fadd_rn(__fmul_rn(mk[q], din[q]), __fmul_rn(__fsub_rn(1.0f, mk[q]), p.drop_const));
        // tanh_to_sigmoid + clamp_(0,1); NaN survives the clamp as in torch
        float inv = __fmul_rn(__fadd_rn(dout[q], 1.0f), 0.5f);
        inv = inv < 0.0f ? 0.0f : (inv > 1.0f ? 1.0f : inv);
        bool valid;
        rng[q] = range_from_inv(inv, p, valid);
        vbits[it] |= valid ? (1u << q) : 0u;
      }
      stg_stream(reinterpret_cast<float4*>(odepth + pix), make_float4(dout[0], dout[1], dout[2], dout[3]));
      if (!COMPACT && a.out_points == nullptr) continue;   // maskout alone: no projection, no trig table
      const float4 ce = *reinterpret_cast<const float4*>(a.trig + pix);
      const float4 se = *reinterpret_cast<const float4*>(a.trig + npix + pix);
      const float4 ca = *reinterpret_cast<const float4*>(a.trig + 2 * npix + pix);
      const float4 sa = *reinterpret_cast<const float4*>(a.trig + 3 * npix + pix);
      const float cev[4] = {ce.x, ce.y, ce.z, ce.w}, sev[4] = {se.x, se.y, se.z, se.w};
      const float cav[4] = {ca.x, ca.y, ca.z, ca.w}, sav[4] = {sa.x, sa.y, sa.z, sa.w};
      float X[4], Y[4], Z[4];
      #pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float rc = __fmul_rn(rng[q], cev[q]);
        X[q] = __fmul_rn(rc, cav[q]);
        Y[q] = __fmul_rn(rc, sav[q]);
        Z[q] = __fmul_rn(rng[q], sev[q]);
      }
      xs[it] = make_float4(X[0], X[1], X[2], X[3]);
      ys[it] = make_float4(Y[0], Y[1], Y[2], Y[3]);
      zs[it] = make_float4(Z[0], Z[1], Z[2], Z[3]);
      if (a.out_points != nullptr) {
        if (p.points_layout == 0) {
          float* o = a.out_points + img * 3 * npix + pix;
          stg_stream(reinterpret_cast<float4*>(o), xs[it]);
          stg_stream(reinterpret_cast<float4*>(o + npix), ys[it]);
          stg_stream(reinterpret_cast<float4*>(o + 2 * npix), zs[it]);
        } else {
          const int wpix = seg * SEG + it * (TPB * 4) + warp * 128;       // first pixel of this warp's 128
          if (wpix + 128 <= npix) {
            // whole warp in range (uniform): transpose through shared memory so that every STG.128 of
            // the warp covers 512 contiguous bytes instead of 16 bytes out of every 48
            float4* stage = reinterpret_cast<float4*>(xyz_stage[warp]);
            stage[lane * 3] = make_float4(X[0], Y[0], Z[0], X[1]);
            stage[lane * 3 + 1] = make_float4(Y[1], Z[1], X[2], Y[2]);
            stage[lane * 3 + 2] = make_float4(Z[2], X[3], Y[3], Z[3]);
            __syncwarp();
            float4* o = reinterpret_cast<float4*>(a.out_points + (img * npix + wpix) * 3);
            stg_stream(o + lane, stage[lane]);
            stg_stream(o + lane + 32, stage[lane + 32]);
            stg_stream(o + lane + 64, stage[lane + 64]);
            __syncwarp();
          } else {
            float4* o = reinterpret_cast<float4*>(a.out_points + (img * npix + pix) * 3);
            stg_stream(o, make_float4(X[0], Y[0], Z[0], X[1]));
            stg_stream(o + 1, make_float4(Y[1], Z[1], X[2], Y[2]));
            stg_stream(o + 2, make_float4(Z[2], X[3], Y[3], Z[3]));
          }
        }
      }
      tcount += __popc(vbits[it]);
    }
  }

  if (!COMPACT) return;

  // ---- ordered compaction: positions follow pixel order within the image ----
  // 1. CTA total -> publish; 2. look back over the earlier segments of this image; 3. scatter.
  int total = tcount;
  #pragma unroll
  for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
  if (lane == 0) wsum[0][warp] = total;
  __syncthreads();
  if (tid == 0) {
    int cta_total = 0;
    #pragma unroll
    for (int w = 0; w < TPB / 32; ++w) cta_total += wsum[0][w];
    volatile unsigned* st = a.seg_state + img * a.segs_per_image;
    st[seg] = 0x80000000u | (unsigned)cta_total;
    __threadfence();
    int base = 0;
    for (int s = 0; s < seg; ++s) {
      unsigned v;
      do { v = st[s]; } while (!(v & 0x80000000u));
      base += (int)(v & 0x7fffffffu);
    }
    s_base = base;
    if (seg == a.segs_per_image - 1 && a.out_count) a.out_count[img] = base + cta_total;
  }
  __syncthreads();
  int running = s_base;
  #pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int c = __popc(vbits[it]);
    int inc = c;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) wsum[it & 1][warp] = inc;
    __syncthreads();
    int before = 0, all = 0;
    #pragma unroll
    for (int w = 0; w < TPB / 32; ++w) { const int s = wsum[it & 1][w]; if (w < warp) before += s; all += s; }
    int pos = running + before + inc - c;
    running += all;
    const int pix = seg * SEG + it * (TPB * 4) + tid * 4;
    const float X[4] = {xs[it].x, xs[it].y, xs[it].z, xs[it].w};
    const float Y[4] = {ys[it].x, ys[it].y, ys[it].z, ys[it].w};
    const float Z[4] = {zs[it].x, zs[it].y, zs[it].z, zs[it].w};
    #pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (vbits[it] & (1u << q)) {
        if (a.out_index) a.out_index[img * npix + pos] = pix + q;
        if (a.out_compact) {
          float* o = a.out_compact + (img * npix + pos) * 3;
          o[0] = X[q]; o[1] = Y[q]; o[2] = Z[q];
        }
        ++pos;
      }
    }
  }
}

__global__ void __launch_bounds__(TPB) inv_to_xyz_kernel(const dusty_head_params p, const float* __restrict__ inv_all,
                                                         const float* __restrict__ trig, float* __restrict__ out, int npix) {
  const long long img = blockIdx.y;
  const int pix = (blockIdx.x * TPB + threadIdx.x) * 4;
  if (pix >= npix) return;
  const float4 iv = ldg_stream(reinterpret_cast<const float4*>(inv_all + img * npix + pix));
  const float in[4] = {iv.x, iv.y, iv.z, iv.w};
  const float4 ce = *reinterpret_cast<const float4*>(trig + pix);
  const float4 se = *reinterpret_cast<const float4*>(trig + npix + pix);
  const float4 ca = *reinterpret_cast<const float4*>(trig + 2 * npix + pix);
  const float4 sa = *reinterpret_cast<const float4*>(trig + 3 * npix + pix);
  const float cev[4] = {ce.x, ce.y, ce.z, ce.w}, sev[4] = {se.x, se.y, se.z, se.w};
  const float cav[4] = {ca.x, ca.y, ca.z, ca.w}, sav[4] = {sa.x, sa.y, sa.z, sa.w};
  float X[4], Y[4], Z[4];
  #pragma unroll
  for (int q = 0; q < 4; ++q) {
    bool valid;
    const float r = range_from_inv(in[q], p, valid);
    const float rc = __fmul_rn(r, cev[q]);
    X[q] = __fmul_rn(rc, cav[q]);
    Y[q] = __fmul_rn(rc, sav[q]);
    Z[q] = __fmul_rn(r, sev[q]);
  }
  if (p.points_layout == 0) {
    float* o = out + img * 3 * npix + pix;
    stg_stream(reinterpret_cast<float4*>(o), make_float4(X[0], X[1], X[2], X[3]));
    stg_stream(reinterpret_cast<float4*>(o + npix), make_float4(Y[0], Y[1], Y[2], Y[3]));
    stg_stream(reinterpret_cast<float4*>(o + 2 * npix), make_float4(Z[0], Z[1], Z[2], Z[3]));
  } else {
    float4* o = reinterpret_cast<float4*>(out + (img * npix + pix) * 3);
    stg_stream(o, make_float4(X[0], Y[0], Z[0], X[1]));
    stg_stream(o + 1, make_float4(Y[1], Z[1], X[2], Y[2]));
    stg_stream(o + 2, make_float4(Z[2], X[3], Y[3], Z[3]));
  }
}

__global__ void __launch_bounds__(TPB) logistic_noise_kernel(const float* __restrict__ u1, const float* __restrict__ u2,
                                                             float eps, size_t count, float* __restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * TPB + threadIdx.x; i < count; i += (size_t)gridDim.x * TPB)
    out[i] = logistic_from_uniform(u1[i], u2[i], eps);
}

__global__ void __launch_bounds__(TPB) gumbel_sigmoid_kernel(const float* __restrict__ logits, const GateDev g,
                                                             const dusty_head_params p, int npix, float* __restrict__ out) {
  const long long img = blockIdx.y;
  const int pix = (blockIdx.x * TPB + threadIdx.x) * 4;
  if (pix >= npix) return;
  const float4 cv = ldg_stream(reinterpret_cast<const float4*>(logits + img * npix + pix));
  float4 na = make_float4(0, 0, 0, 0), nb = na;
  if (g.mode != DUSTY_NOISE_NONE) na = load_noise(g, g.a, img, pix);
  if (g.mode == DUSTY_NOISE_UNIFORM) nb = load_noise(g, g.b, img, pix);
  stg_stream(reinterpret_cast<float4*>(out + img * npix + pix),
             make_float4(gate_value(cv.x, g.mode, na.x, nb.x, p), gate_value(cv.y, g.mode, na.y, nb.y, p),
                         gate_value(cv.z, g.mode, na.z, nb.z, p), gate_value(cv.w, g.mode, na.w, nb.w, p)));
}

// Pixel groups per thread. Measured on B200 (tests/perf_head_iters.py, batch 256 of 64x512, DUSty-I):
// ITERS = 4 (62 registers, 4 CTAs/SM, 2048 CTAs = 3.46 waves) 43.3 us; ITERS = 2 (42 registers) 39.1 us;
// ITERS = 1 (32 registers, 8 CTAs/SM = 2048 threads/SM, 8192 CTAs) 37.4 us = 6.29 TB/s. One group per
// thread wins at every batch size: full occupancy hides the latency that the extra loads per thread
// were meant to hide, and the last wave is small. DUSTY_HEAD_ITERS=2|4 re-creates the A/B.
static int pick_iters(int, int) {
  if (const char* e = getenv("DUSTY_HEAD_ITERS")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4) return v; }
  return 1;
}

static int check_params(const dusty_head_params* p, const char* who) {
  if (!p) return fail_arg(DUSTY_EINVAL, "%s: null params", who);
  if (p->b < 0 || p->h <= 0 || p->w <= 0) return fail_arg(DUSTY_EINVAL, "%s: bad shape b=%d h=%d w=%d", who, p->b, p->h, p->w);
  if (p->w % 4 != 0) return fail_arg(DUSTY_EINVAL, "%s: w=%d must be a multiple of 4", who, p->w);
  if ((long long)p->h * p->w > 0x7fffffffLL / 4) return fail_arg(DUSTY_EINVAL, "%s: image too large", who);
  if (p->points_layout != 0 && p->points_layout != 1) return fail_arg(DUSTY_EINVAL, "%s: points_layout must be 0 or 1", who);
  return 0;
}

static int check_gate(const dusty_gate& g, const char* name) {
  if (g.mode < DUSTY_NOISE_NONE || g.mode > DUSTY_NOISE_UNIFORM) return fail_arg(DUSTY_EINVAL, "head_project: %s gate mode %d", name, g.mode);
  if (g.mode != DUSTY_NOISE_NONE && !g.noise_a) return fail_arg(DUSTY_EINVAL, "head_project: %s gate needs noise_a", name);
  if (g.mode == DUSTY_NOISE_UNIFORM && !g.noise_b) return fail_arg(DUSTY_EINVAL, "head_project: %s gate needs noise_b", name);
  if (g.pixel_stride != 0 && g.pixel_stride != 1) return fail_arg(DUSTY_EINVAL, "head_project: %s gate pixel_stride must be 0 or 1", name);
  if (g.mode != DUSTY_NOISE_NONE && g.pixel_stride == 1) {
    if (!aligned16(g.noise_a) || (g.noise_b && !aligned16(g.noise_b)) || (g.batch_stride % 4) != 0)
      return fail_arg(DUSTY_EALIGN, "head_project: %s gate noise must be 16-byte aligned with batch_stride %% 4 == 0", name);
  }
  return 0;
}

}  // namespace head
}  // namespace dusty

using namespace dusty;
using namespace dusty::head;

extern "C" size_t dusty_head_project_workspace_bytes(int b, int h, int w) {
  if (b <= 0 || h <= 0 || w <= 0) return 0;
  const long long npix = (long long)h * w;
  const long long segs = (npix + GROUP - 1) / GROUP;      // the finest segmentation any launch uses
  return align_up((size_t)b * segs * sizeof(unsigned), 256);
}

extern "C" int dusty_head_project(const dusty_head_params* p, const float* depth, const float* confidence,
                                  const float* trig, float* out_mask, float* out_depth, float* out_points,
                                  int32_t* out_count, int32_t* out_index, float* out_compact, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (int rc = check_params(p, "head_project")) return rc;
  if (p->conf_channels != 1 && p->conf_channels != 2) return fail_arg(DUSTY_EINVAL, "head_project: conf_channels must be 1 or 2");
  if (p->b == 0) return 0;
  if (int rc = check_device()) return rc;
  const bool compact = out_count || out_index || out_compact;
  if (!depth || !confidence || !out_mask || !out_depth || (!trig && (out_points || compact)))
    return fail_arg(DUSTY_EINVAL, "head_project: null pointer");
  if (!aligned16(depth) || !aligned16(confidence) || (trig && !aligned16(trig)) || !aligned16(out_mask) || !aligned16(out_depth) ||
      (out_points && !aligned16(out_points)))
    return fail_arg(DUSTY_EALIGN, "head_project: tensors must be 16-byte aligned");
  if (int rc = check_gate(p->gate_pixel, "pixel")) return rc;
  if (p->conf_channels == 2) if (int rc = check_gate(p->gate_image, "image")) return rc;
  Args a{};
  a.p = *p;
  a.gp = GateDev{p->gate_pixel.mode, p->gate_pixel.noise_a, p->gate_pixel.noise_b, p->gate_pixel.batch_stride, (int)p->gate_pixel.pixel_stride};
  a.gi = GateDev{p->gate_image.mode, p->gate_image.noise_a, p->gate_image.noise_b, p->gate_image.batch_stride, (int)p->gate_image.pixel_stride};
  if (p->conf_channels == 1) a.gi.mode = DUSTY_NOISE_NONE;
  a.depth = depth; a.conf = confidence; a.trig = trig;
  a.out_mask = out_mask; a.out_depth = out_depth; a.out_points = out_points;
  a.out_count = out_count; a.out_index = out_index; a.out_compact = out_compact;
  a.npix = p->h * p->w;
  const int iters = pick_iters(p->b, a.npix);
  const int seg = GROUP * iters;
  a.segs_per_image = (a.npix + seg - 1) / seg;
  const long long ctas = (long long)p->b * a.segs_per_image;
  if (ctas > 0x7fffffffLL) return fail_arg(DUSTY_EINVAL, "head_project: grid too large");
  if (compact) {
    if (!workspace || workspace_bytes < dusty_head_project_workspace_bytes(p->b, p->h, p->w))
      return fail_arg(DUSTY_ENOSPACE, "head_project: compaction needs %zu workspace bytes", dusty_head_project_workspace_bytes(p->b, p->h, p->w));
    a.seg_state = static_cast<unsigned*>(workspace);
    DUSTY_CUDA(cudaMemsetAsync(workspace, 0, (size_t)ctas * sizeof(unsigned), st));
  }
#define DUSTY_HEAD_LAUNCH(C, COMPACT)                                                                     \
  switch (iters) {                                                                                        \
    case 1: head_project_kernel<C, COMPACT, 1><<<(unsigned)ctas, TPB, 0, st>>>(a); break;                 \
    case 2: head_project_kernel<C, COMPACT, 2><<<(unsigned)ctas, TPB, 0, st>>>(a); break;                 \
    default: head_project_kernel<C, COMPACT, 4><<<(unsigned)ctas, TPB, 0, st>>>(a); break;                \
  }
  if (p->conf_channels == 1) {
    if (compact) { DUSTY_HEAD_LAUNCH(1, true) } else { DUSTY_HEAD_LAUNCH(1, false) }
  } else {
    if (compact) { DUSTY_HEAD_LAUNCH(2, true) } else { DUSTY_HEAD_LAUNCH(2, false) }
  }
#undef DUSTY_HEAD_LAUNCH
  DUSTY_AFTER_LAUNCH("head_project_kernel");
  return 0;
}

extern "C" int dusty_gumbel_sigmoid(const float* logits, const dusty_gate* gate, float inv_tau, float threshold, float eps,
                                    int b, int npix, float* out, void* stream) {
  if (b < 0 || npix <= 0 || npix % 4 != 0) return fail_arg(DUSTY_EINVAL, "gumbel_sigmoid: bad shape b=%d npix=%d", b, npix);
  if (b == 0) return 0;
  if (b > 65535) return fail_arg(DUSTY_EINVAL, "gumbel_sigmoid: batch %d exceeds 65535", b);
  if (int rc = check_device()) return rc;
  if (!logits || !gate || !out) return fail_arg(DUSTY_EINVAL, "gumbel_sigmoid: null pointer");
  if (!aligned16(logits) || !aligned16(out)) return fail_arg(DUSTY_EALIGN, "gumbel_sigmoid: tensors must be 16-byte aligned");
  if (int rc = check_gate(*gate, "gumbel")) return rc;
  dusty_head_params p{};
  p.inv_tau = inv_tau; p.threshold = threshold; p.eps = eps;
  const GateDev g{gate->mode, gate->noise_a, gate->noise_b, gate->batch_stride, (int)gate->pixel_stride};
  gumbel_sigmoid_kernel<<<dim3((npix / 4 + TPB - 1) / TPB, b), TPB, 0, static_cast<cudaStream_t>(stream)>>>(logits, g, p, npix, out);
  DUSTY_AFTER_LAUNCH("gumbel_sigmoid_kernel");
  return 0;
}

extern "C" int dusty_inv_to_xyz(const dusty_head_params* p, const float* inv, const float* trig, float* out_points,
                                void* stream) {
  if (int rc = check_params(p, "inv_to_xyz")) return rc;
  if (p->b == 0) return 0;
  if (p->b > 65535) return fail_arg(DUSTY_EINVAL, "inv_to_xyz: batch %d exceeds 65535", p->b);
  if (int rc = check_device()) return rc;
  if (!inv || !trig || !out_points) return fail_arg(DUSTY_EINVAL, "inv_to_xyz: null pointer");
  if (!aligned16(inv) || !aligned16(trig) || !aligned16(out_points)) return fail_arg(DUSTY_EALIGN, "inv_to_xyz: tensors must be 16-byte aligned");
  const int npix = p->h * p->w;
  inv_to_xyz_kernel<<<dim3((npix / 4 + TPB - 1) / TPB, p->b), TPB, 0, static_cast<cudaStream_t>(stream)>>>(*p, inv, trig, out_points, npix);
  DUSTY_AFTER_LAUNCH("inv_to_xyz_kernel");
  return 0;
}

extern "C" int dusty_logistic_noise(const float* u1, const float* u2, float eps, size_t count, float* out, void* stream) {
  if (count == 0) return 0;
  if (int rc = check_device()) return rc;
  if (!u1 || !u2 || !out) return fail_arg(DUSTY_EINVAL, "logistic_noise: null pointer");
  const unsigned grid = (unsigned)std::min<size_t>((count + TPB - 1) / TPB, (size_t)kNumSMs * 16);
  logistic_noise_kernel<<<grid, TPB, 0, static_cast<cudaStream_t>(stream)>>>(u1, u2, eps, count, out);
  DUSTY_AFTER_LAUNCH("logistic_noise_kernel");
  return 0;
}
