// samplers.cu -- the texture samplers either side of `interpolate` in real DRTK pipelines:
//   mipmap_grid_sample  (src/mipmap_grid_sampler/mipmap_grid_sampler_kernel.cu:20-898): trilinear + anisotropic
//                       lookup of a mip pyramid at per-pixel uv, footprint from the uv Jacobian
//   grid_scatter        (src/grid_scatter/grid_scatter_kernel.cu:18-519): the splatting transpose of grid_sample
// forward and backward, bilinear / bicubic, zeros / border / reflection padding.
//
// Coordinate conventions are those of torch.nn.functional.grid_sample (ATen/native/cuda/GridSampler.cuh), which
// the reference pulls in through src/include/grid_utils.h:7-24; they are restated here (this library has no torch
// headers).  Quirks of the reference that a drop-in has to keep are marked "reference quirk".
//
// Work mapping: one thread per grid pixel (w fastest), so grid / Jacobian / output-plane accesses are coalesced;
// the forward keeps four channels of the pixel in registers over all taps of all samples and stores each output
// once (the reference does a global read-modify-write per tap and channel, :69-80); gradients into textures go
// out as fire-and-forget `red.global.add.f32`.
#include <math_constants.h>

#include "common.cuh"

namespace drtk {
namespace {

constexpr int kMaxLevels = 11;  // max_mipmap_count, mipmap_grid_sampler_kernel.cu:16
constexpr int kZeros = 0, kBorder = 1, kReflection = 2;

// I = int32_t when every offset into every level fits 31 bits (the common case; like the reference's
// canUse32BitIndexMath dispatch), else int64_t: 64-bit multiplies cost four integer instructions each on the SM
template <typename I>
struct Tex {
  const float* p;
  float* g;  // gradient accumulator (dense NCHW) or nullptr
  int H, W;
  I sN, sC, sH, sW;
};
template <typename I>
struct TexList { Tex<I> t[kMaxLevels]; };

// ---- grid_sample coordinate helpers -------------------------------------------------------
__device__ __forceinline__ float unnormalize(float c, int size, bool align, float* mult) {
  if (align) { *mult = float(size - 1) / 2.f; return ((c + 1.f) / 2.f) * float(size - 1); }
  *mult = float(size) / 2.f;
  return ((c + 1.f) * float(size) - 1.f) / 2.f;
}
__device__ __forceinline__ float clip_coord(float c, int size, float* mult) {
  if (c <= 0.f) { *mult = 0.f; return 0.f; }
  const float mx = float(size - 1);
  if (c >= mx) { *mult = 0.f; return mx; }
  *mult = 1.f;
  return c;
}
__device__ __forceinline__ float clip_plain(float c, int size) { return fminf(float(size - 1), fmaxf(c, 0.f)); }
__device__ __forceinline__ float reflect_coord(float c, int twice_low, int twice_high, float* mult) {
  if (twice_low == twice_high) { *mult = 0.f; return 0.f; }
  const float mn = float(twice_low) / 2.f, span = float(twice_high - twice_low) / 2.f;
  c -= mn;
  float sgn = 1.f;
  if (c < 0.f) { sgn = -1.f; c = -c; }
  const float extra = fmodf(c, span);
  const int flips = int(floorf(c / span));
  if ((flips & 1) == 0) { *mult = sgn; return extra + mn; }
  *mult = -sgn;
  return span - extra + mn;
}
__device__ __forceinline__ float to_int_range(float x) {
  return (x > float(INT_MAX - 1) || x < float(INT_MIN) || !isfinite(x)) ? -100.f : x;
}
// compute_coordinates (grid_utils.h:84-105): padding applied to an already unnormalised coordinate
__device__ __forceinline__ float pad_coord(float c, int size, int pad, bool align, float* mult) {
  float m = 1.f;
  if (pad == kBorder) {
    c = clip_coord(c, size, &m);
  } else if (pad == kReflection) {
    float mr, mc;
    c = align ? reflect_coord(c, 0, 2 * (size - 1), &mr) : reflect_coord(c, -1, 2 * size - 1, &mr);
    c = clip_coord(c, size, &mc);
    m = mr * mc;
  }
  *mult = m;
  return to_int_range(c);
}
__device__ __forceinline__ int pad_tap(float c, int size, int pad, bool align) {
  // the value version clips with min/max (GridSampler.cuh clip_coordinates), same result as clip_coord
  if (pad == kBorder) c = clip_plain(c, size);
  else if (pad == kReflection) {
    float m;
    c = align ? reflect_coord(c, 0, 2 * (size - 1), &m) : reflect_coord(c, -1, 2 * size - 1, &m);
    c = clip_plain(c, size);
  }
  return int(to_int_range(c));
}
// grid_sampler_compute_source_index[_set_grad]
__device__ __forceinline__ float source_index(float c, int size, int pad, bool align, float* mult) {
  float m0, m1;
  c = unnormalize(c, size, align, &m0);
  c = pad_coord(c, size, pad, align, &m1);
  *mult = m0 * m1;
  return c;
}

// cubic convolution, A = -0.75 (ATen UpSample.cuh get_cubic_upsampling_coefficients; derivative grid_utils.h:131-146)
__device__ __forceinline__ void cubic_coeffs(float t, float (&c)[4]) {
  const float A = -0.75f;
  float x = t + 1.f;
  c[0] = ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A;
  x = t;
  c[1] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
  x = 1.f - t;
  c[2] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
  x = 2.f - t;
  c[3] = ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A;
}
__device__ __forceinline__ void cubic_coeffs_grad(float t, float (&c)[4]) {
  const float A = -0.75f;
  float x = -1.f - t;
  c[0] = (-3.f * A * x - 10.f * A) * x - 8.f * A;
  x = -t;
  c[1] = (-3.f * (A + 2.f) * x - 2.f * (A + 3.f)) * x;
  x = 1.f - t;
  c[2] = (3.f * (A + 2.f) * x - 2.f * (A + 3.f)) * x;
  x = 2.f - t;
  c[3] = (3.f * A * x - 10.f * A) * x + 8.f * A;
}

// ---- footprints ---------------------------------------------------------------------------
// Bilinear: 4 taps.  Bicubic: 4x4 taps with separable weights.  `off` < 0 marks a tap outside the image.
struct Bilinear {
  int x0, y0;       // north-west tap; taps q = 0..3 are nw ne sw se = (x0 + (q & 1), y0 + (q >> 1))
  unsigned inside;  // bit q set when tap q lies in the image
  float w[4];       // area weights
  float dx[4], dy[4];  // d w / d ix, d w / d iy
  float mx, my;     // d ix / d u, d iy / d v
  __device__ __forceinline__ Bilinear(float u, float v, int H, int W, int pad, bool align) {
    const float ix = source_index(u, W, pad, align, &mx), iy = source_index(v, H, pad, align, &my);
    x0 = int(floorf(ix)); y0 = int(floorf(iy));
    const int x1 = x0 + 1, y1 = y0 + 1;
    const float fx1 = float(x1) - ix, fx0 = ix - float(x0), fy1 = float(y1) - iy, fy0 = iy - float(y0);
    w[0] = fx1 * fy1; w[1] = fx0 * fy1; w[2] = fx1 * fy0; w[3] = fx0 * fy0;
    dx[0] = -fy1; dx[1] = fy1; dx[2] = -fy0; dx[3] = fy0;
    dy[0] = -fx1; dy[1] = -fx0; dy[2] = fx1; dy[3] = fx0;
    const bool bx0 = x0 >= 0 && x0 < W, bx1 = x1 >= 0 && x1 < W, by0 = y0 >= 0 && y0 < H, by1 = y1 >= 0 && y1 < H;
    inside = unsigned(bx0 && by0) | unsigned(bx1 && by0) << 1 | unsigned(bx0 && by1) << 2 | unsigned(bx1 && by1) << 3;
  }
  __device__ __forceinline__ bool in(int q) const { return (inside >> q) & 1u; }
  template <typename I>
  __device__ __forceinline__ I off(int q, I sH, I sW) const {
    return I(y0 + (q >> 1)) * sH + I(x0 + (q & 1)) * sW;
  }
};

struct Bicubic {
  int xi[4], yi[4];  // padded tap columns / rows, -1 when outside the image
  float cx[4], cy[4], gx[4], gy[4];
  float mx, my;
  // pad_centre: grid_scatter pads the sample position itself (grid_scatter_kernel.cu:141-142); the mipmap sampler
  // only unnormalises it (mipmap_grid_sampler_kernel.cu:108-109).  Every tap is padded individually in both.
  __device__ __forceinline__ Bicubic(float u, float v, int H, int W, int pad, bool align, bool pad_centre,
                                     bool want_grad) {
    float ix, iy;
    if (pad_centre) { ix = source_index(u, W, pad, align, &mx); iy = source_index(v, H, pad, align, &my); }
    else { ix = unnormalize(u, W, align, &mx); iy = unnormalize(v, H, align, &my); }
    const float x0 = floorf(ix), y0 = floorf(iy);
    cubic_coeffs(ix - x0, cx); cubic_coeffs(iy - y0, cy);
    if (want_grad) { cubic_coeffs_grad(ix - x0, gx); cubic_coeffs_grad(iy - y0, gy); }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int x = pad_tap(x0 - 1.f + float(i), W, pad, align), y = pad_tap(y0 - 1.f + float(i), H, pad, align);
      xi[i] = (x >= 0 && x < W) ? x : -1;
      yi[i] = (y >= 0 && y < H) ? y : -1;
    }
  }
  __device__ __forceinline__ bool in(int q, int j) const { return (xi[q] | yi[j]) >= 0; }
  template <typename I>
  __device__ __forceinline__ I off(int q, int j, I sH, I sW) const { return I(yi[j]) * sH + I(xi[q]) * sW; }
};

// ---- mip level selection (mipmap_grid_sampler_kernel.cu:441-498) ---------------------------
struct Footprint {
  float du, dv;  // uv step of the major axis
  int d1, n;     // lower level, number of samples
  float a;       // blend towards level d1 + 1
};
__device__ __forceinline__ Footprint select_levels(float dudx, float dvdx, float dudy, float dvdy, int W0, int H0,
                                                   int levels, float lmax, int max_aniso, bool force_max,
                                                   bool clip_grad) {
  const float ax = dudx * float(W0), bx = dvdx * float(H0), ay = dudy * float(W0), by = dvdy * float(H0);
  const float px = sqrtf(ax * ax + bx * bx + 1e-12f), py = sqrtf(ay * ay + by * by + 1e-12f);
  const float pmax = fmaxf(px, py), pmin = fminf(px, py);
  float N = fminf(ceilf(pmax / pmin), float(max_aniso));
  if (pmin == 0.f || N == 0.f) N = 1.f;
  float lambda = log2f(pmax / N);
  if (isnan(lambda) || isinf(lambda)) lambda = 0.f;
  // lmax = float(double(levels - 1) - 1e-6), rounded on the host the way the reference's mixed float/double `min` does
  float l = fminf(lambda, lmax);
  if (clip_grad && lambda > float(levels - 1)) {  // missing coarse levels: shrink the footprint instead
    const float s = exp2f(l) * N / pmax;
    dudx *= s; dvdx *= s; dudy *= s; dvdy *= s;
  }
  l = fmaxf(l, 0.f);
  Footprint f;
  f.d1 = int(floorf(l));
  f.a = l - float(f.d1);
  f.n = force_max ? max_aniso : int(N);
  const bool major_x = px > py;
  f.du = major_x ? dudx : dudy;
  f.dv = major_x ? dvdx : dvdy;
  return f;
}
// position of sample i of n along the major axis, (i+1)/(n+1)*2-1 = (2i+1-n)/(n+1).  The reference evaluates it in
// double (:500-501); one float division instead keeps fp64 instructions (1/64 rate on B200) out of the tap loop
__device__ __forceinline__ float sample_pos(int i, int n) { return float(2 * i + 1 - n) / float(n + 1); }

struct PixelArgs {
  const float* grid; Strides4 gs;
  const float* jac; int64_t js[5];
  int N, C, H, W;
  int levels, max_aniso, pad;
  float lmax;
  bool align, force_max, clip_grad;
};

constexpr int kCh = 4;  // channels kept in registers per pass

template <bool BICUBIC, typename I>
__global__ void __launch_bounds__(256)
mipmap_fwd_kernel(TexList<I> tex, PixelArgs a, float* __restrict__ out) {
  const int64_t total = int64_t(a.N) * a.H * a.W, plane = int64_t(a.H) * a.W;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total; idx += int64_t(gridDim.x) * blockDim.x) {
    const int w = int(idx % a.W), h = int((idx / a.W) % a.H), n = int(idx / plane);
    const float* pg = a.grid + n * a.gs.s0 + h * a.gs.s1 + w * a.gs.s2;
    const float u = __ldg(pg), v = __ldg(pg + a.gs.s3);
    const float* pj = a.jac + n * a.js[0] + h * a.js[1] + w * a.js[2];
    const Footprint f = select_levels(__ldg(pj), __ldg(pj + a.js[4]), __ldg(pj + a.js[3]), __ldg(pj + a.js[3] + a.js[4]),
                                      tex.t[0].W, tex.t[0].H, a.levels, a.lmax, a.max_aniso, a.force_max, a.clip_grad);
    const int nlev = a.levels > 1 ? 2 : 1;
    const float inv_n = 1.f / float(f.n);
    float* po = out + int64_t(n) * a.C * plane + int64_t(h) * a.W + w;
    for (int c0 = 0; c0 < a.C; c0 += kCh) {
      float acc[kCh] = {0.f, 0.f, 0.f, 0.f};
      for (int i = 0; i < f.n; ++i) {
        const float t = sample_pos(i, f.n), su = u + f.du * t, sv = v + f.dv * t;
        for (int lv = 0; lv < nlev; ++lv) {
          const Tex<I>& T = tex.t[f.d1 + lv];
          const float alpha = (lv == 0 ? 1.f - f.a : f.a) * inv_n;
          // a magnified texture (footprint below one texel) has a == 0: the coarser level would be fetched only to
          // be multiplied by zero, as the reference does (:505-528); skipped here (identical for finite texels)
          if (alpha == 0.f) continue;
          const float* base = T.p + I(n) * T.sN + I(c0) * T.sC;
          // reference quirk: the forward kernel overrides align_corners with false (:423)
          if (!BICUBIC) {
            const Bilinear b(su, sv, T.H, T.W, a.pad, false);
#pragma unroll
            for (int k = 0; k < kCh; ++k) {
              if (c0 + k < a.C) {
                float s = 0.f;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  if (b.in(q)) s += __ldg(base + I(k) * T.sC + b.off(q, T.sH, T.sW)) * b.w[q];
                acc[k] += s * alpha;
              }
            }
          } else {
            const Bicubic b(su, sv, T.H, T.W, a.pad, false, false, false);
#pragma unroll
            for (int k = 0; k < kCh; ++k) {
              if (c0 + k < a.C) {
                float s = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  float r = 0.f;
#pragma unroll
                  for (int q = 0; q < 4; ++q)
                    if (b.in(q, j)) r += __ldg(base + I(k) * T.sC + b.off(q, j, T.sH, T.sW)) * b.cx[q];
                  s += r * b.cy[j];
                }
                acc[k] += s * alpha;
              }
            }
          }
        }
      }
#pragma unroll
      for (int k = 0; k < kCh; ++k)
        if (c0 + k < a.C) po[int64_t(c0 + k) * plane] = acc[k];
    }
  }
}

template <bool BICUBIC, typename I>
__global__ void __launch_bounds__(256)
mipmap_bwd_kernel(TexList<I> tex, PixelArgs a, const float* __restrict__ gout, Strides4 gos, float* __restrict__ grad_grid) {
  const int64_t total = int64_t(a.N) * a.H * a.W, plane = int64_t(a.H) * a.W;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total; idx += int64_t(gridDim.x) * blockDim.x) {
    const int w = int(idx % a.W), h = int((idx / a.W) % a.H), n = int(idx / plane);
    const float* pg = a.grid + n * a.gs.s0 + h * a.gs.s1 + w * a.gs.s2;
    const float u = __ldg(pg), v = __ldg(pg + a.gs.s3);
    const float* pj = a.jac + n * a.js[0] + h * a.js[1] + w * a.js[2];
    const Footprint f = select_levels(__ldg(pj), __ldg(pj + a.js[4]), __ldg(pj + a.js[3]), __ldg(pj + a.js[3] + a.js[4]),
                                      tex.t[0].W, tex.t[0].H, a.levels, a.lmax, a.max_aniso, a.force_max, a.clip_grad);
    const int nlev = a.levels > 1 ? 2 : 1;
    const float inv_n = 1.f / float(f.n);
    const float* pgo = gout + n * gos.s0 + h * gos.s2 + w * gos.s3;
    float gu = 0.f, gv = 0.f;
    for (int i = 0; i < f.n; ++i) {
      const float t = sample_pos(i, f.n), su = u + f.du * t, sv = v + f.dv * t;
      for (int lv = 0; lv < nlev; ++lv) {
        const Tex<I>& T = tex.t[f.d1 + lv];
        const float alpha = (lv == 0 ? 1.f - f.a : f.a) * inv_n;
        if (alpha == 0.f) continue;  // contributes exact zeros to every gradient
        const float* base = T.p + I(n) * T.sN;
        const I tplane = I(T.H) * I(T.W);
        float* gbase = T.g ? T.g + I(n) * I(a.C) * tplane : nullptr;
        float sx = 0.f, sy = 0.f;
        // the backward kernel honours align_corners (reference quirk: unlike its forward)
        if (!BICUBIC) {
          const Bilinear b(su, sv, T.H, T.W, a.pad, a.align);
          for (int c = 0; c < a.C; ++c) {
            const float go = __ldg(pgo + c * gos.s1) * alpha;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (b.in(q)) {
                if (gbase) red_add(gbase + I(c) * tplane + b.off(q, I(T.W), I(1)), b.w[q] * go);
                const float val = __ldg(base + I(c) * T.sC + b.off(q, T.sH, T.sW));
                sx += val * b.dx[q] * go;
                sy += val * b.dy[q] * go;
              }
            }
          }
          gu += b.mx * sx; gv += b.my * sy;
        } else {
          const Bicubic b(su, sv, T.H, T.W, a.pad, a.align, false, true);
          for (int c = 0; c < a.C; ++c) {
            const float go = __ldg(pgo + c * gos.s1) * alpha;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                if (b.in(q, j)) {
                  if (gbase) red_add(gbase + I(c) * tplane + b.off(q, j, I(T.W), I(1)), go * b.cx[q] * b.cy[j]);
                  const float val = __ldg(base + I(c) * T.sC + b.off(q, j, T.sH, T.sW));
                  sx -= go * val * b.gx[q] * b.cy[j];
                  sy -= go * val * b.gy[j] * b.cx[q];
                }
              }
            }
          }
          gu += b.mx * sx; gv += b.my * sy;
        }
      }
    }
    if (grad_grid) {
      reinterpret_cast<float2*>(grad_grid)[idx] = make_float2(gu, gv);
    }
  }
}

// ---- grid_scatter ---------------------------------------------------------------------------
struct ScatterArgs {
  const float* input; Strides4 is;
  const float* grid; Strides4 gs;
  int N, C, H, W, Ho, Wo, pad;
  bool align;
};

template <bool BICUBIC>
__global__ void __launch_bounds__(256)
scatter_fwd_kernel(ScatterArgs a, float* __restrict__ out) {
  const int64_t total = int64_t(a.N) * a.H * a.W, plane = int64_t(a.H) * a.W, oplane = int64_t(a.Ho) * a.Wo;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total; idx += int64_t(gridDim.x) * blockDim.x) {
    const int w = int(idx % a.W), h = int((idx / a.W) % a.H), n = int(idx / plane);
    const float* pg = a.grid + n * a.gs.s0 + h * a.gs.s1 + w * a.gs.s2;
    const float u = __ldg(pg), v = __ldg(pg + a.gs.s3);
    const float* pin = a.input + n * a.is.s0 + h * a.is.s2 + w * a.is.s3;
    float* ob = out + int64_t(n) * a.C * oplane;
    if (!BICUBIC) {
      const Bilinear b(u, v, a.Ho, a.Wo, a.pad, a.align);
      for (int c = 0; c < a.C; ++c) {
        const float val = __ldg(pin + c * a.is.s1);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (b.in(q)) red_add(ob + c * oplane + b.off(q, int64_t(a.Wo), int64_t(1)), b.w[q] * val);
      }
    } else {
      const Bicubic b(u, v, a.Ho, a.Wo, a.pad, a.align, true, false);
      for (int c = 0; c < a.C; ++c) {
        const float val = __ldg(pin + c * a.is.s1);
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (b.in(q, j)) red_add(ob + c * oplane + b.off(q, j, int64_t(a.Wo), int64_t(1)), val * b.cx[q] * b.cy[j]);
      }
    }
  }
}

template <bool BICUBIC>
__global__ void __launch_bounds__(256)
scatter_bwd_kernel(ScatterArgs a, const float* __restrict__ gout, Strides4 gos, float* __restrict__ grad_input,
                   float* __restrict__ grad_grid) {
  const int64_t total = int64_t(a.N) * a.H * a.W, plane = int64_t(a.H) * a.W;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total; idx += int64_t(gridDim.x) * blockDim.x) {
    const int w = int(idx % a.W), h = int((idx / a.W) % a.H), n = int(idx / plane);
    const float* pg = a.grid + n * a.gs.s0 + h * a.gs.s1 + w * a.gs.s2;
    const float u = __ldg(pg), v = __ldg(pg + a.gs.s3);
    const float* pin = a.input + n * a.is.s0 + h * a.is.s2 + w * a.is.s3;
    const float* gb = gout + n * gos.s0;
    float* pgi = grad_input ? grad_input + int64_t(n) * a.C * plane + int64_t(h) * a.W + w : nullptr;
    float sx = 0.f, sy = 0.f, mx, my;
    if (!BICUBIC) {
      const Bilinear b(u, v, a.Ho, a.Wo, a.pad, a.align);
      mx = b.mx; my = b.my;
      for (int c = 0; c < a.C; ++c) {
        const float val = grad_grid ? __ldg(pin + c * a.is.s1) : 0.f;
        float gi = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (b.in(q)) {
            const float go = __ldg(gb + c * gos.s1 + b.off(q, gos.s2, gos.s3));
            gi += go * b.w[q];
            sx += val * b.dx[q] * go;
            sy += val * b.dy[q] * go;
          }
        }
        if (pgi) pgi[int64_t(c) * plane] = gi;
      }
    } else {
      const Bicubic b(u, v, a.Ho, a.Wo, a.pad, a.align, true, true);
      mx = b.mx; my = b.my;
      for (int c = 0; c < a.C; ++c) {
        const float val = grad_grid ? __ldg(pin + c * a.is.s1) : 0.f;
        float gi = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float r = 0.f;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (b.in(q, j)) {
              const float go = __ldg(gb + c * gos.s1 + b.off(q, j, gos.s2, gos.s3));
              r += go * b.cx[q];
              sx -= go * val * b.gx[q] * b.cy[j];
              sy -= go * val * b.gy[j] * b.cx[q];
            }
          }
          gi += r * b.cy[j];
        }
        if (pgi) pgi[int64_t(c) * plane] = gi;
      }
    }
    if (grad_grid) reinterpret_cast<float2*>(grad_grid)[idx] = make_float2(mx * sx, my * sy);
  }
}

inline unsigned blocks_for(int64_t total) {
  const int64_t need = (total + 255) / 256, cap = int64_t(num_sms()) * 16;
  return unsigned(need < 1 ? 1 : (need > cap ? cap : need));
}
inline bool bad_enum(int pad, int interp) { return pad < 0 || pad > 2 || (interp != 0 && interp != 2); }

}  // namespace
}  // namespace drtk

using namespace drtk;

// true when every element offset of every level (input strides and the dense gradient planes) fits 31 bits
static bool levels_fit_int32(const int64_t* hw, const int64_t* strides, int L, int64_t N, int64_t C) {
  for (int i = 0; i < L; ++i) {
    const int64_t H = hw[2 * i], W = hw[2 * i + 1], *s = strides + 4 * i;
    int64_t span = 0;
    const int64_t ext[4] = {N, C, H, W};
    for (int d = 0; d < 4; ++d) {
      if (s[d] < 0) return false;
      span += (ext[d] > 0 ? ext[d] - 1 : 0) * s[d];
    }
    if (span >= INT32_MAX || N * C * H * W >= INT32_MAX) return false;
  }
  return true;
}

template <typename I>
static int fill_levels(TexList<I>& tl, const float* const* levels, float* const* grad_levels, const int64_t* hw,
                       const int64_t* strides, int L) {
  for (int i = 0; i < L; ++i) {
    if (hw[2 * i] <= 0 || hw[2 * i + 1] <= 0 || hw[2 * i] > INT32_MAX || hw[2 * i + 1] > INT32_MAX)
      return DRTK_B200_EINVAL;
    Tex<I>& t = tl.t[i];
    t.p = levels[i];
    t.g = grad_levels ? grad_levels[i] : nullptr;
    t.H = int(hw[2 * i]); t.W = int(hw[2 * i + 1]);
    t.sN = I(strides[4 * i]); t.sC = I(strides[4 * i + 1]); t.sH = I(strides[4 * i + 2]); t.sW = I(strides[4 * i + 3]);
  }
  for (int i = L; i < kMaxLevels; ++i) tl.t[i] = tl.t[L - 1];
  return 0;
}

static int fill_pixel_args(PixelArgs& a, const float* grid, const int64_t* gs, const float* jac, const int64_t* js,
                           int64_t N, int64_t C, int64_t H, int64_t W, int L, int max_aniso, int pad, int align,
                           int force_max, int clip_grad) {
  if (!gs || !js || N < 0 || C < 0 || H < 0 || W < 0 || max_aniso < 1) return DRTK_B200_EINVAL;
  if (N * H * W > 0 && (!grid || !jac)) return DRTK_B200_EINVAL;
  if (N > INT32_MAX || C > INT32_MAX || H > INT32_MAX || W > INT32_MAX) return DRTK_B200_EUNSUPPORTED;
  a.grid = grid; a.gs = make4(gs); a.jac = jac;
  for (int i = 0; i < 5; ++i) a.js[i] = js[i];
  a.N = int(N); a.C = int(C); a.H = int(H); a.W = int(W);
  a.levels = L; a.max_aniso = max_aniso; a.pad = pad;
  a.lmax = float(double(L - 1) - 1e-6);
  a.align = align != 0; a.force_max = force_max != 0; a.clip_grad = clip_grad != 0;
  return 0;
}

extern "C" int drtk_b200_mipmap_grid_sample_forward(
    const float* const* levels, const int64_t* level_hw, const int64_t* level_strides, int num_levels,
    const float* grid, const int64_t* grid_strides, const float* vt_dxdy_img, const int64_t* vt_strides, int64_t N,
    int64_t C, int64_t H, int64_t W, int max_aniso, int padding_mode, int interpolation_mode, int align_corners,
    int force_max_aniso, int clip_grad, float* out, void* stream) {
  if (!levels || !level_hw || !level_strides || num_levels < 1 || num_levels > kMaxLevels ||
      bad_enum(padding_mode, interpolation_mode))
    return DRTK_B200_EINVAL;
  PixelArgs a;
  int rc = fill_pixel_args(a, grid, grid_strides, vt_dxdy_img, vt_strides, N, C, H, W, num_levels, max_aniso, padding_mode,
                           align_corners, force_max_aniso, clip_grad);
  if (rc) return rc;
  const int64_t total = N * H * W;
  if (total == 0 || C == 0) return 0;
  for (int i = 0; i < num_levels; ++i)
    if (!levels[i]) return DRTK_B200_EINVAL;
  if (!out) return DRTK_B200_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool bicubic = interpolation_mode != 0;
  if (levels_fit_int32(level_hw, level_strides, num_levels, N, C)) {
    TexList<int32_t> tl;
    if ((rc = fill_levels(tl, levels, nullptr, level_hw, level_strides, num_levels))) return rc;
    if (bicubic) mipmap_fwd_kernel<true, int32_t><<<blocks_for(total), 256, 0, st>>>(tl, a, out);
    else mipmap_fwd_kernel<false, int32_t><<<blocks_for(total), 256, 0, st>>>(tl, a, out);
  } else {
    TexList<int64_t> tl;
    if ((rc = fill_levels(tl, levels, nullptr, level_hw, level_strides, num_levels))) return rc;
    if (bicubic) mipmap_fwd_kernel<true, int64_t><<<blocks_for(total), 256, 0, st>>>(tl, a, out);
    else mipmap_fwd_kernel<false, int64_t><<<blocks_for(total), 256, 0, st>>>(tl, a, out);
  }
  DRTK_CHECK_LAUNCH();
  return 0;
}

extern "C" int drtk_b200_mipmap_grid_sample_backward(
    const float* grad_out, const int64_t* grad_out_strides, const float* const* levels, const int64_t* level_hw,
    const int64_t* level_strides, int num_levels, const float* grid, const int64_t* grid_strides,
    const float* vt_dxdy_img, const int64_t* vt_strides, int64_t N, int64_t C, int64_t H, int64_t W, int max_aniso,
    int padding_mode, int interpolation_mode, int align_corners, int force_max_aniso, int clip_grad,
    float* const* grad_levels, float* grad_grid, void* stream) {
  if (!grad_out_strides || !levels || !level_hw || !level_strides || num_levels < 1 ||
      num_levels > kMaxLevels || bad_enum(padding_mode, interpolation_mode))
    return DRTK_B200_EINVAL;
  PixelArgs a;
  int rc = fill_pixel_args(a, grid, grid_strides, vt_dxdy_img, vt_strides, N, C, H, W, num_levels, max_aniso, padding_mode,
                           align_corners, force_max_aniso, clip_grad);
  if (rc) return rc;
  for (int i = 0; i < num_levels; ++i)
    if (level_hw[2 * i] <= 0 || level_hw[2 * i + 1] <= 0) return DRTK_B200_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (grad_levels)
    for (int i = 0; i < num_levels; ++i)
      if (grad_levels[i] && N * C > 0)
        DRTK_CUDA(cudaMemsetAsync(grad_levels[i], 0, size_t(N) * C * level_hw[2 * i] * level_hw[2 * i + 1] * sizeof(float), st));
  const int64_t total = N * H * W;
  if (total == 0 || C == 0) return 0;
  if (!grad_out) return DRTK_B200_EINVAL;
  for (int i = 0; i < num_levels; ++i)
    if (!levels[i]) return DRTK_B200_EINVAL;
  const bool bicubic = interpolation_mode != 0;
  const Strides4 gos = make4(grad_out_strides);
  if (levels_fit_int32(level_hw, level_strides, num_levels, N, C)) {
    TexList<int32_t> tl;
    if ((rc = fill_levels(tl, levels, grad_levels, level_hw, level_strides, num_levels))) return rc;
    if (bicubic) mipmap_bwd_kernel<true, int32_t><<<blocks_for(total), 256, 0, st>>>(tl, a, grad_out, gos, grad_grid);
    else mipmap_bwd_kernel<false, int32_t><<<blocks_for(total), 256, 0, st>>>(tl, a, grad_out, gos, grad_grid);
  } else {
    TexList<int64_t> tl;
    if ((rc = fill_levels(tl, levels, grad_levels, level_hw, level_strides, num_levels))) return rc;
    if (bicubic) mipmap_bwd_kernel<true, int64_t><<<blocks_for(total), 256, 0, st>>>(tl, a, grad_out, gos, grad_grid);
    else mipmap_bwd_kernel<false, int64_t><<<blocks_for(total), 256, 0, st>>>(tl, a, grad_out, gos, grad_grid);
  }
  DRTK_CHECK_LAUNCH();
  return 0;
}

static int fill_scatter(ScatterArgs& a, const float* input, const int64_t* is, const float* grid, const int64_t* gs,
                        int64_t N, int64_t C, int64_t H, int64_t W, int64_t Ho, int64_t Wo, int pad, int interp,
                        int align) {
  if (!is || !gs || N < 0 || C < 0 || H < 0 || W < 0 || Ho <= 0 || Wo <= 0 || bad_enum(pad, interp))
    return DRTK_B200_EINVAL;
  if (N * C * H * W > 0 && (!input || !grid)) return DRTK_B200_EINVAL;
  if (N > INT32_MAX || C > INT32_MAX || H > INT32_MAX || W > INT32_MAX || Ho > INT32_MAX || Wo > INT32_MAX)
    return DRTK_B200_EUNSUPPORTED;
  a.input = input; a.is = make4(is); a.grid = grid; a.gs = make4(gs);
  a.N = int(N); a.C = int(C); a.H = int(H); a.W = int(W); a.Ho = int(Ho); a.Wo = int(Wo);
  a.pad = pad; a.align = align != 0;
  return 0;
}

extern "C" int drtk_b200_grid_scatter_forward(const float* input, const int64_t* input_strides, const float* grid,
                                              const int64_t* grid_strides, int64_t N, int64_t C, int64_t H, int64_t W,
                                              int64_t out_H, int64_t out_W, int padding_mode, int interpolation_mode,
                                              int align_corners, float* out, void* stream) {
  ScatterArgs a;
  int rc = fill_scatter(a, input, input_strides, grid, grid_strides, N, C, H, W, out_H, out_W, padding_mode,
                        interpolation_mode, align_corners);
  if (rc) return rc;
  if (!out && N * C > 0) return DRTK_B200_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (N * C > 0) DRTK_CUDA(cudaMemsetAsync(out, 0, size_t(N) * C * out_H * out_W * sizeof(float), st));
  const int64_t total = N * H * W;
  if (total == 0 || C == 0) return 0;
  if (interpolation_mode == 0) scatter_fwd_kernel<false><<<blocks_for(total), 256, 0, st>>>(a, out);
  else scatter_fwd_kernel<true><<<blocks_for(total), 256, 0, st>>>(a, out);
  DRTK_CHECK_LAUNCH();
  return 0;
}

extern "C" int drtk_b200_grid_scatter_backward(const float* grad_out, const int64_t* grad_out_strides,
                                               const float* input, const int64_t* input_strides, const float* grid,
                                               const int64_t* grid_strides, int64_t N, int64_t C, int64_t H, int64_t W,
                                               int64_t out_H, int64_t out_W, int padding_mode, int interpolation_mode,
                                               int align_corners, float* grad_input, float* grad_grid, void* stream) {
  ScatterArgs a;
  int rc = fill_scatter(a, input, input_strides, grid, grid_strides, N, C, H, W, out_H, out_W, padding_mode,
                        interpolation_mode, align_corners);
  if (rc) return rc;
  if (!grad_out_strides) return DRTK_B200_EINVAL;
  const int64_t total = N * H * W;
  if (total == 0 || (!grad_input && !grad_grid)) return 0;
  if (!grad_out) return DRTK_B200_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (interpolation_mode == 0)
    scatter_bwd_kernel<false><<<blocks_for(total), 256, 0, st>>>(a, grad_out, make4(grad_out_strides), grad_input, grad_grid);
  else
    scatter_bwd_kernel<true><<<blocks_for(total), 256, 0, st>>>(a, grad_out, make4(grad_out_strides), grad_input, grad_grid);
  DRTK_CHECK_LAUNCH();
  return 0;
}
