// fp64.cu -- double-precision dispatch of the hot path (the reference instantiates every kernel for float and
// double, src/include/kernel_utils.h:47-57).  fp64 is what gradient checks and fp64 reference runs use; it is
// not a throughput path on B200 (fp64 runs at 1/64 of the fp32 rate), so these kernels are deliberately plain:
// one thread per pixel (per triangle for the rasteriser), `atomicAdd(double)` scatters, arbitrary element
// strides, no staging.  The fp32 kernels in the other files are the optimised ones.
//
//   rasterize   rasterize_kernel + unpack_kernel (src/rasterize/rasterize_kernel.cu:42-168, :402-415)
//   render      render_kernel / render_backward_kernel (src/render/render_kernel.cu:19-281)
//   interpolate interpolate_kernel / interpolate_backward_kernel (src/interpolate/interpolate_kernel.cu:38-299)
//   edge_grad   edge_grad_backward_kernel (src/edge_grad/edge_grad_kernel.cu:217-449)
// This file is compiled without --use_fast_math (IEEE division and sqrt in double either way).
#include "common.cuh"

namespace drtk {
namespace {

using T = double;
constexpr T kEps = 1e-16;  // math::epsilon<double>, src/include/cuda_math_helper.h:63-69

__device__ __forceinline__ T epsclamp_d(T v) { return v < 0 ? fmin(v, -kEps) : fmax(v, kEps); }
__device__ __forceinline__ T sign_d(T v) { return v > 0 ? T(1) : (v < 0 ? T(-1) : T(0)); }

struct Vert { T x, y, z; };
__device__ __forceinline__ Vert load_vert(const T* vn, Strides3 s, int i) {
  const T* p = vn + int64_t(i) * s.s1;
  return Vert{p[0], p[s.s2], p[2 * s.s2]};
}
__device__ __forceinline__ void load_tri(const int32_t* vin, Strides3 s, int t, int& i0, int& i1, int& i2) {
  const int32_t* p = vin + int64_t(t) * s.s1;
  i0 = p[0]; i1 = p[s.s2]; i2 = p[2 * s.s2];
}

// ---- rasterize ----------------------------------------------------------------------------
__device__ __forceinline__ T edge_fn(T ax, T ay, T bx, T by, T px, T py) {  // :19-27
  return (py - ay) * (bx - ax) - (px - ax) * (by - ay);
}
__device__ __forceinline__ T canon_edge(int ia, int ib, T ax, T ay, T bx, T by, T px, T py) {  // :29-40
  return ia <= ib ? edge_fn(ax, ay, bx, by, px, py) : -edge_fn(bx, by, ax, ay, px, py);
}
__device__ __forceinline__ void top_left(T den, T v01x, T v01y, T v02x, T v02y, T v12x, T v12y, bool (&tl)[3]) {
  if (den > 0) {  // :133-141
    tl[0] = v12y < 0 || (v12y == 0 && v12x > 0);
    tl[1] = v02y > 0 || (v02y == 0 && v02x < 0);
    tl[2] = v01y < 0 || (v01y == 0 && v01x > 0);
  } else {
    tl[0] = v12y > 0 || (v12y == 0 && v12x < 0);
    tl[1] = v02y < 0 || (v02y == 0 && v02x > 0);
    tl[2] = v01y > 0 || (v01y == 0 && v01x < 0);
  }
}

__global__ void raster_tri_kernel(const T* __restrict__ v, Strides3 vs, const int32_t* __restrict__ vi, Strides3 is,
                                  int N, int F, int H, int W, unsigned long long* __restrict__ packed) {
  const int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (idx >= int64_t(N) * F) return;
  const int n = int(idx / F), id = int(idx % F);
  int i0, i1, i2;
  load_tri(vi + n * is.s0, is, id, i0, i1, i2);
  i0 &= 0x0FFFFFFF;  // :74
  if (i0 == i1 && i1 == i2) return;
  const T* vn = v + n * vs.s0;
  const Vert p0 = load_vert(vn, vs, i0), p1 = load_vert(vn, vs, i1), p2 = load_vert(vn, vs, i2);
  if (!(p0.z > 1e-8f && p1.z > 1e-8f && p2.z > 1e-8f)) return;  // :96 (float literal in the reference)
  const T mnx = fmin(fmin(p0.x, p1.x), p2.x), mny = fmin(fmin(p0.y, p1.y), p2.y);
  const T mxx = fmax(fmax(p0.x, p1.x), p2.x), mxy = fmax(fmax(p0.y, p1.y), p2.y);
  if (!(mnx <= T(W - 1) && mny <= T(H - 1) && mxx > 0 && mxy > 0)) return;
  const T v01x = p1.x - p0.x, v01y = p1.y - p0.y, v02x = p2.x - p0.x, v02y = p2.y - p0.y;
  const T v12x = p2.x - p1.x, v12y = p2.y - p1.y;
  const T den = v01x * v02y - v01y * v02x;
  if (den == 0) return;
  const int bx0 = max(0, int(mnx)), by0 = max(0, int(mny));
  // (the clamp before the conversion keeps `+ 1` from overflowing on huge coordinates)
  const int bx1 = min(W - 1, int(fmin(mxx, T(W))) + 1), by1 = min(H - 1, int(fmin(mxy, T(H))) + 1);
  bool tl[3];
  top_left(den, v01x, v01y, v02x, v02y, v12x, v12y, tl);
  const T s = sign_d(den), aden = fabs(den);
  const T d0 = T(1) / epsclamp_d(p0.z), d1 = T(1) / epsclamp_d(p1.z), d2 = T(1) / epsclamp_d(p2.z);
  unsigned long long* pk = packed + int64_t(n) * H * W;
  for (int y = by0; y <= by1; ++y)
    for (int x = bx0; x <= bx1; ++x) {
      const T px = T(x), py = T(y);
      T b0 = canon_edge(i1, i2, p1.x, p1.y, p2.x, p2.y, px, py) * s;
      T b1 = canon_edge(i2, i0, p2.x, p2.y, p0.x, p0.y, px, py) * s;
      T b2 = canon_edge(i0, i1, p0.x, p0.y, p1.x, p1.y, px, py) * s;
      if (!(b0 >= 0 && b1 >= 0 && b2 >= 0)) continue;
      if ((b0 == 0 && !tl[0]) || (b1 == 0 && !tl[1]) || (b2 == 0 && !tl[2])) continue;
      b0 /= aden; b1 /= aden; b2 /= aden;
      const float depth = float(T(1) / epsclamp_d(d0 * b0 + d1 * b1 + d2 * b2));
      atomicMin(pk + int64_t(y) * W + x, (static_cast<unsigned long long>(__float_as_uint(depth)) << 32) | unsigned(id));
    }
}

// ---- wireframe (src/rasterize/rasterize_kernel.cu:171-400, instantiated for double at :492-535) -------------------
// A pixel carries a triangle's id when one of its VISIBLE edges (bits 0-2 of the top nibble of vi[...,0], :293-303)
// crosses the diamond |dx| + |dy| = 0.5 around the pixel centre (:220-259); the interior still writes depth with id
// 0xFFFFFFFF (= -1) so that surfaces occlude the lines behind them (:376-393).
struct Line64 { T a, b, c; };
__device__ __forceinline__ Line64 line_through(T p1x, T p1y, T p2x, T p2y) {  // :171-181
  return Line64{p1y - p2y, p2x - p1x, p1x * p2y - p2x * p1y};
}
__device__ __forceinline__ bool in_segment(T p1x, T p1y, T p2x, T p2y, T cx, T cy) {  // :183-191
  return (((p2x >= cx) && (cx >= p1x)) || ((p2x <= cx) && (cx <= p1x))) &&
         (((p2y >= cy) && (cy >= p1y)) || ((p2y <= cy) && (cy <= p1y)));
}
__device__ __forceinline__ bool crosses_side(const Line64& l, T p1x, T p1y, T p2x, T p2y, T s0x, T s0y, T s1x, T s1y) {
  const Line64 m = line_through(s0x, s0y, s1x, s1y);
  const T d = l.a * m.b - m.a * l.b;  // :205-218
  T cx = 1.7976931348623157e308, cy = 0;  // TVec2{max}: x = DBL_MAX, y = 0
  if (d != 0) {
    cx = (l.b * m.c - m.b * l.c) / d;
    cy = (m.a * l.c - l.a * m.c) / d;
  }
  return in_segment(s0x, s0y, s1x, s1y, cx, cy) && in_segment(p1x, p1y, p2x, p2y, cx, cy);
}
__device__ __forceinline__ bool crosses_diamond(T p1x, T p1y, T p2x, T p2y, T px, T py) {  // :220-259
  const Line64 l = line_through(p1x, p1y, p2x, p2y);
  const T h = 0.5;
  bool hit = crosses_side(l, p1x, p1y, p2x, p2y, px, py - h, px + h, py);
  hit |= crosses_side(l, p1x, p1y, p2x, p2y, px + h, py, px, py + h);
  hit |= crosses_side(l, p1x, p1y, p2x, p2y, px, py + h, px - h, py);
  hit |= crosses_side(l, p1x, p1y, p2x, p2y, px - h, py, px, py - h);
  return hit;
}

// one WARP per triangle: the lanes stride over the padded bounding box (the reference walks it with one thread)
__global__ void raster_lines_kernel64(const T* __restrict__ v, Strides3 vs, const int32_t* __restrict__ vi, Strides3 is,
                                      int N, int F, int H, int W, unsigned long long* __restrict__ packed) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5, nwarps = (gridDim.x * int64_t(blockDim.x)) >> 5;
  for (int64_t idx = warp0; idx < int64_t(N) * F; idx += nwarps) {
    const int n = int(idx / F), id = int(idx % F);
    int i0, i1, i2;
    load_tri(vi + n * is.s0, is, id, i0, i1, i2);
    const int flag = int((unsigned(i0) & 0xF0000000u) >> 28);
    i0 &= 0x0FFFFFFF;
    if (i0 == i1 && i1 == i2) continue;  // :296
    const bool vis0 = flag & 1, vis1 = flag & 2, vis2 = flag & 4;
    const T* vn = v + n * vs.s0;
    const Vert p0 = load_vert(vn, vs, i0), p1 = load_vert(vn, vs, i1), p2 = load_vert(vn, vs, i2);
    if (!(p0.z > 1e-8f && p1.z > 1e-8f && p2.z > 1e-8f)) continue;  // :321
    const T mnx = fmin(fmin(p0.x, p1.x), p2.x), mny = fmin(fmin(p0.y, p1.y), p2.y);
    const T mxx = fmax(fmax(p0.x, p1.x), p2.x), mxy = fmax(fmax(p0.y, p1.y), p2.y);
    if (!(mnx <= T(W - 1) && mny <= T(H - 1) && mxx > 0 && mxy > 0)) continue;  // :322-323
    const T v01x = p1.x - p0.x, v01y = p1.y - p0.y, v02x = p2.x - p0.x, v02y = p2.y - p0.y;
    const T v12x = p2.x - p1.x, v12y = p2.y - p1.y;
    const T den = v01x * v02y - v01y * v02x;  // :330
    if (den == 0) continue;
    // bounds with the extra border (:333-337); the clamps before the conversions keep them from overflowing
    const int bx0 = max(1, int(fmax(mnx, T(-4))) - 2), by0 = max(1, int(fmax(mny, T(-4))) - 2);
    const int bx1 = min(W - 2, int(fmin(mxx, T(W))) + 2), by1 = min(H - 2, int(fmin(mxy, T(H))) + 2);
    if (bx0 > bx1 || by0 > by1) continue;
    bool tl[3];
    top_left(den, v01x, v01y, v02x, v02y, v12x, v12y, tl);
    const T s = sign_d(den), aden = fabs(den);
    const T d0 = T(1) / epsclamp_d(p0.z), d1 = T(1) / epsclamp_d(p1.z), d2 = T(1) / epsclamp_d(p2.z);
    unsigned long long* pk = packed + int64_t(n) * H * W;
    const int bw = bx1 - bx0 + 1;
    const int64_t count = int64_t(bw) * (by1 - by0 + 1);
    for (int64_t q = lane; q < count; q += 32) {
      const int y = by0 + int(q / bw), x = bx0 + int(q % bw);
      const T px = T(x), py = T(y);
      bool hit = crosses_diamond(p0.x, p0.y, p1.x, p1.y, px, py) && vis0;  // :343-346
      hit |= crosses_diamond(p1.x, p1.y, p2.x, p2.y, px, py) && vis1;
      hit |= crosses_diamond(p0.x, p0.y, p2.x, p2.y, px, py) && vis2;
      T b0 = canon_edge(i1, i2, p1.x, p1.y, p2.x, p2.y, px, py) * s;  // :348-353
      T b1 = canon_edge(i2, i0, p2.x, p2.y, p0.x, p0.y, px, py) * s;
      T b2 = canon_edge(i0, i1, p0.x, p0.y, p1.x, p1.y, px, py) * s;
      const bool inside = b0 >= 0 && b1 >= 0 && b2 >= 0;
      const bool keep = inside && !((b0 == 0 && !tl[0]) || (b1 == 0 && !tl[1]) || (b2 == 0 && !tl[2]));
      if (!(keep || hit)) continue;  // :375
      b0 = fmin(fmax(b0 / aden, T(0)), T(1)); b1 = fmin(fmax(b1 / aden, T(0)), T(1)); b2 = fmin(fmax(b2 / aden, T(0)), T(1));
      const T sum = b0 + b1 + b2;
      b0 /= sum; b1 /= sum; b2 /= sum;
      const float depth = float(T(1) / epsclamp_d(d0 * b0 + d1 * b1 + d2 * b2));
      atomicMin(pk + int64_t(y) * W + x,
                (static_cast<unsigned long long>(__float_as_uint(depth)) << 32) | (hit ? (unsigned long long)unsigned(id) : 0xFFFFFFFFull));
    }
  }
}

__global__ void unpack_kernel(const unsigned long long* __restrict__ packed, int64_t total, float* __restrict__ depth,
                              int32_t* __restrict__ index) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const unsigned long long p = packed[i];
  const unsigned hi = unsigned(p >> 32);
  depth[i] = hi == 0xFFFFFFFFu ? 0.f : __uint_as_float(hi);  // :409-413
  index[i] = int32_t(unsigned(p));
}

// ---- render ---------------------------------------------------------------------------------
struct RenderPx {
  int i0, i1, i2;
  T v01x, v01y, v02x, v02y, den_raw, den, qx, qy, b0, b1, b2, z0e, z1e, z2e, d0, d1, d2, dinv, dinv_e, depth;
  bool c0, c1, c2;
  __device__ __forceinline__ RenderPx(const T* vn, Strides3 vs, const int32_t* vin, Strides3 is, int t, int w, int h) {
    load_tri(vin, is, t, i0, i1, i2);
    const Vert p0 = load_vert(vn, vs, i0), p1 = load_vert(vn, vs, i1), p2 = load_vert(vn, vs, i2);
    v01x = p1.x - p0.x; v01y = p1.y - p0.y; v02x = p2.x - p0.x; v02y = p2.y - p0.y;
    den_raw = v01x * v02y - v01y * v02x;
    den = epsclamp_d(den_raw);
    qx = T(w) - p0.x; qy = T(h) - p0.y;
    b1 = (qx * v02y - qy * v02x) / den;
    b2 = (qy * v01x - qx * v01y) / den;
    b0 = T(1) - b1 - b2;
    z0e = epsclamp_d(p0.z); z1e = epsclamp_d(p1.z); z2e = epsclamp_d(p2.z);
    c0 = z0e != p0.z; c1 = z1e != p1.z; c2 = z2e != p2.z;
    d0 = T(1) / z0e; d1 = T(1) / z1e; d2 = T(1) / z2e;
    dinv = d0 * b0 + d1 * b1 + d2 * b2;
    dinv_e = epsclamp_d(dinv);
    depth = T(1) / dinv_e;
  }
};

struct ImgArgs { int N, H, W; };

__global__ void render_fwd_kernel64(const T* __restrict__ v, Strides3 vs, const int32_t* __restrict__ vi, Strides3 is,
                                    const int32_t* __restrict__ index, Strides3 xs, ImgArgs a, T* __restrict__ depth,
                                    T* __restrict__ bary) {
  const int64_t HW = int64_t(a.H) * a.W, idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (idx >= a.N * HW) return;
  const int n = int(idx / HW), h = int((idx % HW) / a.W), w = int(idx % a.W);
  const int t = index[n * xs.s0 + h * xs.s1 + w * xs.s2];
  T* bo = bary + int64_t(n) * 3 * HW + (idx % HW);
  if (t == -1) { bo[0] = 0; bo[HW] = 0; bo[2 * HW] = 0; depth[idx] = 0; return; }  // :110-115
  const RenderPx r(v + n * vs.s0, vs, vi + n * is.s0, is, t, w, h);
  bo[0] = r.d0 * r.b0 * r.depth; bo[HW] = r.d1 * r.b1 * r.depth; bo[2 * HW] = r.d2 * r.b2 * r.depth;
  depth[idx] = r.depth;
}

__global__ void render_bwd_kernel64(const T* __restrict__ v, Strides3 vs, const int32_t* __restrict__ vi, Strides3 is,
                                    const int32_t* __restrict__ index, Strides3 xs, const T* __restrict__ gdepth,
                                    Strides3 gds, const T* __restrict__ gbary, Strides4 gbs, ImgArgs a, int V,
                                    T* __restrict__ grad_v) {
  const int64_t HW = int64_t(a.H) * a.W, idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (idx >= a.N * HW) return;
  const int n = int(idx / HW), h = int((idx % HW) / a.W), w = int(idx % a.W);
  const int t = index[n * xs.s0 + h * xs.s1 + w * xs.s2];
  if (t == -1) return;
  const RenderPx r(v + n * vs.s0, vs, vi + n * is.s0, is, t, w, h);
  T g0 = 0, g1 = 0, g2 = 0, gd = 0;
  if (gbary) {
    const T* p = gbary + n * gbs.s0 + h * gbs.s2 + w * gbs.s3;
    g0 = p[0]; g1 = p[gbs.s1]; g2 = p[2 * gbs.s1];
  }
  if (gdepth) gd = gdepth[n * gds.s0 + h * gds.s1 + w * gds.s2];
  const bool den_clamped = r.den != r.den_raw, dinv_clamped = r.dinv_e != r.dinv;
  const T dL_depth = gd + (g0 * r.d0 * r.b0 + g1 * r.d1 * r.b1 + g2 * r.d2 * r.b2);     // :226
  const T dL_dinv = dinv_clamped ? T(0) : -dL_depth / (r.dinv * r.dinv);                 // :228-229
  const T dLd0 = g0 * r.b0 * r.depth + dL_dinv * r.b0, dLd1 = g1 * r.b1 * r.depth + dL_dinv * r.b1,
          dLd2 = g2 * r.b2 * r.depth + dL_dinv * r.b2;
  T* gv = grad_v + int64_t(n) * V * 3;
  atomicAdd(gv + r.i0 * 3 + 2, r.c0 ? T(0) : -dLd0 / (r.z0e * r.z0e));                   // :231-250
  atomicAdd(gv + r.i1 * 3 + 2, r.c1 ? T(0) : -dLd1 / (r.z1e * r.z1e));
  atomicAdd(gv + r.i2 * 3 + 2, r.c2 ? T(0) : -dLd2 / (r.z2e * r.z2e));
  const T dLb0 = g0 * r.d0 * r.depth + dL_dinv * r.d0, dLb1 = g1 * r.d1 * r.depth + dL_dinv * r.d1,
          dLb2 = g2 * r.d2 * r.depth + dL_dinv * r.d2;
  const T e1 = (-dLb0 + dLb1) / r.den, e2 = (-dLb0 + dLb2) / r.den;                      // :253-254
  const T dL_den = den_clamped ? T(0) : -(e1 * r.b1 + e2 * r.b2);                        // :256
  const T dqx = e1 * r.v02y - e2 * r.v01y, dqy = -e1 * r.v02x + e2 * r.v01x;
  const T dv02x = -e1 * r.qy - dL_den * r.v01y, dv02y = e1 * r.qx + dL_den * r.v01x;
  const T dv01x = e2 * r.qy + dL_den * r.v02y, dv01y = -e2 * r.qx - dL_den * r.v02x;
  atomicAdd(gv + r.i0 * 3, -dv02x - dv01x - dqx); atomicAdd(gv + r.i0 * 3 + 1, -dv02y - dv01y - dqy);  // :269-278
  atomicAdd(gv + r.i1 * 3, dv01x); atomicAdd(gv + r.i1 * 3 + 1, dv01y);
  atomicAdd(gv + r.i2 * 3, dv02x); atomicAdd(gv + r.i2 * 3 + 1, dv02y);
}

// ---- interpolate ----------------------------------------------------------------------------
__global__ void interp_fwd_kernel64(const T* __restrict__ attr, Strides3 as, const int32_t* __restrict__ vi, Strides3 is,
                                    const int32_t* __restrict__ index, Strides3 xs, const T* __restrict__ bary,
                                    Strides4 bs, ImgArgs a, int C, T* __restrict__ out) {
  const int64_t HW = int64_t(a.H) * a.W, idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (idx >= a.N * HW) return;
  const int n = int(idx / HW), h = int((idx % HW) / a.W), w = int(idx % a.W);
  const int t = index[n * xs.s0 + h * xs.s1 + w * xs.s2];
  T* po = out + int64_t(n) * C * HW + (idx % HW);
  if (t == -1) {  // coordinate sweep, computed in float like the reference (:104-109)
    const T sx = T((float(w) * 2.0f + 1.0f) / float(a.W) - 1.0f), sy = T((float(h) * 2.0f + 1.0f) / float(a.H) - 1.0f);
    for (int c = 0; c < C; ++c) po[int64_t(c) * HW] = (c & 1) ? sy : sx;
    return;
  }
  int i0, i1, i2;
  load_tri(vi + n * is.s0, is, t, i0, i1, i2);
  const T* pb = bary + n * bs.s0 + h * bs.s2 + w * bs.s3;
  const T b0 = pb[0], b1 = pb[bs.s1], b2 = pb[2 * bs.s1];
  const T* an = attr + n * as.s0;
  const T *a0 = an + i0 * as.s1, *a1 = an + i1 * as.s1, *a2 = an + i2 * as.s1;
  for (int c = 0; c < C; ++c) po[int64_t(c) * HW] = a0[c * as.s2] * b0 + a1[c * as.s2] * b1 + a2[c * as.s2] * b2;
}

__global__ void interp_bwd_kernel64(const T* __restrict__ gout, Strides4 gs, const T* __restrict__ attr, Strides3 as,
                                    const int32_t* __restrict__ vi, Strides3 is, const int32_t* __restrict__ index,
                                    Strides3 xs, const T* __restrict__ bary, Strides4 bs, ImgArgs a, int V, int C,
                                    T* __restrict__ attr_grad, T* __restrict__ bary_grad) {
  const int64_t HW = int64_t(a.H) * a.W, idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (idx >= a.N * HW) return;
  const int n = int(idx / HW), h = int((idx % HW) / a.W), w = int(idx % a.W);
  const int t = index[n * xs.s0 + h * xs.s1 + w * xs.s2];
  T* pbg = bary_grad ? bary_grad + int64_t(n) * 3 * HW + (idx % HW) : nullptr;
  if (t == -1) {
    if (pbg) { pbg[0] = 0; pbg[HW] = 0; pbg[2 * HW] = 0; }  // :282-297
    return;
  }
  int i0, i1, i2;
  load_tri(vi + n * is.s0, is, t, i0, i1, i2);
  const T* pb = bary + n * bs.s0 + h * bs.s2 + w * bs.s3;
  const T b0 = pb[0], b1 = pb[bs.s1], b2 = pb[2 * bs.s1];
  const T* an = attr + n * as.s0;
  const T *a0 = an + i0 * as.s1, *a1 = an + i1 * as.s1, *a2 = an + i2 * as.s1;
  const T* pg = gout + n * gs.s0 + h * gs.s2 + w * gs.s3;
  T* gn = attr_grad ? attr_grad + int64_t(n) * V * C : nullptr;
  T gb0 = 0, gb1 = 0, gb2 = 0;
  for (int c = 0; c < C; ++c) {
    const T g = pg[c * gs.s1];
    gb0 += g * a0[c * as.s2]; gb1 += g * a1[c * as.s2]; gb2 += g * a2[c * as.s2];
    if (gn) {
      atomicAdd(gn + int64_t(i0) * C + c, g * b0);
      atomicAdd(gn + int64_t(i1) * C + c, g * b1);
      atomicAdd(gn + int64_t(i2) * C + c, g * b2);
    }
  }
  if (pbg) { pbg[0] = gb0; pbg[HW] = gb1; pbg[2 * HW] = gb2; }
}

// ---- edge_grad ------------------------------------------------------------------------------
struct TriInfo { T p0x, p0y, p1x, p1y, v01x, v01y, v02x, v02y, v12x, v12y, den; };
__device__ __forceinline__ TriInfo tri_info(const T* vn, Strides3 vs, int i0, int i1, int i2) {  // :72-87
  const Vert p0 = load_vert(vn, vs, i0), p1 = load_vert(vn, vs, i1), p2 = load_vert(vn, vs, i2);
  TriInfo t;
  t.p0x = p0.x; t.p0y = p0.y; t.p1x = p1.x; t.p1y = p1.y;
  t.v01x = p1.x - p0.x; t.v01y = p1.y - p0.y; t.v02x = p2.x - p0.x; t.v02y = p2.y - p0.y;
  t.v12x = p2.x - p1.x; t.v12y = p2.y - p1.y;
  t.den = t.v01x * t.v02y - t.v01y * t.v02x;
  return t;
}
__device__ __forceinline__ bool pix_in_tri(const TriInfo& t, int x, int y) {  // :30-70, plain edge functions
  if (t.den == 0) return false;
  const T px = T(x), py = T(y), q0x = px - t.p0x, q0y = py - t.p0y, q1x = px - t.p1x, q1y = py - t.p1y;
  const T s = sign_d(t.den);
  const T b0 = (q1y * t.v12x - q1x * t.v12y) * s, b1 = (q0x * t.v02y - q0y * t.v02x) * s,
          b2 = (q0y * t.v01x - q0x * t.v01y) * s;
  if (!(b0 >= 0 && b1 >= 0 && b2 >= 0)) return false;
  bool tl[3];
  top_left(t.den, t.v01x, t.v01y, t.v02x, t.v02y, t.v12x, t.v12y, tl);
  return !((b0 == 0 && !tl[0]) || (b1 == 0 && !tl[1]) || (b2 == 0 && !tl[2]));
}
__device__ __forceinline__ void tri_normal(const T* vn, Strides3 vs, int i0, int i1, int i2, T (&nrm)[3]) {  // :89-100
  const Vert p0 = load_vert(vn, vs, i0), p1 = load_vert(vn, vs, i1), p2 = load_vert(vn, vs, i2);
  const T ax = p0.x - p2.x, ay = p0.y - p2.y, az = p0.z - p2.z, bx = p1.x - p0.x, by = p1.y - p0.y, bz = p1.z - p0.z;
  const T cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx;
  const T r = rnorm3d(cx, cy, cz);
  nrm[0] = cx * r; nrm[1] = cy * r; nrm[2] = cz * r;
}
__device__ __forceinline__ void dp_dr(T nvx, T nvy, T nfx, T nfy, T max_mag, T& ox, T& oy) {  // :102-203
  const T rv = rsqrt(nvx * nvx + nvy * nvy), rf = rsqrt(nfx * nfx + nfy * nfy);
  nvx *= rv; nvy *= rv; nfx *= rf; nfy *= rf;
  const T bx = -nfy, by = nfx, d = bx * nvx + by * nvy;
  T k;
  if (max_mag > 0) {
    const T safe = (d >= 0 ? T(1) : T(-1)) * epsclamp_d(fmax(fabs(d), fabs(bx) / max_mag));
    k = bx / safe;
  } else {
    k = bx / epsclamp_d(d);
  }
  ox = k * nvx; oy = k * nvy;
}

__global__ void edge_grad_bwd_kernel64(const T* __restrict__ v, Strides3 vs, const T* __restrict__ img, Strides4 ms,
                                       const int32_t* __restrict__ index, Strides3 xs, const int32_t* __restrict__ vi,
                                       Strides3 is, const T* __restrict__ gout, Strides4 gs, ImgArgs a, int C,
                                       T max_dp_dr, T* __restrict__ out) {
  const int64_t HW = int64_t(a.H) * a.W, idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (idx >= a.N * HW) return;
  const int n = int(idx / HW), y = int((idx % HW) / a.W), x = int(idx % a.W);
  if (x >= a.W - 1 || y >= a.H - 1) return;  // :270
  const int32_t* ix = index + n * xs.s0;
  const int ci = ix[y * xs.s1 + x * xs.s2], ri = ix[y * xs.s1 + (x + 1) * xs.s2], di = ix[(y + 1) * xs.s1 + x * xs.s2];
  const bool cv = ci >= 0, rv = ri >= 0, dv = di >= 0;
  const bool lr = ci != ri, ud = ci != di;
  if (!lr && !ud) return;
  const int32_t* vin = vi + n * is.s0;
  const T* vn = v + n * vs.s0;
  int c0 = 0, c1 = 0, c2 = 0, r0 = 0, r1 = 0, r2 = 0, e0 = 0, e1 = 0, e2 = 0;  // :296-301
  if (cv) load_tri(vin, is, ci, c0, c1, c2);
  if (rv) load_tri(vin, is, ri, r0, r1, r2);
  if (dv) load_tri(vin, is, di, e0, e1, e2);
  const bool xb = cv && rv, yb = cv && dv;
  const TriInfo tc = tri_info(vn, vs, c0, c1, c2), tr = tri_info(vn, vs, r0, r1, r2), td = tri_info(vn, vs, e0, e1, e2);
  const bool c_in_r = lr && xb && pix_in_tri(tr, x, y), r_in_c = lr && xb && pix_in_tri(tc, x + 1, y);  // :320-325
  const bool c_in_d = ud && yb && pix_in_tri(td, x, y), d_in_c = ud && yb && pix_in_tri(tc, x, y + 1);
  const bool l_over_r = c_in_r && !r_in_c, r_over_l = r_in_c && !c_in_r, u_over_d = c_in_d && !d_in_c,
             d_over_u = d_in_c && !c_in_d;
  const bool horiz_int = c_in_r && r_in_c, vert_int = c_in_d && d_in_c;
  const bool horiz_adj = lr && xb && !c_in_r && !r_in_c, vert_adj = ud && yb && !c_in_d && !d_in_c;
  const T* im = img + n * ms.s0 + y * ms.s2 + x * ms.s3;
  const T* go = gout + n * gs.s0 + y * gs.s2 + x * gs.s3;
  T gdx = 0, gdy = 0;
  for (int c = 0; c < C; ++c) {  // :351-380
    const T ic = im[c * ms.s1], gc = go[c * gs.s1];
    if (lr) gdx += (im[c * ms.s1 + ms.s3] - ic) * (T(0.5) * (go[c * gs.s1 + gs.s3] + gc));
    if (ud) gdy += (im[c * ms.s1 + ms.s2] - ic) * (T(0.5) * (go[c * gs.s1 + gs.s2] + gc));
  }
  T gc3[3] = {0, 0, 0}, gr3[3] = {0, 0, 0}, gd3[3] = {0, 0, 0};
  if (!horiz_int) {  // :391-393
    gc3[0] += (!cv || r_over_l || horiz_adj) ? T(0) : gdx;
    gr3[0] += (!rv || l_over_r || horiz_adj) ? T(0) : gdx;
  } else {  // :394-406
    T nc[3], nr[3], ox, oz;
    tri_normal(vn, vs, c0, c1, c2, nc); tri_normal(vn, vs, r0, r1, r2, nr);
    dp_dr(nc[0], nc[2], nr[0], nr[2], max_dp_dr, ox, oz); gc3[0] += gdx * ox; gc3[2] += gdx * oz;
    dp_dr(nr[0], nr[2], nc[0], nc[2], max_dp_dr, ox, oz); gr3[0] += gdx * ox; gr3[2] += gdx * oz;
  }
  if (!vert_int) {  // :408-410
    gc3[1] += (!cv || d_over_u || vert_adj) ? T(0) : gdy;
    gd3[1] += (!dv || u_over_d || vert_adj) ? T(0) : gdy;
  } else {  // :411-423
    T nc[3], nd[3], oy, oz;
    tri_normal(vn, vs, c0, c1, c2, nc); tri_normal(vn, vs, e0, e1, e2, nd);
    dp_dr(nc[1], nc[2], nd[1], nd[2], max_dp_dr, oy, oz); gc3[1] += gdy * oy; gc3[2] += gdy * oz;
    dp_dr(nd[1], nd[2], nc[1], nc[2], max_dp_dr, oy, oz); gd3[1] += gdy * oy; gd3[2] += gdy * oz;
  }
  T* o = out + int64_t(n) * 3 * HW + int64_t(y) * a.W + x;
#pragma unroll
  for (int k = 0; k < 3; ++k) {  // negated sums, :427-445
    if (gc3[k] != 0) atomicAdd(o + k * HW, -gc3[k]);
    if (gr3[k] != 0) atomicAdd(o + k * HW + 1, -gr3[k]);
    if (gd3[k] != 0) atomicAdd(o + k * HW + a.W, -gd3[k]);
  }
}

inline unsigned nblk(int64_t n) { return unsigned((n + 255) / 256); }
inline bool too_big(int64_t N, int64_t H, int64_t W) {
  return N > INT32_MAX || H > INT32_MAX || W > INT32_MAX || N * H * W / 256 > INT32_MAX;
}

}  // namespace
}  // namespace drtk

using namespace drtk;

extern "C" size_t drtk_b200_rasterize_f64_workspace_bytes(int64_t N, int64_t H, int64_t W) {
  return size_t(N) * H * W * sizeof(unsigned long long);
}

extern "C" int drtk_b200_rasterize_f64(const double* v, const int64_t* v_strides, const int32_t* vi,
                                       const int64_t* vi_strides, int64_t N, int64_t V, int64_t F, int64_t H, int64_t W,
                                       int wireframe, float* depth_img, int32_t* index_img, void* workspace,
                                       size_t workspace_bytes, void* stream) {
  if (!v_strides || !vi_strides || N < 0 || V < 0 || F < 0 || H <= 0 || W <= 0) return DRTK_B200_EINVAL;
  if (too_big(N, H, W) || N * F / 256 > INT32_MAX || F > INT32_MAX) return DRTK_B200_EUNSUPPORTED;
  if (N == 0) return 0;
  if (!depth_img || !index_img || (F > 0 && (!v || !vi))) return DRTK_B200_EINVAL;
  if (!workspace || workspace_bytes < drtk_b200_rasterize_f64_workspace_bytes(N, H, W)) return DRTK_B200_EWORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto* packed = static_cast<unsigned long long*>(workspace);
  DRTK_CUDA(cudaMemsetAsync(packed, 0xFF, size_t(N) * H * W * 8, st));  // :484-488
  if (F > 0 && wireframe) {
    const int64_t want = (N * F * 32 + 255) / 256, cap = int64_t(num_sms()) * 32;
    raster_lines_kernel64<<<unsigned(want < cap ? want : cap), 256, 0, st>>>(v, make3(v_strides), vi, make3(vi_strides), int(N),
                                                                             int(F), int(H), int(W), packed);
    DRTK_CHECK_LAUNCH();
  } else if (F > 0) {
    raster_tri_kernel<<<nblk(N * F), 256, 0, st>>>(v, make3(v_strides), vi, make3(vi_strides), int(N), int(F), int(H),
                                                   int(W), packed);
    DRTK_CHECK_LAUNCH();
  }
  unpack_kernel<<<nblk(N * H * W), 256, 0, st>>>(packed, N * H * W, depth_img, index_img);
  DRTK_CHECK_LAUNCH();
  return 0;
}

extern "C" int drtk_b200_render_forward_f64(const double* v, const int64_t* v_strides, const int32_t* vi,
                                            const int64_t* vi_strides, const int32_t* index_img,
                                            const int64_t* index_strides, int64_t N, int64_t V, int64_t F, int64_t H,
                                            int64_t W, double* depth_img, double* bary_img, void* stream) {
  if (!v_strides || !vi_strides || !index_strides || N < 0 || H < 0 || W < 0) return DRTK_B200_EINVAL;
  if (too_big(N, H, W)) return DRTK_B200_EUNSUPPORTED;
  if (N * H * W == 0) return 0;
  if (!v || !vi || !index_img || !depth_img || !bary_img) return DRTK_B200_EINVAL;
  render_fwd_kernel64<<<nblk(N * H * W), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      v, make3(v_strides), vi, make3(vi_strides), index_img, make3(index_strides), ImgArgs{int(N), int(H), int(W)},
      depth_img, bary_img);
  DRTK_CHECK_LAUNCH();
  return 0;
}

extern "C" int drtk_b200_render_backward_f64(const double* v, const int64_t* v_strides, const int32_t* vi,
                                             const int64_t* vi_strides, const int32_t* index_img,
                                             const int64_t* index_strides, const double* grad_depth,
                                             const int64_t* grad_depth_strides, const double* grad_bary,
                                             const int64_t* grad_bary_strides, int64_t N, int64_t V, int64_t F,
                                             int64_t H, int64_t W, double* grad_v, void* stream) {
  if (!v_strides || !vi_strides || !index_strides || N < 0 || V < 0 || H < 0 || W < 0) return DRTK_B200_EINVAL;
  if ((grad_depth && !grad_depth_strides) || (grad_bary && !grad_bary_strides)) return DRTK_B200_EINVAL;
  if (too_big(N, H, W)) return DRTK_B200_EUNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (N * V > 0) {
    if (!grad_v) return DRTK_B200_EINVAL;
    DRTK_CUDA(cudaMemsetAsync(grad_v, 0, size_t(N) * V * 3 * sizeof(double), st));
  }
  if (N * H * W == 0 || V == 0 || (!grad_depth && !grad_bary)) return 0;
  if (!v || !vi || !index_img) return DRTK_B200_EINVAL;
  const Strides3 z3{0, 0, 0};
  const Strides4 z4{0, 0, 0, 0};
  render_bwd_kernel64<<<nblk(N * H * W), 256, 0, st>>>(
      v, make3(v_strides), vi, make3(vi_strides), index_img, make3(index_strides), grad_depth,
      grad_depth ? make3(grad_depth_strides) : z3, grad_bary, grad_bary ? make4(grad_bary_strides) : z4,
      ImgArgs{int(N), int(H), int(W)}, int(V), grad_v);
  DRTK_CHECK_LAUNCH();
  return 0;
}

extern "C" int drtk_b200_interpolate_forward_f64(const double* vert_attributes, const int64_t* attr_strides,
                                                 const int32_t* vi, const int64_t* vi_strides, const int32_t* index_img,
                                                 const int64_t* index_strides, const double* bary_img,
                                                 const int64_t* bary_strides, int64_t N, int64_t V, int64_t F, int64_t C,
                                                 int64_t H, int64_t W, double* out, void* stream) {
  if (!attr_strides || !vi_strides || !index_strides || !bary_strides || N < 0 || C < 0 || H < 0 || W < 0)
    return DRTK_B200_EINVAL;
  if (too_big(N, H, W) || C > INT32_MAX) return DRTK_B200_EUNSUPPORTED;
  if (N * H * W * C == 0) return 0;
  if (!vert_attributes || !vi || !index_img || !bary_img || !out) return DRTK_B200_EINVAL;
  interp_fwd_kernel64<<<nblk(N * H * W), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      vert_attributes, make3(attr_strides), vi, make3(vi_strides), index_img, make3(index_strides), bary_img,
      make4(bary_strides), ImgArgs{int(N), int(H), int(W)}, int(C), out);
  DRTK_CHECK_LAUNCH();
  return 0;
}

extern "C" int drtk_b200_interpolate_backward_f64(const double* grad_out, const int64_t* grad_out_strides,
                                                  const double* vert_attributes, const int64_t* attr_strides,
                                                  const int32_t* vi, const int64_t* vi_strides, const int32_t* index_img,
                                                  const int64_t* index_strides, const double* bary_img,
                                                  const int64_t* bary_strides, int64_t N, int64_t V, int64_t F, int64_t C,
                                                  int64_t H, int64_t W, double* vert_attributes_grad,
                                                  double* bary_img_grad, void* stream) {
  if (!grad_out_strides || !attr_strides || !vi_strides || !index_strides || !bary_strides || N < 0 || V < 0 || C < 0 ||
      H < 0 || W < 0)
    return DRTK_B200_EINVAL;
  if (too_big(N, H, W) || C > INT32_MAX) return DRTK_B200_EUNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (vert_attributes_grad && N * V * C > 0)
    DRTK_CUDA(cudaMemsetAsync(vert_attributes_grad, 0, size_t(N) * V * C * sizeof(double), st));
  if (N * H * W == 0 || (!vert_attributes_grad && !bary_img_grad)) return 0;
  if (!grad_out || !vert_attributes || !vi || !index_img || !bary_img) return DRTK_B200_EINVAL;
  interp_bwd_kernel64<<<nblk(N * H * W), 256, 0, st>>>(
      grad_out, make4(grad_out_strides), vert_attributes, make3(attr_strides), vi, make3(vi_strides), index_img,
      make3(index_strides), bary_img, make4(bary_strides), ImgArgs{int(N), int(H), int(W)}, int(V), int(C),
      vert_attributes_grad, bary_img_grad);
  DRTK_CHECK_LAUNCH();
  return 0;
}

extern "C" int drtk_b200_edge_grad_backward_f64(const double* v_pix, const int64_t* v_strides, const double* img,
                                                const int64_t* img_strides, const int32_t* index_img,
                                                const int64_t* index_strides, const int32_t* vi,
                                                const int64_t* vi_strides, const double* grad_output,
                                                const int64_t* grad_output_strides, int64_t N, int64_t V, int64_t F,
                                                int64_t C, int64_t H, int64_t W, double max_dp_dr,
                                                double* grad_v_pix_img, void* stream) {
  if (!v_strides || !img_strides || !index_strides || !vi_strides || !grad_output_strides || N < 0 || C < 0 || H < 0 ||
      W < 0)
    return DRTK_B200_EINVAL;
  if (too_big(N, H, W) || C > INT32_MAX) return DRTK_B200_EUNSUPPORTED;
  if (N * H * W == 0) return 0;
  if (!v_pix || !img || !index_img || !vi || !grad_output || !grad_v_pix_img) return DRTK_B200_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DRTK_CUDA(cudaMemsetAsync(grad_v_pix_img, 0, size_t(N) * 3 * H * W * sizeof(double), st));
  edge_grad_bwd_kernel64<<<nblk(N * H * W), 256, 0, st>>>(
      v_pix, make3(v_strides), img, make4(img_strides), index_img, make3(index_strides), vi, make3(vi_strides),
      grad_output, make4(grad_output_strides), ImgArgs{int(N), int(H), int(W)}, int(C), max_dp_dr, grad_v_pix_img);
  DRTK_CHECK_LAUNCH();
  return 0;
}
