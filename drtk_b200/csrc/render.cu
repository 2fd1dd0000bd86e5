// render.cu -- per-pixel perspective-correct barycentrics + depth, forward and backward.
//
// Semantics: src/render/render_kernel.cu:19-117 (forward) and :119-281 (backward) of the
// reference.  fp32 results agree with the reference within 1e-5 relative (they are not
// bit-pinned: only rasterize has a bit-exact contract).
//
// Forward: 20 B/px of compulsory HBM traffic (4 read + 16 written).  One thread owns four
// horizontally adjacent pixels: one 128-bit streaming load of index_img, four 128-bit
// streaming stores (3 bary planes + depth).  Vertex/index gathers hit L1/L2 (tables are
// a few MB, L2 is 126 MB).
//
// Backward: 20 B/px read.  Each covered pixel yields 9 partial derivatives for the 3 vertices
// of its triangle.  Pixels of a warp are consecutive along x, so runs of equal triangle id are
// reduced with a segmented shuffle scan and only the head lane of a run issues the 9
// reductions (REDG.ADD.F32) -- ~5x fewer atomics than one-per-pixel on the 100k-triangle
// config, many more on large triangles.
#include "common.cuh"

namespace drtk {
namespace {

struct RenderArgs {
  const float* v;
  Strides3 vs;
  const int32_t* vi;
  Strides3 vis;
  const int32_t* index_img;
  Strides3 is;
  int N, V, F, H, W;
};

struct TriVerts {
  float p0x, p0y, z0, p1x, p1y, z1, p2x, p2y, z2;
  int i0, i1, i2;
};

__device__ __forceinline__ void load_tri(const RenderArgs& a, int n, int t, TriVerts& r) {
  const int32_t* vip = a.vi + (int64_t)n * a.vis.s0 + (int64_t)t * a.vis.s1;
  r.i0 = vip[0]; r.i1 = vip[a.vis.s2]; r.i2 = vip[2 * a.vis.s2];  // not nibble-masked (:70-72)
  const float* vp = a.v + (int64_t)n * a.vs.s0;
  const float* q0 = vp + (int64_t)r.i0 * a.vs.s1;
  const float* q1 = vp + (int64_t)r.i1 * a.vs.s1;
  const float* q2 = vp + (int64_t)r.i2 * a.vs.s1;
  r.p0x = q0[0]; r.p0y = q0[a.vs.s2]; r.z0 = q0[2 * a.vs.s2];
  r.p1x = q1[0]; r.p1y = q1[a.vs.s2]; r.z1 = q1[2 * a.vs.s2];
  r.p2x = q2[0]; r.p2y = q2[a.vs.s2]; r.z2 = q2[2 * a.vs.s2];
}

// Dense layout fast path: v [N,V,3] and vi [.,F,3] rows contiguous -> 32-bit index arithmetic
// (the generic stride path costs ~45 % of render_bwd's instructions in 64-bit address math).
__device__ __forceinline__ void load_tri_dense(const int32_t* __restrict__ vin, const float* __restrict__ vn,
                                               int t, TriVerts& r) {
  const int32_t* vip = vin + (unsigned)t * 3u;
  r.i0 = vip[0]; r.i1 = vip[1]; r.i2 = vip[2];
  const float* q0 = vn + (unsigned)r.i0 * 3u;
  const float* q1 = vn + (unsigned)r.i1 * 3u;
  const float* q2 = vn + (unsigned)r.i2 * 3u;
  r.p0x = q0[0]; r.p0y = q0[1]; r.z0 = q0[2];
  r.p1x = q1[0]; r.p1y = q1[1]; r.z1 = q1[2];
  r.p2x = q2[0]; r.p2y = q2[1]; r.z2 = q2[2];
}

struct PixOut { float b0, b1, b2, depth; };

__device__ __forceinline__ PixOut shade(const TriVerts& t, float px, float py) {
  const float v01x = t.p1x - t.p0x, v01y = t.p1y - t.p0y;
  const float v02x = t.p2x - t.p0x, v02y = t.p2y - t.p0y;
  const float den = epsclamp(v01x * v02y - v01y * v02x);  // (:88)
  const float rden = rcp_approx(den);
  const float qx = px - t.p0x, qy = py - t.p0y;           // (:90)
  const float b1 = (qx * v02y - qy * v02x) * rden;        // (:92-96)
  const float b2 = (qy * v01x - qx * v01y) * rden;
  const float b0 = 1.f - b1 - b2;                          // (:97)
  const float d0 = rcp_approx(epsclamp(t.z0)), d1 = rcp_approx(epsclamp(t.z1)),
              d2 = rcp_approx(epsclamp(t.z2));              // (:99-100)
  const float dinv = d0 * b0 + d1 * b1 + d2 * b2;          // (:102)
  const float depth = rcp_approx(epsclamp(dinv));           // (:103)
  PixOut o;
  o.b0 = d0 * b0 * depth; o.b1 = d1 * b1 * depth; o.b2 = d2 * b2 * depth;  // (:105)
  o.depth = depth;
  return o;
}

// VEC: four pixels per thread, 128-bit accesses (requires W % 4 == 0 and a dense, aligned index_img)
template <bool VEC>
__global__ void __launch_bounds__(256) render_fwd_kernel(RenderArgs a, float* __restrict__ depth_img,
                                                         float* __restrict__ bary_img) {
  // blockIdx.y = image; 32-bit pixel arithmetic inside an image (host guarantees H*W < 2^31)
  const int HW = a.H * a.W;
  const int n = blockIdx.y;
  const int32_t* ibase = a.index_img + (int64_t)n * a.is.s0;
  float* bbase = bary_img + (int64_t)n * 3 * HW;
  float* dbase = depth_img + (int64_t)n * HW;
  if (VEC) {
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < HW / 4; q += gridDim.x * blockDim.x) {
      const int rem = q * 4;
      const int h = rem / a.W, w = rem - h * a.W;
      const int4 id = ldg_stream_i4(ibase + (int64_t)h * a.is.s1 + w);
      const int ids[4] = {id.x, id.y, id.z, id.w};
      float o0[4], o1[4], o2[4], od[4];
      TriVerts tv;
      int cached = -1;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (ids[j] != -1) {
          if (ids[j] != cached) { load_tri(a, n, ids[j], tv); cached = ids[j]; }
          const PixOut p = shade(tv, (float)(w + j), (float)h);
          o0[j] = p.b0; o1[j] = p.b1; o2[j] = p.b2; od[j] = p.depth;
        } else {
          o0[j] = 0.f; o1[j] = 0.f; o2[j] = 0.f; od[j] = 0.f;  // (:110-115)
        }
      }
      float* bp = bbase + rem;
      stg_stream_f4(bp, make_float4(o0[0], o0[1], o0[2], o0[3]));
      stg_stream_f4(bp + HW, make_float4(o1[0], o1[1], o1[2], o1[3]));
      stg_stream_f4(bp + 2 * (int64_t)HW, make_float4(o2[0], o2[1], o2[2], o2[3]));
      stg_stream_f4(dbase + rem, make_float4(od[0], od[1], od[2], od[3]));
    }
  } else {
    for (int rem = blockIdx.x * blockDim.x + threadIdx.x; rem < HW; rem += gridDim.x * blockDim.x) {
      const int h = rem / a.W, w = rem - h * a.W;
      const int id = ibase[(int64_t)h * a.is.s1 + (int64_t)w * a.is.s2];
      float* bp = bbase + rem;
      if (id != -1) {
        TriVerts tv;
        load_tri(a, n, id, tv);
        const PixOut p = shade(tv, (float)w, (float)h);
        bp[0] = p.b0; bp[HW] = p.b1; bp[2 * (int64_t)HW] = p.b2; dbase[rem] = p.depth;
      } else {
        bp[0] = 0.f; bp[HW] = 0.f; bp[2 * (int64_t)HW] = 0.f; dbase[rem] = 0.f;
      }
    }
  }
}

// Dense fast path of the forward (v rows and vi rows contiguous, index_img 128-bit accessible): 32-bit gather
// arithmetic, and the per-triangle part of `shade` (edge vectors, 1/den, 1/z_k: 4 MUFU + ~15 FP32) is evaluated
// once per distinct triangle of the thread's four pixels (runs are ~4 px on the 100k-triangle mesh) instead of
// once per pixel.
struct TriSetupFwd { float p0x, p0y, v01x, v01y, v02x, v02y, rden, d0, d1, d2; };

__device__ __forceinline__ void setup_fwd(const int32_t* __restrict__ vin, const float* __restrict__ vn, int t,
                                          TriSetupFwd& s) {
  TriVerts r;
  load_tri_dense(vin, vn, t, r);
  s.p0x = r.p0x; s.p0y = r.p0y;
  s.v01x = r.p1x - r.p0x; s.v01y = r.p1y - r.p0y;
  s.v02x = r.p2x - r.p0x; s.v02y = r.p2y - r.p0y;
  s.rden = rcp_approx(epsclamp(s.v01x * s.v02y - s.v01y * s.v02x));
  s.d0 = rcp_approx(epsclamp(r.z0)); s.d1 = rcp_approx(epsclamp(r.z1)); s.d2 = rcp_approx(epsclamp(r.z2));
}

__device__ __forceinline__ PixOut shade_fwd(const TriSetupFwd& s, float px, float py) {
  const float qx = px - s.p0x, qy = py - s.p0y;
  const float b1 = (qx * s.v02y - qy * s.v02x) * s.rden;
  const float b2 = (qy * s.v01x - qx * s.v01y) * s.rden;
  const float b0 = 1.f - b1 - b2;
  const float dinv = s.d0 * b0 + s.d1 * b1 + s.d2 * b2;
  const float depth = rcp_approx(epsclamp(dinv));
  PixOut o;
  o.b0 = s.d0 * b0 * depth; o.b1 = s.d1 * b1 * depth; o.b2 = s.d2 * b2 * depth;
  o.depth = depth;
  return o;
}

__global__ void __launch_bounds__(256) render_fwd_dense_kernel(RenderArgs a, float* __restrict__ depth_img,
                                                               float* __restrict__ bary_img) {
  const int HW = a.H * a.W;
  const int n = blockIdx.y;
  const int32_t* ibase = a.index_img + (int64_t)n * a.is.s0;
  const int32_t* vin = a.vi + (int64_t)n * a.vis.s0;
  const float* vn = a.v + (int64_t)n * a.vs.s0;
  float* bbase = bary_img + (int64_t)n * 3 * HW;
  float* dbase = depth_img + (int64_t)n * HW;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < HW / 4; q += gridDim.x * blockDim.x) {
    const int rem = q * 4;
    const int h = rem / a.W, w = rem - h * a.W;
    const int4 id = ldg_stream_i4(ibase + (int64_t)h * a.is.s1 + w);
    const int ids[4] = {id.x, id.y, id.z, id.w};
    float o0[4], o1[4], o2[4], od[4];
    TriSetupFwd ts;
    int cached = -1;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (ids[j] != -1) {
        if (ids[j] != cached) { setup_fwd(vin, vn, ids[j], ts); cached = ids[j]; }
        const PixOut p = shade_fwd(ts, (float)(w + j), (float)h);
        o0[j] = p.b0; o1[j] = p.b1; o2[j] = p.b2; od[j] = p.depth;
      } else {
        o0[j] = 0.f; o1[j] = 0.f; o2[j] = 0.f; od[j] = 0.f;  // (:110-115)
      }
    }
    float* bp = bbase + rem;
    stg_stream_f4(bp, make_float4(o0[0], o0[1], o0[2], o0[3]));
    stg_stream_f4(bp + HW, make_float4(o1[0], o1[1], o1[2], o1[3]));
    stg_stream_f4(bp + 2 * (int64_t)HW, make_float4(o2[0], o2[1], o2[2], o2[3]));
    stg_stream_f4(dbase + rem, make_float4(od[0], od[1], od[2], od[3]));
  }
}

struct RenderBwdArgs {
  RenderArgs r;
  const float* grad_depth;  // may be null
  Strides3 gds;
  const float* grad_bary;   // may be null
  Strides4 gbs;
};

// One thread per pixel; a warp covers 32 consecutive pixels of one image row segment.
// DENSE: every tensor contiguous (the common case) -> offsets are n*HW + rem with 32-bit arithmetic.
template <bool DENSE>
__global__ void __launch_bounds__(256) render_bwd_kernel(RenderBwdArgs b, float* __restrict__ grad_v) {
  const RenderArgs& a = b.r;
  const int HW = a.H * a.W;  // blockIdx.y = image, 32-bit pixel arithmetic inside it
  const int lane = threadIdx.x & 31;
  const int rem = blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_range = rem < HW;
  const int n = blockIdx.y;
  int h = 0, w = 0, id = -1;
  if (in_range) {
    h = rem / a.W; w = rem - h * a.W;
    id = DENSE ? a.index_img[(int64_t)n * HW + rem]
               : a.index_img[(int64_t)n * a.is.s0 + (int64_t)h * a.is.s1 + (int64_t)w * a.is.s2];
  }
  // key of a run: the triangle id (a block never spans two images); -1 lanes get unique keys
  const int key = (id == -1) ? (-2 - lane) : id;
  if (__all_sync(0xffffffffu, id == -1)) return;

  float g[9];  // dL/d(p0.x, p0.y, z0, p1.x, p1.y, z1, p2.x, p2.y, z2)
#pragma unroll
  for (int i = 0; i < 9; ++i) g[i] = 0.f;
  TriVerts t;
  t.i0 = t.i1 = t.i2 = 0;
  if (id != -1) {
    if (DENSE) load_tri_dense(a.vi + (int64_t)n * a.vis.s0, a.v + (int64_t)n * a.V * 3, id, t);
    else load_tri(a, n, id, t);
    const float v01x = t.p1x - t.p0x, v01y = t.p1y - t.p0y;
    const float v02x = t.p2x - t.p0x, v02y = t.p2y - t.p0y;
    const float den_raw = v01x * v02y - v01y * v02x;
    const float den = epsclamp(den_raw);
    const bool den_clamped = den != den_raw;  // (:198)
    const float rden = rcp_approx(den);
    const float qx = (float)w - t.p0x, qy = (float)h - t.p0y;
    const float b1 = (qx * v02y - qy * v02x) * rden;
    const float b2 = (qy * v01x - qx * v01y) * rden;
    const float b0 = 1.f - b1 - b2;
    const float z0e = epsclamp(t.z0), z1e = epsclamp(t.z1), z2e = epsclamp(t.z2);
    const bool c0 = z0e != t.z0, c1 = z1e != t.z1, c2 = z2e != t.z2;  // (:211-213)
    const float d0 = rcp_approx(z0e), d1 = rcp_approx(z1e), d2 = rcp_approx(z2e);
    const float dinv = d0 * b0 + d1 * b1 + d2 * b2;
    const float dinv_e = epsclamp(dinv);
    const bool dinv_clamped = dinv_e != dinv;  // (:219)
    const float depth = rcp_approx(dinv_e);

    float g0 = 0.f, g1 = 0.f, g2 = 0.f, gd = 0.f;
    if (b.grad_bary) {
      if (DENSE) {
        const float* gp = b.grad_bary + (int64_t)n * 3 * HW + rem;
        g0 = ldg_stream_f(gp); g1 = ldg_stream_f(gp + HW); g2 = ldg_stream_f(gp + 2 * (int64_t)HW);
      } else {
        const float* gp = b.grad_bary + (int64_t)n * b.gbs.s0 + (int64_t)h * b.gbs.s2 + (int64_t)w * b.gbs.s3;
        g0 = ldg_stream_f(gp); g1 = ldg_stream_f(gp + b.gbs.s1); g2 = ldg_stream_f(gp + 2 * b.gbs.s1);
      }
    }
    if (b.grad_depth)
      gd = DENSE ? ldg_stream_f(b.grad_depth + (int64_t)n * HW + rem)
                 : ldg_stream_f(b.grad_depth + (int64_t)n * b.gds.s0 + (int64_t)h * b.gds.s1 + (int64_t)w * b.gds.s2);

    const float dL_depth = gd + (g0 * d0 * b0 + g1 * d1 * b1 + g2 * d2 * b2);               // (:226)
    const float dL_dinv = dinv_clamped ? 0.f : (-dL_depth * rcp_approx(dinv * dinv));        // (:228-229)
    const float dLd0 = g0 * b0 * depth + dL_dinv * b0;                                       // (:230)
    const float dLd1 = g1 * b1 * depth + dL_dinv * b1;
    const float dLd2 = g2 * b2 * depth + dL_dinv * b2;
    g[2] = c0 ? 0.f : -dLd0 * rcp_approx(z0e * z0e);                                          // (:231-250)
    g[5] = c1 ? 0.f : -dLd1 * rcp_approx(z1e * z1e);
    g[8] = c2 ? 0.f : -dLd2 * rcp_approx(z2e * z2e);
    const float dLb0 = g0 * d0 * depth + dL_dinv * d0;                                       // (:252)
    const float dLb1 = g1 * d1 * depth + dL_dinv * d1;
    const float dLb2 = g2 * d2 * depth + dL_dinv * d2;
    const float e1 = (-dLb0 + dLb1) * rden, e2 = (-dLb0 + dLb2) * rden;                      // (:253-254)
    const float dL_den = den_clamped ? 0.f : -(e1 * b1 + e2 * b2);                           // (:256)
    const float dqx = e1 * v02y - e2 * v01y, dqy = -e1 * v02x + e2 * v01x;                   // (:258-260)
    const float dv02x = -e1 * qy - dL_den * v01y, dv02y = e1 * qx + dL_den * v01x;           // (:262-264)
    const float dv01x = e2 * qy + dL_den * v02y, dv01y = -e2 * qx - dL_den * v02x;           // (:265-267)
    g[0] = -dv02x - dv01x - dqx; g[1] = -dv02y - dv01y - dqy;                                // (:269)
    g[3] = dv01x; g[4] = dv01y; g[6] = dv02x; g[7] = dv02y;                                  // (:270-271)
  }

  // segmented reduction over runs of equal (image, triangle)
  const int key_up = __shfl_up_sync(0xffffffffu, key, 1);
  const int key_dn = __shfl_down_sync(0xffffffffu, key, 1);
  const bool head = (lane == 0) || (key_up != key);
  const bool tail = (lane == 31) || (key_dn != key);
  const unsigned tail_mask = __ballot_sync(0xffffffffu, tail);
  seg_reduce_to_head<9>(g, tail_mask, lane);
  if (head && id != -1) {
    float* gv = grad_v + (int64_t)n * a.V * 3;
    float* q0 = gv + (int64_t)t.i0 * 3;
    float* q1 = gv + (int64_t)t.i1 * 3;
    float* q2 = gv + (int64_t)t.i2 * 3;
    red_add(q0 + 0, g[0]); red_add(q0 + 1, g[1]); red_add(q0 + 2, g[2]);
    red_add(q1 + 0, g[3]); red_add(q1 + 1, g[4]); red_add(q1 + 2, g[5]);
    red_add(q2 + 0, g[6]); red_add(q2 + 1, g[7]); red_add(q2 + 2, g[8]);
  }
}

// Walker variant of the backward (dense tensors, W % 8 == 0, 16-B aligned planes).
//   1. tri_table_kernel: one thread per (image, triangle) derives the per-triangle constants of the backward
//      once and stores them as one 64-B row (4 x 128-bit), so the hot kernel replaces the two-level dependent
//      gather index -> vi (3 x 4 B) -> v (9 x 4 B) + ~40 setup instructions by four independent LDG.128.
//   2. render_bwd_walk_kernel: one thread owns EIGHT consecutive pixels of a row, streams them with 128-bit
//      loads and walks them in order; the nine partial derivatives of a run of equal ids are accumulated in
//      registers and flushed when the id changes with THREE 128-bit reductions into a [N,V,4]-padded
//      accumulator (rows 16-B aligned; red.global.add.v4.f32) -- no shuffles, 3 instead of 9 REDs per run.
//   3. unpad_kernel: [N,V,4] -> grad_v [N,V,3].
// The per-pixel kernel above spends ~310 thread-instructions per pixel, two thirds of them in the segmented
// shuffle reduction and in re-deriving the triangle for every pixel.
constexpr int kWalkPx = 8;
#ifndef DRTK_RENDER_BWD_CHUNKS
#define DRTK_RENDER_BWD_CHUNKS 1  // 8-pixel chunks per thread; measured on B200: 1 -> 0.243 ms, 2 -> 0.271 ms, 4 -> 0.349 ms (config 4)
#endif

struct RunSetup {  // per-triangle constants of the backward (:186-219 of the reference)
  float p0x, p0y, v01x, v01y, v02x, v02y, rden, d0, d1, d2, rz0, rz1, rz2;  // rzk = 1/zk^2 (0 when zk was clamped)
  bool den_clamped;
  int i0, i1, i2;
};

// table row: {p0x,p0y,v01x,v01y} {v02x,v02y,rden,d0} {d1,d2,rz1,rz2} {rz0 | sign bit = den_clamped, i0,i1,i2}
// Also zero-fills the padded gradient accumulator the walker reduces into (saves a memset launch per step).
__global__ void __launch_bounds__(256) tri_table_kernel(RenderArgs a, float4* __restrict__ table,
                                                        float4* __restrict__ zero, int64_t zero_count) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  {
    const int64_t gid = ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * gridDim.y * blockDim.x;
    for (int64_t i = gid; i < zero_count; i += nthreads) zero[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (f >= a.F) return;
  TriVerts t;
  load_tri_dense(a.vi + (int64_t)n * a.vis.s0, a.v + (int64_t)n * a.V * 3, f, t);
  const float v01x = t.p1x - t.p0x, v01y = t.p1y - t.p0y;
  const float v02x = t.p2x - t.p0x, v02y = t.p2y - t.p0y;
  const float den_raw = v01x * v02y - v01y * v02x;
  const float den = epsclamp(den_raw);
  const float z0e = epsclamp(t.z0), z1e = epsclamp(t.z1), z2e = epsclamp(t.z2);
  const float rz0 = (z0e != t.z0) ? 0.f : rcp_approx(z0e * z0e);
  const float rz1 = (z1e != t.z1) ? 0.f : rcp_approx(z1e * z1e);
  const float rz2 = (z2e != t.z2) ? 0.f : rcp_approx(z2e * z2e);
  float4* row = table + ((int64_t)n * a.F + f) * 4;
  row[0] = make_float4(t.p0x, t.p0y, v01x, v01y);
  row[1] = make_float4(v02x, v02y, rcp_approx(den), rcp_approx(z0e));
  row[2] = make_float4(rcp_approx(z1e), rcp_approx(z2e), rz1, rz2);
  row[3] = make_float4(__uint_as_float(__float_as_uint(rz0) | (den != den_raw ? 0x80000000u : 0u)),
                       __int_as_float(t.i0), __int_as_float(t.i1), __int_as_float(t.i2));
}

__device__ __forceinline__ void run_setup(const float4* __restrict__ row, RunSetup& r) {
  // one 64-B table row = two 256-bit loads (LDG.E.ENL2.256): a divergent gather costs the L1 data pipe one
  // wavefront per lane and instruction, so halving the instructions halves that cost
  const float8 lo = ldg_f8(reinterpret_cast<const float*>(row)), hi = ldg_f8(reinterpret_cast<const float*>(row + 2));
  const float4 a = lo.lo, b = lo.hi, c = hi.lo, d = hi.hi;
  r.p0x = a.x; r.p0y = a.y; r.v01x = a.z; r.v01y = a.w;
  r.v02x = b.x; r.v02y = b.y; r.rden = b.z; r.d0 = b.w;
  r.d1 = c.x; r.d2 = c.y; r.rz1 = c.z; r.rz2 = c.w;
  r.rz0 = fabsf(d.x); r.den_clamped = (__float_as_uint(d.x) >> 31) != 0u;
  r.i0 = __float_as_int(d.y); r.i1 = __float_as_int(d.z); r.i2 = __float_as_int(d.w);
}

// Measured and NOT adopted in round 2 (config 4, B200, profiles/r02_opbench_render_bwd_*.txt; baseline 0.244-0.252 ms):
//   * merging the runs cut at a thread's 8-pixel boundary with the neighbouring thread's run (first run parked in
//     shared memory, added by the left neighbour): 28 % fewer reductions, but 0.289 ms with a CTA barrier and
//     0.302 ms warp-synchronous -- the kernel is not bound by the number of REDs;
//   * vertex carry (the next triangle of a row shares an edge: keep the shared vertices' sums, reduce one row instead
//     of three per run): 0.291 ms, the 9 compares + 27 selects per run boundary cost more than the REDs they save;
//   * L1 prefetch (CCTL.PF1) of the later runs' table rows at the head of the thread: 0.271 ms;
//   * 7 / 8 CTAs per SM through launch bounds (spills): 0.270 / 0.273 ms.
// CHUNKS: consecutive 8-pixel chunks walked by one thread with the run state carried from chunk to chunk (a run cut
// at a thread boundary costs an extra table fetch and an extra flush: 8 px per thread = 3 runs per 8 px on the
// 100k-triangle mesh, 16 px per thread = 5 runs per 16 px).
template <bool HAS_GB, bool HAS_GD, int CHUNKS>
__global__ void __launch_bounds__(128) render_bwd_walk_kernel(RenderBwdArgs b, const float4* __restrict__ table,
                                                              float* __restrict__ gpad) {
  const RenderArgs& a = b.r;
  const int HW = a.H * a.W;
  const int n = blockIdx.y;
  const int rem0 = (blockIdx.x * blockDim.x + threadIdx.x) * (kWalkPx * CHUNKS);
  if (rem0 >= HW) return;
  const float4* tn = table + (int64_t)n * a.F * 4;
  float* gvn = gpad + (int64_t)n * a.V * 4;

  RunSetup r;
  int cur = -1;
  float acc[9];
  auto flush = [&]() {
    red_add_v4(gvn + (int64_t)r.i0 * 4, acc[0], acc[1], acc[2], 0.f);
    red_add_v4(gvn + (int64_t)r.i1 * 4, acc[3], acc[4], acc[5], 0.f);
    red_add_v4(gvn + (int64_t)r.i2 * 4, acc[6], acc[7], acc[8], 0.f);
  };
#pragma unroll 1
  for (int ch = 0; ch < CHUNKS; ++ch) {
    const int rem = rem0 + ch * kWalkPx;
    if (rem >= HW) break;
    const int32_t* ip = a.index_img + (int64_t)n * HW + rem;
    const int4 ia = ldg_stream_i4(ip), ib = ldg_stream_i4(ip + 4);
    const int ids[kWalkPx] = {ia.x, ia.y, ia.z, ia.w, ib.x, ib.y, ib.z, ib.w};
    if ((ia.x & ia.y & ia.z & ia.w & ib.x & ib.y & ib.z & ib.w) == -1) {  // all empty: a run cannot continue through it
      if (cur != -1) { flush(); cur = -1; }
      continue;
    }
    float gb0[kWalkPx], gb1[kWalkPx], gb2[kWalkPx], gdp[kWalkPx];
    if (HAS_GB) {
      const float* gp = b.grad_bary + (int64_t)n * 3 * HW + rem;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const float4 x = ldg_stream_f4(gp + 4 * q), y = ldg_stream_f4(gp + HW + 4 * q),
                     z = ldg_stream_f4(gp + 2 * (int64_t)HW + 4 * q);
        gb0[4 * q] = x.x; gb0[4 * q + 1] = x.y; gb0[4 * q + 2] = x.z; gb0[4 * q + 3] = x.w;
        gb1[4 * q] = y.x; gb1[4 * q + 1] = y.y; gb1[4 * q + 2] = y.z; gb1[4 * q + 3] = y.w;
        gb2[4 * q] = z.x; gb2[4 * q + 1] = z.y; gb2[4 * q + 2] = z.z; gb2[4 * q + 3] = z.w;
      }
    }
    if (HAS_GD) {
      const float* gp = b.grad_depth + (int64_t)n * HW + rem;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const float4 x = ldg_stream_f4(gp + 4 * q);
        gdp[4 * q] = x.x; gdp[4 * q + 1] = x.y; gdp[4 * q + 2] = x.z; gdp[4 * q + 3] = x.w;
      }
    }
    const int h = rem / a.W, w0 = rem - h * a.W;  // W % 8 == 0: the eight pixels share the row
#pragma unroll
    for (int j = 0; j < kWalkPx; ++j) {
      const int id = ids[j];
      if (id == -1) continue;
      if (id != cur) {
        if (cur != -1) flush();
        run_setup(tn + (int64_t)id * 4, r);
        cur = id;
#pragma unroll
        for (int i = 0; i < 9; ++i) acc[i] = 0.f;
      }
      const float qx = (float)(w0 + j) - r.p0x, qy = (float)h - r.p0y;
      const float b1 = (qx * r.v02y - qy * r.v02x) * r.rden;
      const float b2 = (qy * r.v01x - qx * r.v01y) * r.rden;
      const float b0 = 1.f - b1 - b2;
      const float dinv = r.d0 * b0 + r.d1 * b1 + r.d2 * b2;
      const float dinv_e = epsclamp(dinv);
      const float depth = rcp_approx(dinv_e);
      const float g0 = HAS_GB ? gb0[j] : 0.f, g1 = HAS_GB ? gb1[j] : 0.f, g2 = HAS_GB ? gb2[j] : 0.f;
      const float gd = HAS_GD ? gdp[j] : 0.f;
      const float dL_depth = gd + (g0 * r.d0 * b0 + g1 * r.d1 * b1 + g2 * r.d2 * b2);
      const float dL_dinv = (dinv_e != dinv) ? 0.f : (-dL_depth * rcp_approx(dinv * dinv));
      const float dLd0 = g0 * b0 * depth + dL_dinv * b0;
      const float dLd1 = g1 * b1 * depth + dL_dinv * b1;
      const float dLd2 = g2 * b2 * depth + dL_dinv * b2;
      acc[2] += -dLd0 * r.rz0; acc[5] += -dLd1 * r.rz1; acc[8] += -dLd2 * r.rz2;
      const float dLb0 = g0 * r.d0 * depth + dL_dinv * r.d0;
      const float dLb1 = g1 * r.d1 * depth + dL_dinv * r.d1;
      const float dLb2 = g2 * r.d2 * depth + dL_dinv * r.d2;
      const float e1 = (-dLb0 + dLb1) * r.rden, e2 = (-dLb0 + dLb2) * r.rden;
      const float dL_den = r.den_clamped ? 0.f : -(e1 * b1 + e2 * b2);
      const float dqx = e1 * r.v02y - e2 * r.v01y, dqy = -e1 * r.v02x + e2 * r.v01x;
      const float dv02x = -e1 * qy - dL_den * r.v01y, dv02y = e1 * qx + dL_den * r.v01x;
      const float dv01x = e2 * qy + dL_den * r.v02y, dv01y = -e2 * qx - dL_den * r.v02x;
      acc[0] += -dv02x - dv01x - dqx; acc[1] += -dv02y - dv01y - dqy;
      acc[3] += dv01x; acc[4] += dv01y; acc[6] += dv02x; acc[7] += dv02y;
    }
  }
  if (cur != -1) flush();
}

__global__ void __launch_bounds__(256) unpad_kernel(const float4* __restrict__ gpad, float* __restrict__ grad_v, int64_t rows) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  const float4 g = gpad[i];
  float* o = grad_v + i * 3;
  o[0] = g.x; o[1] = g.y; o[2] = g.z;
}

// workspace of the walker path: the triangle table, then the padded accumulators
inline size_t walk_table_bytes(int64_t N, int64_t F) { return (size_t)(N * F) * 64; }
inline size_t walk_gpad_bytes(int64_t N, int64_t V) { return (size_t)(N * V) * 16; }

inline unsigned grid_for(int64_t work_items, int threads, int ctas_per_sm) {
  const int64_t need = (work_items + threads - 1) / threads;
  const int64_t cap = (int64_t)num_sms() * ctas_per_sm;
  return (unsigned)(need < cap ? (need > 0 ? need : 1) : cap);
}

}  // namespace
}  // namespace drtk

using namespace drtk;

extern "C" int drtk_b200_render_forward(const float* v, const int64_t* v_strides, const int32_t* vi,
                                        const int64_t* vi_strides, const int32_t* index_img,
                                        const int64_t* index_strides, int64_t N, int64_t V, int64_t F,
                                        int64_t H, int64_t W, float* depth_img, float* bary_img,
                                        void* stream_) {
  if (N < 0 || H < 0 || W < 0) return DRTK_B200_EINVAL;
  if (N * H * W == 0) return 0;
  // an empty face list (F == 0) may come with null v / vi pointers: no pixel can reference a triangle then
  if (((!v || !vi) && F > 0) || !index_img || !depth_img || !bary_img) return DRTK_B200_EINVAL;
  if (H > (1 << 30) || W > (1 << 30) || N > (1 << 30)) return DRTK_B200_EUNSUPPORTED;
  if (N > kMaxBatchPerLaunch) {  // batch index rides on gridDim.y: slices of the batch
    for (int64_t n0 = 0; n0 < N; n0 += kMaxBatchPerLaunch) {
      const int64_t nn = (N - n0 < kMaxBatchPerLaunch) ? N - n0 : kMaxBatchPerLaunch;
      const int rc = drtk_b200_render_forward(v ? v + n0 * v_strides[0] : nullptr, v_strides, vi ? vi + n0 * vi_strides[0] : nullptr, vi_strides,
                                              index_img + n0 * index_strides[0], index_strides, nn, V, F, H, W,
                                              depth_img + n0 * H * W, bary_img + n0 * 3 * H * W, stream_);
      if (rc) return rc;
    }
    return 0;
  }
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RenderArgs a;
  a.v = v; a.vs = make3(v_strides); a.vi = vi; a.vis = make3(vi_strides);
  a.index_img = index_img; a.is = make3(index_strides);
  a.N = (int)N; a.V = (int)V; a.F = (int)F; a.H = (int)H; a.W = (int)W;
  const bool vec = VecOk::image(index_img, W, a.is.s2, a.is.s1, a.is.s0);
  if (H * W >= (int64_t)0x7FFFFFF0) return DRTK_B200_EUNSUPPORTED;
  // grid.y = image; grid.x sized so that all CTAs of all images are co-resident (one wave, grid-stride loops
  // inside): a partly filled last wave cost 8 % in interp_fwd_kernel
  const int64_t items = vec ? H * W / 4 : H * W;
  auto launch = [&](auto kern) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, 0) != cudaSuccess || occ < 1) occ = 1;
    int64_t gx = (int64_t)num_sms() * occ / N;
    const int64_t need = (items + 255) / 256;
    if (gx < 1) gx = 1;
    if (gx > need) gx = need;
    kern<<<dim3((unsigned)gx, (unsigned)N), 256, 0, stream>>>(a, depth_img, bary_img);
  };
  const bool dense_tables = a.vs.s2 == 1 && a.vs.s1 == 3 && a.vis.s2 == 1 && a.vis.s1 == 3 && V * 3 < INT32_MAX &&
                            F * 3 < INT32_MAX;
  if (vec && dense_tables) launch(render_fwd_dense_kernel);
  else if (vec) launch(render_fwd_kernel<true>);
  else launch(render_fwd_kernel<false>);
  DRTK_CHECK_LAUNCH();
  return 0;
}

extern "C" size_t drtk_b200_render_backward_workspace_bytes(int64_t N, int64_t V, int64_t F) {
  if (N <= 0 || V <= 0 || F <= 0) return 0;
  return walk_table_bytes(N, F) + walk_gpad_bytes(N, V) + 32;  // + slack to align the table to 32 B (256-bit loads)
}

extern "C" int drtk_b200_render_backward(const float* v, const int64_t* v_strides, const int32_t* vi,
                                         const int64_t* vi_strides, const int32_t* index_img,
                                         const int64_t* index_strides, const float* grad_depth,
                                         const int64_t* grad_depth_strides, const float* grad_bary,
                                         const int64_t* grad_bary_strides, int64_t N, int64_t V,
                                         int64_t F, int64_t H, int64_t W, float* grad_v, void* workspace,
                                         size_t workspace_bytes, void* stream_) {
  if (N < 0 || V < 0 || H < 0 || W < 0) return DRTK_B200_EINVAL;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (N * V > 0 && !grad_v) return DRTK_B200_EINVAL;
  const int64_t npix = N * H * W;
  const bool trivial = npix == 0 || N * V == 0 || F == 0 || (!grad_depth && !grad_bary);  // zero grad_v
  const auto al16 = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
  if (trivial) {
    if (N * V > 0) DRTK_CUDA(cudaMemsetAsync(grad_v, 0, sizeof(float) * (size_t)(N * V * 3), stream));  // (:397)
    return 0;
  }
  if (!v || !vi || !index_img) return DRTK_B200_EINVAL;
  if (H > (1 << 30) || W > (1 << 30) || N > (1 << 30)) return DRTK_B200_EUNSUPPORTED;
  if (N > kMaxBatchPerLaunch) {  // batch index rides on gridDim.y: slices of the batch (the workspace is reused)
    for (int64_t n0 = 0; n0 < N; n0 += kMaxBatchPerLaunch) {
      const int64_t nn = (N - n0 < kMaxBatchPerLaunch) ? N - n0 : kMaxBatchPerLaunch;
      const int rc = drtk_b200_render_backward(
          v + n0 * v_strides[0], v_strides, vi + n0 * vi_strides[0], vi_strides, index_img + n0 * index_strides[0],
          index_strides, grad_depth ? grad_depth + n0 * grad_depth_strides[0] : nullptr, grad_depth_strides,
          grad_bary ? grad_bary + n0 * grad_bary_strides[0] : nullptr, grad_bary_strides, nn, V, F, H, W,
          grad_v + n0 * V * 3, workspace, workspace_bytes, stream_);
      if (rc) return rc;
    }
    return 0;
  }
  RenderBwdArgs b;
  b.r.v = v; b.r.vs = make3(v_strides); b.r.vi = vi; b.r.vis = make3(vi_strides);
  b.r.index_img = index_img; b.r.is = make3(index_strides);
  b.r.N = (int)N; b.r.V = (int)V; b.r.F = (int)F; b.r.H = (int)H; b.r.W = (int)W;
  b.grad_depth = grad_depth; b.gds = grad_depth ? make3(grad_depth_strides) : Strides3{0, 0, 0};
  b.grad_bary = grad_bary; b.gbs = grad_bary ? make4(grad_bary_strides) : Strides4{0, 0, 0, 0};
  if (H * W >= (int64_t)0x7FFFFFF0) return DRTK_B200_EUNSUPPORTED;
  auto dense3 = [&](const Strides3& s, int64_t d1, int64_t d2) { return s.s2 == 1 && s.s1 == d2 && (N == 1 || s.s0 == d1 * d2); };
  const bool dense = dense3(b.r.is, H, W) && b.r.vs.s2 == 1 && b.r.vs.s1 == 3 && (N == 1 || b.r.vs.s0 == V * 3) &&
                     b.r.vis.s2 == 1 && b.r.vis.s1 == 3 && V * 3 < (int64_t)0x7FFFFFF0 && F * 3 < (int64_t)0x7FFFFFF0 &&
                     (!grad_depth || dense3(b.gds, H, W)) &&
                     (!grad_bary || (b.gbs.s3 == 1 && b.gbs.s2 == W && b.gbs.s1 == H * W && (N == 1 || b.gbs.s0 == 3 * H * W)));
  if (dense && W % kWalkPx == 0 && al16(index_img) && (!grad_bary || al16(grad_bary)) && (!grad_depth || al16(grad_depth)) &&
      F <= 65535LL * 256) {
    const size_t tb = walk_table_bytes(N, F), gb = walk_gpad_bytes(N, V);
    if (!workspace || workspace_bytes < tb + gb + 32) return DRTK_B200_EWORKSPACE;
    char* ws32 = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 31) & ~uintptr_t(31));
    float4* table = reinterpret_cast<float4*>(ws32);
    float* gpad = reinterpret_cast<float*>(ws32 + tb);
    tri_table_kernel<<<dim3((unsigned)((F + 255) / 256), (unsigned)N), 256, 0, stream>>>(
        b.r, table, reinterpret_cast<float4*>(gpad), (int64_t)(gb / sizeof(float4)));
    constexpr int CH = DRTK_RENDER_BWD_CHUNKS;
    const dim3 wgrid((unsigned)((H * W / (kWalkPx * CH) + 127) / 128 + 1), (unsigned)N);
    if (grad_bary && grad_depth) render_bwd_walk_kernel<true, true, CH><<<wgrid, 128, 0, stream>>>(b, table, gpad);
    else if (grad_bary) render_bwd_walk_kernel<true, false, CH><<<wgrid, 128, 0, stream>>>(b, table, gpad);
    else render_bwd_walk_kernel<false, true, CH><<<wgrid, 128, 0, stream>>>(b, table, gpad);
    unpad_kernel<<<(unsigned)((N * V + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const float4*>(gpad), grad_v, N * V);
    DRTK_CHECK_LAUNCH();
    return 0;
  }
  DRTK_CUDA(cudaMemsetAsync(grad_v, 0, sizeof(float) * (size_t)(N * V * 3), stream));  // (:397)
  const dim3 grid((unsigned)((H * W + 255) / 256), (unsigned)N);
  if (dense) render_bwd_kernel<true><<<grid, 256, 0, stream>>>(b, grad_v);
  else render_bwd_kernel<false><<<grid, 256, 0, stream>>>(b, grad_v);
  DRTK_CHECK_LAUNCH();
  return 0;
}
