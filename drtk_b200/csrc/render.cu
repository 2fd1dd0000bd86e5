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

inline unsigned grid_for(int64_t work_items, int threads, int ctas_per_sm) {
  const int64_t need = (work_items + threads - 1) / threads;
  const int64_t cap = (int64_t)kNumSMs * ctas_per_sm;
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
  if (!v || !vi || !index_img || !depth_img || !bary_img) return DRTK_B200_EINVAL;
  if (H > (1 << 30) || W > (1 << 30) || N > (1 << 30)) return DRTK_B200_EUNSUPPORTED;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RenderArgs a;
  a.v = v; a.vs = make3(v_strides); a.vi = vi; a.vis = make3(vi_strides);
  a.index_img = index_img; a.is = make3(index_strides);
  a.N = (int)N; a.V = (int)V; a.F = (int)F; a.H = (int)H; a.W = (int)W;
  const bool vec = VecOk::image(index_img, W, a.is.s2, a.is.s1, a.is.s0);
  if (H * W >= (int64_t)0x7FFFFFF0 || N > 65535) return DRTK_B200_EUNSUPPORTED;
  // grid.y = image; grid.x sized so that the whole grid is ~8 CTAs of 256 threads per SM
  const int64_t items = vec ? H * W / 4 : H * W;
  const unsigned gx = grid_for(items, 256, (int)((8 + N - 1) / N > 0 ? (8 + N - 1) / N : 1));
  if (vec) render_fwd_kernel<true><<<dim3(gx, (unsigned)N), 256, 0, stream>>>(a, depth_img, bary_img);
  else render_fwd_kernel<false><<<dim3(gx, (unsigned)N), 256, 0, stream>>>(a, depth_img, bary_img);
  DRTK_CHECK_LAUNCH();
  return 0;
}

extern "C" int drtk_b200_render_backward(const float* v, const int64_t* v_strides, const int32_t* vi,
                                         const int64_t* vi_strides, const int32_t* index_img,
                                         const int64_t* index_strides, const float* grad_depth,
                                         const int64_t* grad_depth_strides, const float* grad_bary,
                                         const int64_t* grad_bary_strides, int64_t N, int64_t V,
                                         int64_t F, int64_t H, int64_t W, float* grad_v, void* stream_) {
  if (N < 0 || V < 0 || H < 0 || W < 0) return DRTK_B200_EINVAL;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (N * V > 0) {
    if (!grad_v) return DRTK_B200_EINVAL;
    DRTK_CUDA(cudaMemsetAsync(grad_v, 0, sizeof(float) * (size_t)(N * V * 3), stream));  // (:397)
  }
  const int64_t npix = N * H * W;
  if (npix == 0 || N * V == 0) return 0;
  if (!v || !vi || !index_img) return DRTK_B200_EINVAL;
  if (!grad_depth && !grad_bary) return 0;  // all-zero upstream gradient -> zero grad_v
  if (H > (1 << 30) || W > (1 << 30) || N > (1 << 30)) return DRTK_B200_EUNSUPPORTED;
  RenderBwdArgs b;
  b.r.v = v; b.r.vs = make3(v_strides); b.r.vi = vi; b.r.vis = make3(vi_strides);
  b.r.index_img = index_img; b.r.is = make3(index_strides);
  b.r.N = (int)N; b.r.V = (int)V; b.r.F = (int)F; b.r.H = (int)H; b.r.W = (int)W;
  b.grad_depth = grad_depth; b.gds = grad_depth ? make3(grad_depth_strides) : Strides3{0, 0, 0};
  b.grad_bary = grad_bary; b.gbs = grad_bary ? make4(grad_bary_strides) : Strides4{0, 0, 0, 0};
  if (H * W >= (int64_t)0x7FFFFFF0 || N > 65535) return DRTK_B200_EUNSUPPORTED;
  auto dense3 = [&](const Strides3& s, int64_t d1, int64_t d2) { return s.s2 == 1 && s.s1 == d2 && (N == 1 || s.s0 == d1 * d2); };
  const bool dense = dense3(b.r.is, H, W) && b.r.vs.s2 == 1 && b.r.vs.s1 == 3 && (N == 1 || b.r.vs.s0 == V * 3) &&
                     b.r.vis.s2 == 1 && b.r.vis.s1 == 3 && V * 3 < (int64_t)0x7FFFFFF0 && F * 3 < (int64_t)0x7FFFFFF0 &&
                     (!grad_depth || dense3(b.gds, H, W)) &&
                     (!grad_bary || (b.gbs.s3 == 1 && b.gbs.s2 == W && b.gbs.s1 == H * W && (N == 1 || b.gbs.s0 == 3 * H * W)));
  const dim3 grid((unsigned)((H * W + 255) / 256), (unsigned)N);
  if (dense) render_bwd_kernel<true><<<grid, 256, 0, stream>>>(b, grad_v);
  else render_bwd_kernel<false><<<grid, 256, 0, stream>>>(b, grad_v);
  DRTK_CHECK_LAUNCH();
  return 0;
}
