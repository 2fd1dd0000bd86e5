// edge_grad.cu -- backward of edge_grad_estimator: image-space gradients at visibility
// discontinuities -> per-pixel dL/d(v_pix_img) [N,3,H,W].
//
// Semantics: src/edge_grad/edge_grad_kernel.cu:217-449 (helpers :18-215) of the reference
// (paper: "Rasterized Edge Gradients: Handling Discontinuities Differentiably").
//
// The reference is a SCATTER: every pixel (x < W-1, y < H-1) acts as "centre", classifies the
// (centre,right) and (centre,down) pairs and atomically adds 9 values into a zero-initialised
// output (12 B/px memset + 9 REDG per pixel, interior pixels included), and reads img/grad for
// every pair that shows two different triangles.  Here a CTA owns a 64x8 output tile: pairs
// showing different triangles are compacted into a job list, each job is evaluated once by one
// thread, its contributions go to per-role shared-memory slots (single writer each) and every
// pixel sums its slots and writes its three outputs exactly once.  No memset, no atomics,
// deterministic; img/grad_output are read only for pairs that actually contribute (adjacent
// triangles of a watertight mesh -- the bulk of all pairs -- do not).
//
// The discrete inside/outside tests use the reference's compiled arithmetic
// (x*y - z*w  ==  FFMA(x, y, -FMUL(z, w)), FTZ) so that classification agrees sample for sample.
#include "common.cuh"

namespace drtk {
namespace {

struct EdgeArgs {
  const float* v;
  Strides3 vs;
  const float* img;
  Strides4 ims;
  const int32_t* index_img;
  Strides3 is;
  const int32_t* vi;
  Strides3 vis;
  const float* go;
  Strides4 gs;
  int N, V, F, C, H, W;
  float max_dp_dr;
};

struct Tri2 {  // screen-space part of a triangle (get_tri_info, :72-87)
  int i0, i1, i2;
  float p0x, p0y, p1x, p1y, v01x, v01y, v02x, v02y, v12x, v12y, den;
};

__device__ __forceinline__ void fetch_tri(const EdgeArgs& a, int n, int id, Tri2& t) {
  const int32_t* vip = a.vi + (int64_t)n * a.vis.s0 + (int64_t)id * a.vis.s1;
  t.i0 = vip[0]; t.i1 = vip[a.vis.s2]; t.i2 = vip[2 * a.vis.s2];
  const float* vp = a.v + (int64_t)n * a.vs.s0;
  const float* q0 = vp + (int64_t)t.i0 * a.vs.s1;
  const float* q1 = vp + (int64_t)t.i1 * a.vs.s1;
  const float* q2 = vp + (int64_t)t.i2 * a.vs.s1;
  t.p0x = q0[0]; t.p0y = q0[a.vs.s2];
  t.p1x = q1[0]; t.p1y = q1[a.vs.s2];
  const float p2x = q2[0], p2y = q2[a.vs.s2];
  t.v01x = sub_rn(t.p1x, t.p0x); t.v01y = sub_rn(t.p1y, t.p0y);
  t.v02x = sub_rn(p2x, t.p0x);   t.v02y = sub_rn(p2y, t.p0y);
  t.v12x = sub_rn(p2x, t.p1x);   t.v12y = sub_rn(p2y, t.p1y);
  t.den = diff_of_products(t.v01x, t.v02y, t.v01y, t.v02x);
}

// pix_in_tri (:30-70): top-left rule with plain (non-canonical) edge functions
__device__ __forceinline__ bool pix_in_tri(const Tri2& t, int x, int y) {
  if (t.den == 0.f) return false;
  const float px = (float)x, py = (float)y;
  const float q0x = sub_rn(px, t.p0x), q0y = sub_rn(py, t.p0y);
  const float q1x = sub_rn(px, t.p1x), q1y = sub_rn(py, t.p1y);
  const float s = t.den > 0.f ? 1.f : -1.f;
  const float b0 = mul_rn(diff_of_products(q1y, t.v12x, q1x, t.v12y), s);
  const float b1 = mul_rn(diff_of_products(q0x, t.v02y, q0y, t.v02x), s);
  const float b2 = mul_rn(diff_of_products(q0y, t.v01x, q0x, t.v01y), s);
  if (!(b0 >= 0.f && b1 >= 0.f && b2 >= 0.f)) return false;
  bool tl0, tl1, tl2;
  if (t.den > 0.f) {
    tl0 = (t.v12y < 0.f) || (t.v12y == 0.f && t.v12x > 0.f);
    tl1 = (t.v02y > 0.f) || (t.v02y == 0.f && t.v02x < 0.f);
    tl2 = (t.v01y < 0.f) || (t.v01y == 0.f && t.v01x > 0.f);
  } else {
    tl0 = (t.v12y > 0.f) || (t.v12y == 0.f && t.v12x < 0.f);
    tl1 = (t.v02y < 0.f) || (t.v02y == 0.f && t.v02x > 0.f);
    tl2 = (t.v01y > 0.f) || (t.v01y == 0.f && t.v01x < 0.f);
  }
  return !((b0 == 0.f && !tl0) || (b1 == 0.f && !tl1) || (b2 == 0.f && !tl2));
}

// get_tri_normal (:89-100): normalize(cross(p0 - p2, p1 - p0))
__device__ __forceinline__ float3 tri_normal(const EdgeArgs& a, int n, const Tri2& t) {
  const float* vp = a.v + (int64_t)n * a.vs.s0;
  const float* q0 = vp + (int64_t)t.i0 * a.vs.s1;
  const float* q1 = vp + (int64_t)t.i1 * a.vs.s1;
  const float* q2 = vp + (int64_t)t.i2 * a.vs.s1;
  const float ax = q0[0] - q2[0], ay = q0[a.vs.s2] - q2[a.vs.s2], az = q0[2 * a.vs.s2] - q2[2 * a.vs.s2];
  const float bx = q1[0] - q0[0], by = q1[a.vs.s2] - q0[a.vs.s2], bz = q1[2 * a.vs.s2] - q0[2 * a.vs.s2];
  const float cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx;
  const float r = rsqrt_approx(cx * cx + cy * cy + cz * cz);
  return make_float3(cx * r, cy * r, cz * r);
}

// get_dp_dr (:102-203)
__device__ __forceinline__ float2 dp_dr(float nvx, float nvy, float nfx, float nfy, float max_mag) {
  const float rv = rsqrt_approx(nvx * nvx + nvy * nvy);
  const float rf = rsqrt_approx(nfx * nfx + nfy * nfy);
  nvx *= rv; nvy *= rv; nfx *= rf; nfy *= rf;
  const float bx = -nfy, by = nfx;
  const float d = bx * nvx + by * nvy;
  float k;
  if (max_mag > 0.f) {
    const float safe = (d >= 0.f ? 1.f : -1.f) * epsclamp(fmaxf(fabsf(d), fabsf(bx) * rcp_approx(max_mag)));
    k = bx * rcp_approx(safe);
  } else {
    k = bx * rcp_approx(epsclamp(d));
  }
  return make_float2(k * nvx, k * nvy);
}

// sum_c (img[nb] - img[c]) * 0.5 * (g[nb] + g[c])   (:351-380)
__device__ __forceinline__ float grad_dot(const EdgeArgs& a, int n, int xc, int yc, int xn, int yn) {
  const float* ic = a.img + (int64_t)n * a.ims.s0 + (int64_t)yc * a.ims.s2 + (int64_t)xc * a.ims.s3;
  const float* in_ = a.img + (int64_t)n * a.ims.s0 + (int64_t)yn * a.ims.s2 + (int64_t)xn * a.ims.s3;
  const float* gc = a.go + (int64_t)n * a.gs.s0 + (int64_t)yc * a.gs.s2 + (int64_t)xc * a.gs.s3;
  const float* gn = a.go + (int64_t)n * a.gs.s0 + (int64_t)yn * a.gs.s2 + (int64_t)xn * a.gs.s3;
  float acc = 0.f;
  for (int c = 0; c < a.C; ++c) {
    const float di = in_[(int64_t)c * a.ims.s1] - ic[(int64_t)c * a.ims.s1];
    const float sg = gn[(int64_t)c * a.gs.s1] + gc[(int64_t)c * a.gs.s1];
    acc += di * (0.5f * sg);
  }
  return acc;
}

// Both sides of the pair (centre=(xc,yc) showing triangle ci, neighbour=(xn,yn) showing ni != ci).
// axis: 0 = horizontal pair (x gradient), 1 = vertical pair (y gradient).
// Returns (centre.axis, centre.z, neighbour.axis, neighbour.z) before the final negation.
__device__ __forceinline__ float4 pair_eval(const EdgeArgs& a, int n, int ci, int ni, int xc, int yc,
                                            int xn, int yn, int axis) {
  const bool cv = ci >= 0, nv = ni >= 0;      // (:290-292)
  bool c_in_n = false, n_in_c = false;
  Tri2 tc, tn;
  if (cv && nv) {                              // (:320-325)
    fetch_tri(a, n, ci, tc);
    fetch_tri(a, n, ni, tn);
    c_in_n = pix_in_tri(tn, xc, yc);
    n_in_c = pix_in_tri(tc, xn, yn);
  }
  if (!(c_in_n && n_in_c)) {                               // no intersection (:391-393, :408-410)
    const bool adj = cv && nv && !c_in_n && !n_in_c;      // (:338-341)
    const bool c_over = c_in_n && !n_in_c;                // l_over_r / u_over_d (:328-331)
    const bool n_over = n_in_c && !c_in_n;                // r_over_l / d_over_u
    const bool c_zero = !cv || n_over || adj, n_zero = !nv || c_over || adj;
    if (c_zero && n_zero) return make_float4(0.f, 0.f, 0.f, 0.f);
    const float g = grad_dot(a, n, xc, yc, xn, yn);
    return make_float4(c_zero ? 0.f : g, 0.f, n_zero ? 0.f : g, 0.f);
  }
  // intersection: both triangles valid (:394-406, :411-423)
  const float g = grad_dot(a, n, xc, yc, xn, yn);
  const float3 nc = tri_normal(a, n, tc), nn = tri_normal(a, n, tn);
  const float nca = axis == 0 ? nc.x : nc.y, nna = axis == 0 ? nn.x : nn.y;
  const float2 dc = dp_dr(nca, nc.z, nna, nn.z, a.max_dp_dr);
  const float2 dn = dp_dr(nna, nn.z, nca, nc.z, a.max_dp_dr);
  return make_float4(g * dc.x, g * dc.y, g * dn.x, g * dn.y);
}

// One CTA per 64 x 8 output tile.
//   phase 0: triangle ids of the tile plus a one-pixel ring -> shared memory
//   phase 1: every (centre, right) / (centre, down) pair that touches the tile and shows two
//            different triangles becomes a JOB (warp-ballot compaction; ~20 % of the pair slots on
//            the 100k-triangle config, far fewer on large triangles)
//   phase 2: one thread per job evaluates the pair ONCE (the scatter form's work, all lanes busy)
//            and drops the centre / neighbour contributions into per-role slots of the two pixels
//            (each slot has exactly one writer, so no atomics)
//   phase 3: every pixel adds its <= 8 slots and writes its three outputs once, coalesced.
constexpr int kETW = 64, kETH = 8, kEThreads = 256;
constexpr int kEIW = kETW + 2, kEIH = kETH + 2;                 // id tile with ring
constexpr int kEHSlots = (kETW + 1) * kETH;                     // horizontal pairs: centres x0-1 .. x0+TW-1
constexpr int kEVSlots = kETW * (kETH + 1);                     // vertical pairs:   centres y0-1 .. y0+TH-1
constexpr int kESlots = kEHSlots + kEVSlots;

// FUSED: instead of writing dL/d(v_pix_img), push it straight through the backward of
// interpolate(v_pix, vi, index_img, bary) (drtk/edge_grad_estimator.py:172): grad_v_pix[vi_k] += g * bary_k.
// The estimator's gradient image is sparse (non-zero only at silhouettes / overlaps / intersections), so
// a handful of direct REDs replaces a 12 B/px write, a 28 B/px re-read and a whole reduction kernel.
struct FusedArgs {
  const float* bary;
  Strides4 bs;
  float* grad_v;  // [N,V,3], zero-filled by the launcher
};

template <bool FUSED>
__global__ void __launch_bounds__(kEThreads, 4) edge_grad_tile_kernel(EdgeArgs a, float* __restrict__ out, FusedArgs fz) {
  __shared__ int ids[kEIH * kEIW];
  __shared__ int jobs[kESlots];
  __shared__ int njobs;
  __shared__ __align__(16) float slot[8][kETW * kETH];  // 0 cx, 1 czx, 2 cy, 3 czy (centre roles); 4 rx, 5 rz, 6 dy, 7 dz
  const int tid = threadIdx.x, lane = tid & 31;
  const int n = blockIdx.z;
  const int x0 = blockIdx.x * kETW, y0 = blockIdx.y * kETH;
  const int32_t* ip = a.index_img + (int64_t)n * a.is.s0;

  for (int i = tid; i < kEIH * kEIW; i += kEThreads) {
    const int ly = i / kEIW, lx = i - ly * kEIW;
    const int x = x0 - 1 + lx, y = y0 - 1 + ly;
    ids[i] = (x >= 0 && x < a.W && y >= 0 && y < a.H) ? ip[(int64_t)y * a.is.s1 + (int64_t)x * a.is.s2] : -2;
  }
  for (int i = tid; i < 8 * kETW * kETH / 4; i += kEThreads)  // 16 KB of slots, 128-bit stores
    reinterpret_cast<float4*>(&slot[0][0])[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tid == 0) njobs = 0;
  __syncthreads();

  // ---- phase 1: job list ----
  for (int base = 0; base < kESlots; base += kEThreads) {
    const int sidx = base + tid;
    bool has = false;
    int job = 0;
    if (sidx < kESlots) {
      int lx, ly, axis;  // centre in id-tile coordinates
      if (sidx < kEHSlots) { axis = 0; ly = 1 + sidx / (kETW + 1); lx = sidx % (kETW + 1); }
      else { const int t = sidx - kEHSlots; axis = 1; ly = t / kETW; lx = 1 + t % kETW; }
      const int cx = x0 - 1 + lx, cy = y0 - 1 + ly;
      // only pixels with 0 <= x < W-1 and 0 <= y < H-1 act as centre (:270)
      if (cx >= 0 && cy >= 0 && cx < a.W - 1 && cy < a.H - 1) {
        const int ci = ids[ly * kEIW + lx];
        const int ni = axis == 0 ? ids[ly * kEIW + lx + 1] : ids[(ly + 1) * kEIW + lx];
        has = ci != ni;
        job = (axis << 16) | (ly << 8) | lx;
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, has);
    int wbase = 0;
    if (lane == 0 && m) wbase = atomicAdd(&njobs, __popc(m));
    wbase = __shfl_sync(0xffffffffu, wbase, 0);
    if (has) jobs[wbase + __popc(m & ((1u << lane) - 1u))] = job;
  }
  __syncthreads();

  // ---- phase 2: evaluate each pair once ----
  const int nj = njobs;
  for (int j = tid; j < nj; j += kEThreads) {
    const int job = jobs[j];
    const int axis = job >> 16, ly = (job >> 8) & 0xff, lx = job & 0xff;
    const int cx = x0 - 1 + lx, cy = y0 - 1 + ly;
    const int nx = cx + (axis == 0), ny = cy + (axis == 1);
    const int ci = ids[ly * kEIW + lx];
    const int ni = axis == 0 ? ids[ly * kEIW + lx + 1] : ids[(ly + 1) * kEIW + lx];
    const float4 r = pair_eval(a, n, ci, ni, cx, cy, nx, ny, axis);
    // centre pixel inside the tile?
    if (lx >= 1 && ly >= 1) {
      const int o = (ly - 1) * kETW + (lx - 1);
      slot[axis == 0 ? 0 : 2][o] = r.x;
      slot[axis == 0 ? 1 : 3][o] = r.y;
    }
    const int nlx = lx + (axis == 0), nly = ly + (axis == 1);
    if (nlx <= kETW && nly <= kETH) {  // neighbour pixel inside the tile (it is >= 1 by construction)
      const int o = (nly - 1) * kETW + (nlx - 1);
      slot[axis == 0 ? 4 : 6][o] = r.z;
      slot[axis == 0 ? 5 : 7][o] = r.w;
    }
  }
  __syncthreads();

  // ---- phase 3: combine and store (negated sums, :431-445) ----
  const int64_t HW = (int64_t)a.H * a.W;
  float* ob = FUSED ? nullptr : out + (int64_t)n * 3 * HW;
  for (int i = tid; i < kETW * kETH; i += kEThreads) {
    const int ly = i / kETW, lx = i - ly * kETW;
    const int x = x0 + lx, y = y0 + ly;
    if (x >= a.W || y >= a.H) continue;
    const float gx = -(slot[0][i] + slot[4][i]);
    const float gy = -(slot[2][i] + slot[6][i]);
    const float gz = -((slot[1][i] + slot[3][i]) + slot[5][i] + slot[7][i]);
    if (!FUSED) {
      float* o = ob + (int64_t)y * a.W + x;
      o[0] = gx; o[HW] = gy; o[2 * HW] = gz;
    } else if (gx != 0.f || gy != 0.f || gz != 0.f) {
      const int id = ids[(ly + 1) * kEIW + lx + 1];
      if (id != -1) {  // interpolate's backward only touches covered pixels
        const int32_t* vip = a.vi + (int64_t)n * a.vis.s0 + (int64_t)id * a.vis.s1;
        const float* bp = fz.bary + (int64_t)n * fz.bs.s0 + (int64_t)y * fz.bs.s2 + (int64_t)x * fz.bs.s3;
        float* gv = fz.grad_v + (int64_t)n * a.V * 3;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float bk = bp[(int64_t)k * fz.bs.s1];
          float* q = gv + (int64_t)vip[(int64_t)k * a.vis.s2] * 3;
          red_add(q + 0, gx * bk); red_add(q + 1, gy * bk); red_add(q + 2, gz * bk);
        }
      }
    }
  }
}

}  // namespace
}  // namespace drtk

using namespace drtk;

static int edge_launch(const float* v_pix, const int64_t* v_strides, const float* img, const int64_t* img_strides,
                       const int32_t* index_img, const int64_t* index_strides, const int32_t* vi,
                       const int64_t* vi_strides, const float* grad_output, const int64_t* grad_output_strides,
                       int64_t N, int64_t V, int64_t F, int64_t C, int64_t H, int64_t W, float max_dp_dr,
                       float* grad_v_pix_img, const float* bary_img, const int64_t* bary_strides, float* grad_v_pix,
                       void* stream_) {
  if (N < 0 || C < 0 || H < 0 || W < 0) return DRTK_B200_EINVAL;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool fused = grad_v_pix != nullptr;
  if (fused && N * V > 0) DRTK_CUDA(cudaMemsetAsync(grad_v_pix, 0, sizeof(float) * (size_t)(N * V * 3), stream));
  const int64_t npix = N * H * W;
  if (npix == 0) return 0;
  if (!v_pix || !img || !index_img || !vi || !grad_output || (!fused && !grad_v_pix_img) || (fused && !bary_img))
    return DRTK_B200_EINVAL;
  if (H > (1 << 30) || W > (1 << 30) || N > 65535) return DRTK_B200_EUNSUPPORTED;
  EdgeArgs a;
  a.v = v_pix; a.vs = make3(v_strides); a.img = img; a.ims = make4(img_strides);
  a.index_img = index_img; a.is = make3(index_strides); a.vi = vi; a.vis = make3(vi_strides);
  a.go = grad_output; a.gs = make4(grad_output_strides);
  a.N = (int)N; a.V = (int)V; a.F = (int)F; a.C = (int)C; a.H = (int)H; a.W = (int)W;
  a.max_dp_dr = max_dp_dr;
  const dim3 grid((unsigned)((W + kETW - 1) / kETW), (unsigned)((H + kETH - 1) / kETH), (unsigned)N);
  if (grid.y > 65535) return DRTK_B200_EUNSUPPORTED;
  FusedArgs fz;
  fz.bary = bary_img; fz.bs = fused ? make4(bary_strides) : Strides4{0, 0, 0, 0}; fz.grad_v = grad_v_pix;
  if (fused) edge_grad_tile_kernel<true><<<grid, kEThreads, 0, stream>>>(a, nullptr, fz);
  else edge_grad_tile_kernel<false><<<grid, kEThreads, 0, stream>>>(a, grad_v_pix_img, fz);
  DRTK_CHECK_LAUNCH();
  return 0;
}

extern "C" int drtk_b200_edge_grad_backward(const float* v_pix, const int64_t* v_strides, const float* img,
                                            const int64_t* img_strides, const int32_t* index_img,
                                            const int64_t* index_strides, const int32_t* vi,
                                            const int64_t* vi_strides, const float* grad_output,
                                            const int64_t* grad_output_strides, int64_t N, int64_t V,
                                            int64_t F, int64_t C, int64_t H, int64_t W, float max_dp_dr,
                                            float* grad_v_pix_img, void* stream_) {
  return edge_launch(v_pix, v_strides, img, img_strides, index_img, index_strides, vi, vi_strides, grad_output,
                     grad_output_strides, N, V, F, C, H, W, max_dp_dr, grad_v_pix_img, nullptr, nullptr, nullptr,
                     stream_);
}

extern "C" int drtk_b200_edge_grad_backward_fused(
    const float* v_pix, const int64_t* v_strides, const float* img, const int64_t* img_strides,
    const int32_t* index_img, const int64_t* index_strides, const int32_t* vi, const int64_t* vi_strides,
    const float* grad_output, const int64_t* grad_output_strides, const float* bary_img,
    const int64_t* bary_strides, int64_t N, int64_t V, int64_t F, int64_t C, int64_t H, int64_t W,
    float max_dp_dr, float* grad_v_pix, void* stream_) {
  if (!grad_v_pix && N * V > 0) return DRTK_B200_EINVAL;
  if (N * V == 0) return 0;
  return edge_launch(v_pix, v_strides, img, img_strides, index_img, index_strides, vi, vi_strides, grad_output,
                     grad_output_strides, N, V, F, C, H, W, max_dp_dr, nullptr, bary_img, bary_strides, grad_v_pix,
                     stream_);
}
