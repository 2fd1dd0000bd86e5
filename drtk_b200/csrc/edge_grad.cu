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
  float p0x, p0y, p1x, p1y, v01x, v01y, v02x, v02y, v12x, v12y, den;
};

__device__ __forceinline__ void derive_tri(float p0x, float p0y, float p1x, float p1y, float p2x, float p2y, Tri2& t) {
  t.p0x = p0x; t.p0y = p0y; t.p1x = p1x; t.p1y = p1y;
  t.v01x = sub_rn(t.p1x, t.p0x); t.v01y = sub_rn(t.p1y, t.p0y);
  t.v02x = sub_rn(p2x, t.p0x);   t.v02y = sub_rn(p2y, t.p0y);
  t.v12x = sub_rn(p2x, t.p1x);   t.v12y = sub_rn(p2y, t.p1y);
  t.den = diff_of_products(t.v01x, t.v02y, t.v01y, t.v02x);
}

// two-level gather index -> vi -> v (any strides)
struct GatherFetch {
  const EdgeArgs& a;
  int n;
  __device__ __forceinline__ void operator()(int id, Tri2& t) const {
    const int32_t* vip = a.vi + (int64_t)n * a.vis.s0 + (int64_t)id * a.vis.s1;
    const int i0 = vip[0], i1 = vip[a.vis.s2], i2 = vip[2 * a.vis.s2];
    const float* vp = a.v + (int64_t)n * a.vs.s0;
    const float* q0 = vp + (int64_t)i0 * a.vs.s1;
    const float* q1 = vp + (int64_t)i1 * a.vs.s1;
    const float* q2 = vp + (int64_t)i2 * a.vs.s1;
    derive_tri(q0[0], q0[a.vs.s2], q1[0], q1[a.vs.s2], q2[0], q2[a.vs.s2], t);
  }
};

// one-level gather from the per-(image, triangle) table {p0x,p0y,p1x,p1y | p2x,p2y,-,-} built by xy_table_kernel
struct TableFetch {
  const float4* tn;  // table of image n
  __device__ __forceinline__ void operator()(int id, Tri2& t) const {
    const float8 r = ldg_f8(reinterpret_cast<const float*>(tn + (int64_t)id * 2));  // one 256-bit load per 32-B row
    derive_tri(r.lo.x, r.lo.y, r.lo.z, r.lo.w, r.hi.x, r.hi.y, t);
  }
};

// pix_in_tri (:30-70): top-left rule with plain (non-canonical) edge functions.  The top/left classification of the
// edges is only evaluated for a sample that lies exactly ON an edge (some b == 0); strictly inside / outside -- all
// but a handful of samples -- is decided by the three signs.
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
  if (fminf(fminf(b0, b1), b2) > 0.f) return true;
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
__device__ __forceinline__ float3 tri_normal(const EdgeArgs& a, int n, int id) {
  const int32_t* vip = a.vi + (int64_t)n * a.vis.s0 + (int64_t)id * a.vis.s1;
  const float* vp = a.v + (int64_t)n * a.vs.s0;
  const float* q0 = vp + (int64_t)vip[0] * a.vs.s1;
  const float* q1 = vp + (int64_t)vip[a.vis.s2] * a.vs.s1;
  const float* q2 = vp + (int64_t)vip[2 * a.vis.s2] * a.vs.s1;
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

#ifndef DRTK_EDGE_DOT_BATCH
#define DRTK_EDGE_DOT_BATCH 8  // measured (config 3 / 4 / 5 / 4-overdraw, ms): 1 -> .0886 / .2771 / 1.082 / 1.422, 4 -> same, 8 -> .0825 / .2628 / 1.065 / 1.381, 16 -> .0874 / .2695 / 1.071 / 1.424
#endif
// sum_c (img[nb] - img[c]) * 0.5 * (g[nb] + g[c])   (:351-380)
__device__ __forceinline__ float grad_dot(const EdgeArgs& a, int n, int xc, int yc, int xn, int yn) {
  const float* ic = a.img + (int64_t)n * a.ims.s0 + (int64_t)yc * a.ims.s2 + (int64_t)xc * a.ims.s3;
  const float* in_ = a.img + (int64_t)n * a.ims.s0 + (int64_t)yn * a.ims.s2 + (int64_t)xn * a.ims.s3;
  const float* gc = a.go + (int64_t)n * a.gs.s0 + (int64_t)yc * a.gs.s2 + (int64_t)xc * a.gs.s3;
  const float* gn = a.go + (int64_t)n * a.gs.s0 + (int64_t)yn * a.gs.s2 + (int64_t)xn * a.gs.s3;
  float acc = 0.f;
  int c = 0;
#if DRTK_EDGE_DOT_BATCH > 1
  // channel planes are H*W apart: every load of this loop is its own DRAM round trip, so they are issued in batches
  // (the summation order over c stays that of the reference)
  for (; c + DRTK_EDGE_DOT_BATCH <= a.C; c += DRTK_EDGE_DOT_BATCH) {
    float vi_[DRTK_EDGE_DOT_BATCH], vc_[DRTK_EDGE_DOT_BATCH], gn_[DRTK_EDGE_DOT_BATCH], gc_[DRTK_EDGE_DOT_BATCH];
#pragma unroll
    for (int k = 0; k < DRTK_EDGE_DOT_BATCH; ++k) {
      vi_[k] = in_[(int64_t)(c + k) * a.ims.s1]; vc_[k] = ic[(int64_t)(c + k) * a.ims.s1];
      gn_[k] = gn[(int64_t)(c + k) * a.gs.s1];   gc_[k] = gc[(int64_t)(c + k) * a.gs.s1];
    }
#pragma unroll
    for (int k = 0; k < DRTK_EDGE_DOT_BATCH; ++k) acc += (vi_[k] - vc_[k]) * (0.5f * (gn_[k] + gc_[k]));
  }
#endif
  for (; c < a.C; ++c) {
    const float di = in_[(int64_t)c * a.ims.s1] - ic[(int64_t)c * a.ims.s1];
    const float sg = gn[(int64_t)c * a.gs.s1] + gc[(int64_t)c * a.gs.s1];
    acc += di * (0.5f * sg);
  }
  return acc;
}

// Both sides of the pair (centre=(xc,yc) showing triangle ci, neighbour=(xn,yn) showing ni != ci).
// axis: 0 = horizontal pair (x gradient), 1 = vertical pair (y gradient).
// Returns (centre.axis, centre.z, neighbour.axis, neighbour.z) before the final negation.
template <class Fetch>
__device__ __forceinline__ float4 pair_eval(const EdgeArgs& a, int n, int ci, int ni, int xc, int yc,
                                            int xn, int yn, int axis, const Fetch& fetch) {
  const bool cv = ci >= 0, nv = ni >= 0;      // (:290-292)
  bool c_in_n = false, n_in_c = false;
  Tri2 tc, tn;
  if (cv && nv) {                              // (:320-325)
    fetch(ci, tc);
    fetch(ni, tn);
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
  const float3 nc = tri_normal(a, n, ci), nn = tri_normal(a, n, ni);
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
    const float4 r = pair_eval(a, n, ci, ni, cx, cy, nx, ny, axis, GatherFetch{a, n});
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

// ------------------------------------------------------------------------------------------
// Fused path, warp-strip formulation (dense index rows, W % 8 == 0).
// Because grad_v_pix is LINEAR in the per-pixel gradient, the per-pixel sum over a pixel's <= 4 pairs
// need not be formed at all: every evaluated pair scatters its (rare) non-zero contributions straight to
// the vertices of the two pixels' triangles.  That removes the tile kernel's slots, its zero-fill, its
// combine phase and all three CTA barriers; what remains is "find the pairs showing two different
// triangles and classify them", done warp-privately:
//   * a warp takes a strip of 256 consecutive pixels of one row: lane = 8 pixels, two LDG.128 for its own
//     row and two for the row below (the right neighbour of a lane's last pixel comes from the next lane by
//     shuffle), so the 16 candidate pairs of a lane are 16 register compares;
//   * the lanes' jobs are compacted into a warp-private shared-memory queue (shuffle scan of the per-lane
//     counts, no atomics, __syncwarp only), so the heavy pair classification runs with all lanes busy;
//   * triangles come from a per-(image, triangle) xy table (2 x LDG.128, one dependent level) built by a
//     pre-pass, instead of the index -> vi -> v two-level gather (3 + 6 scalar loads per triangle).
constexpr int kStripPx = 256, kStripWarps = 8;

// Also zero-fills grad_v_pix, which the strip kernel accumulates into (saves a memset launch per step).
__global__ void __launch_bounds__(256) xy_table_kernel(EdgeArgs a, float4* __restrict__ table, float* __restrict__ zero,
                                                       int64_t zero_count, unsigned long long* __restrict__ work_counter) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *work_counter = 0ull;  // the strip kernel's item counter
  {
    const int64_t gid = ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * gridDim.y * blockDim.x;
    for (int64_t i = gid; i < zero_count; i += nthreads) zero[i] = 0.f;
  }
  if (f >= a.F) return;
  const int32_t* vip = a.vi + (int64_t)n * a.vis.s0 + (int64_t)f * a.vis.s1;
  const float* vp = a.v + (int64_t)n * a.vs.s0;
  const float* q0 = vp + (int64_t)vip[0] * a.vs.s1;
  const float* q1 = vp + (int64_t)vip[a.vis.s2] * a.vs.s1;
  const float* q2 = vp + (int64_t)vip[2 * a.vis.s2] * a.vs.s1;
  float4* row = table + ((int64_t)n * a.F + f) * 2;
  row[0] = make_float4(q0[0], q0[a.vs.s2], q1[0], q1[a.vs.s2]);
  row[1] = make_float4(q2[0], q2[a.vs.s2], 0.f, 0.f);
}

// grad_v_pix[vi[id][k]] += (gx, gy, gz) * bary_k(x, y)  for the pixel (x, y) showing triangle id
__device__ __forceinline__ void scatter_pixel(const EdgeArgs& a, const FusedArgs& fz, int n, int id, int x, int y,
                                              int axis, float ga, float gz) {
  if (id < 0 || (ga == 0.f && gz == 0.f)) return;
  const int32_t* vip = a.vi + (int64_t)n * a.vis.s0 + (int64_t)id * a.vis.s1;
  const float* bp = fz.bary + (int64_t)n * fz.bs.s0 + (int64_t)y * fz.bs.s2 + (int64_t)x * fz.bs.s3;
  float* gv = fz.grad_v + (int64_t)n * a.V * 3;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float bk = bp[(int64_t)k * fz.bs.s1];
    float* q = gv + (int64_t)vip[(int64_t)k * a.vis.s2] * 3;
    if (ga != 0.f) red_add(q + axis, ga * bk);
    if (gz != 0.f) red_add(q + 2, gz * bk);
  }
}

// The RARE part of a pair -- everything after the two inside tests of pair_eval, plus the scatter through the conduit.
// The strip kernel runs it in a loop of its own over the few pairs that survive classification, so that its
// registers and 64-bit address arithmetic do not burden the loop that classifies the bulk of the pairs (neighbouring
// triangles of a watertight mesh: "adjacent", no contribution).
__device__ __forceinline__ void pair_contribute(const EdgeArgs& a, const FusedArgs& fz, int n, int ci, int ni, int cx, int cy,
                                             int axis, bool c_in_n, bool n_in_c) {
  const int nx = cx + (axis == 0), ny = cy + (axis == 1);
  const bool cv = ci >= 0, nv = ni >= 0;
  float4 r;
  if (!(c_in_n && n_in_c)) {                               // no intersection (:391-393, :408-410)
    const bool adj = cv && nv && !c_in_n && !n_in_c;      // (:338-341)
    const bool c_over = c_in_n && !n_in_c, n_over = n_in_c && !c_in_n;
    const bool c_zero = !cv || n_over || adj, n_zero = !nv || c_over || adj;
    if (c_zero && n_zero) return;
    const float g = grad_dot(a, n, cx, cy, nx, ny);
    r = make_float4(c_zero ? 0.f : g, 0.f, n_zero ? 0.f : g, 0.f);
  } else {                                                 // intersection (:394-406, :411-423)
    const float g = grad_dot(a, n, cx, cy, nx, ny);
    const float3 nc = tri_normal(a, n, ci), nn = tri_normal(a, n, ni);
    const float nca = axis == 0 ? nc.x : nc.y, nna = axis == 0 ? nn.x : nn.y;
    const float2 dc = dp_dr(nca, nc.z, nna, nn.z, a.max_dp_dr);
    const float2 dn = dp_dr(nna, nn.z, nca, nc.z, a.max_dp_dr);
    r = make_float4(g * dc.x, g * dc.y, g * dn.x, g * dn.y);
  }
  // final negation of the reference (:431-445) folded in
  scatter_pixel(a, fz, n, ci, cx, cy, axis, -r.x, -r.y);
  scatter_pixel(a, fz, n, ni, nx, ny, axis, -r.z, -r.w);
}

// Measured and NOT adopted (profiles/r02_opbench_edge_split_classify_contribute.txt): splitting this kernel into a
// classification kernel (48-64 registers, survivors written as 6 bits per pixel into a byte image) and a contribution
// kernel (80 registers, scans the byte image and evaluates the marked pairs with full warps).  Unconstrained the fused
// kernel wants 158 registers, nearly all for the contribution code, so the split looked like the way to un-spill the
// classification loop -- but config 3 / 4 / 5 ran at 0.091 / 0.295 / 1.26 ms against 0.082 / 0.264 / 1.06 ms fused (the
// classification is bound by its id -> table row -> inside-test latency chain, not by the spills, and the second
// kernel's 1 B/px scan costs 30 us), and the overdraw scene gained only 5-10 % (1.24-1.31 against 1.385 ms).
// Work item of a warp: a block of kStripRows consecutive centre rows x 256 columns.  The row below a centre row is the
// next centre row, so it stays in registers (one index-row load per row instead of two).
#ifndef DRTK_EDGE_PREFETCH
#define DRTK_EDGE_PREFETCH 1
#endif
#ifndef DRTK_EDGE_STRIP_ROWS
#define DRTK_EDGE_STRIP_ROWS 2  // measured on B200 (config 3 / 4 / 5 / 4-overdraw, ms): 2 -> .090 / .279 / 1.09 / 1.41, 4 -> .110 / .291 / 1.05 / 1.51, 8 -> .130 / .276 / 1.02 / -; round-1 kernel .087 / .283 / 1.14 / 1.67
#endif
constexpr int kStripRows = DRTK_EDGE_STRIP_ROWS;

__global__ void __launch_bounds__(kStripWarps * 32, 4) edge_grad_strip_kernel(const __grid_constant__ EdgeArgs a,
                                                                           const __grid_constant__ FusedArgs fz,
                                                                           const float4* __restrict__ table,
                                                                           int strips_per_row, int row_blocks,
                                                                           int64_t num_items,
                                                                           unsigned long long* __restrict__ work_counter) {
  __shared__ int s_own[kStripWarps][kStripPx + 1];
  __shared__ int s_below[kStripWarps][kStripPx];
  __shared__ unsigned short s_jobs[kStripWarps][2 * kStripPx];
  __shared__ unsigned short s_keep[kStripWarps][2 * kStripPx];  // pairs that survive classification
  __shared__ int n_keep[kStripWarps];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int* own = s_own[wid];
  int* below = s_below[wid];
  unsigned short* jobs = s_jobs[wid];
  unsigned short* keepq = s_keep[wid];
  if (lane == 0) n_keep[wid] = 0;
  __syncwarp();
  const int items_per_img = strips_per_row * row_blocks;

  // items are handed out dynamically (one atomic per warp and item): the jobs per item vary with the scene, and a
  // static split left the last wave of warps half empty
  for (;;) {
    unsigned long long s = 0;
    if (lane == 0) s = atomicAdd(work_counter, 1ull);
    s = __shfl_sync(0xffffffffu, s, 0);
    if (s >= (unsigned long long)num_items) break;
    const int n = (int)(s / items_per_img);
    const int r = (int)(s - (int64_t)n * items_per_img);
    const int yb = r / strips_per_row, sx = r - yb * strips_per_row;
    const int y0 = yb * kStripRows;
    const int x = sx * kStripPx + lane * 8;
    const bool live = x < a.W;  // W % 8 == 0: a lane's eight pixels are all inside or all outside
    const int32_t* img_n = a.index_img + (int64_t)n * a.is.s0;
    const TableFetch fetch{table + (int64_t)n * a.F * 2};
    int id[9], dn[8];
    int next_first = -1;  // first pixel of the next lane's... (right neighbour of this lane's last pixel), row below
    {
      const int32_t* row = img_n + (int64_t)y0 * a.is.s1;
      if (live) {
        const int4 p = ldg_stream_i4(row + x), q = ldg_stream_i4(row + x + 4);
        dn[0] = p.x; dn[1] = p.y; dn[2] = p.z; dn[3] = p.w; dn[4] = q.x; dn[5] = q.y; dn[6] = q.z; dn[7] = q.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) dn[j] = -1;
      }
      next_first = __shfl_down_sync(0xffffffffu, dn[0], 1);
      if (lane == 31) next_first = (live && x + 8 < a.W) ? row[x + 8] : -1;
    }
#pragma unroll 1
    for (int ry = 0; ry < kStripRows; ++ry) {
      const int y = y0 + ry;
      if (y >= a.H - 1) break;  // the last image row holds no centre pixel (:270); uniform over the warp
      // the row loaded as "below" in the previous step is this step's centre row
#pragma unroll
      for (int j = 0; j < 8; ++j) id[j] = dn[j];
      id[8] = next_first;
      const int32_t* row = img_n + (int64_t)(y + 1) * a.is.s1;
      if (live) {
        const int4 u = ldg_stream_i4(row + x), w = ldg_stream_i4(row + x + 4);
        dn[0] = u.x; dn[1] = u.y; dn[2] = u.z; dn[3] = u.w; dn[4] = w.x; dn[5] = w.y; dn[6] = w.z; dn[7] = w.w;
      }
      next_first = __shfl_down_sync(0xffffffffu, dn[0], 1);
      if (lane == 31) next_first = (live && x + 8 < a.W) ? row[x + 8] : -1;
      // candidate pairs: bit j = (j, j+1) horizontal, bit 8+j = (j, below) vertical; centres need x < W-1
      unsigned m = 0;
      if (live) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const bool centre = x + j < a.W - 1;
          m |= (centre && id[j] != id[j + 1]) ? (1u << j) : 0u;
          m |= (centre && id[j] != dn[j]) ? (0x100u << j) : 0u;
        }
      }
      if (__all_sync(0xffffffffu, m == 0u)) continue;
      // stage the ids for the job phase, compact the jobs
      __syncwarp();  // previous step's readers are done
#pragma unroll
      for (int j = 0; j < 8; ++j) { own[lane * 8 + j] = id[j]; below[lane * 8 + j] = dn[j]; }
      if (lane == 31) own[kStripPx] = id[8];
      const int cnt = __popc(m);
      int off = cnt;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, off, d);
        if (lane >= d) off += t;
      }
      const int total = __shfl_sync(0xffffffffu, off, 31);
      off -= cnt;
      while (m) {
        const int bit = __ffs(m) - 1;
        m &= m - 1;
        jobs[off++] = (unsigned short)(((bit >> 3) << 8) | (lane * 8 + (bit & 7)));
      }
      __syncwarp();
      // One job per lane and step.  The two table rows of the NEXT step's job are requested into L1 while this
      // step's job is classified (the id -> table row -> inside test chain is what the kernel waits on).
      int job = lane < total ? jobs[lane] : 0;
      for (int q = lane; q < total; q += 32) {
        const int axis = job >> 8, lx = job & 0xff;
        const int ci = own[lx];
        const int ni = axis == 0 ? own[lx + 1] : below[lx];
#if DRTK_EDGE_PREFETCH
        int job_next = 0;
        if (q + 32 < total) {
          job_next = jobs[q + 32];
          const int lxn = job_next & 0xff;
          const int cn = own[lxn], nn = (job_next >> 8) == 0 ? own[lxn + 1] : below[lxn];
          if (cn >= 0 && nn >= 0) {
            prefetch_l1(fetch.tn + (int64_t)cn * 2);
            prefetch_l1(fetch.tn + (int64_t)nn * 2);
          }
        }
#else
        const int job_next = q + 32 < total ? jobs[q + 32] : 0;
#endif
        const int cx = sx * kStripPx + lx;
        bool c_in_n = false, n_in_c = false, keep = true;
        if (ci >= 0 && ni >= 0) {  // (:320-325) the bulk: two neighbouring triangles, neither covers the other's pixel
          Tri2 tc, tn;
          fetch(ci, tc);
          fetch(ni, tn);
          c_in_n = pix_in_tri(tn, cx, y);
          n_in_c = pix_in_tri(tc, cx + (axis == 0), y + (axis == 1));
          keep = c_in_n || n_in_c;  // else adjacent (:338-341): no contribution
        }
        if (keep) keepq[atomicAdd(&n_keep[wid], 1)] = (unsigned short)(job | (c_in_n ? 0x4000 : 0) | (n_in_c ? 0x8000 : 0));
        job = job_next;
      }
      __syncwarp();
      const int nk = n_keep[wid];
      __syncwarp();
      if (nk == 0) continue;
      if (lane == 0) n_keep[wid] = 0;
      // survivors: silhouettes, occlusions, intersections
      for (int q = lane; q < nk; q += 32) {
        const int job = keepq[q];
        const int axis = (job >> 8) & 1, lx = job & 0xff;
        const int ci = own[lx];
        const int ni = axis == 0 ? own[lx + 1] : below[lx];
        pair_contribute(a, fz, n, ci, ni, sx * kStripPx + lx, y, axis, (job & 0x4000) != 0, (job & 0x8000) != 0);
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
                       void* workspace, size_t workspace_bytes, void* stream_) {
  if (N < 0 || C < 0 || H < 0 || W < 0) return DRTK_B200_EINVAL;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool fused = grad_v_pix != nullptr;
  // grad_v_pix is zero-filled by xy_table_kernel on the strip path, by a memset everywhere else
  auto zero_grad_v = [&]() -> int {
    if (fused && N * V > 0) DRTK_CUDA(cudaMemsetAsync(grad_v_pix, 0, sizeof(float) * (size_t)(N * V * 3), stream));
    return 0;
  };
  const int64_t npix = N * H * W;
  if (npix == 0) return zero_grad_v();
  if (!v_pix || !img || !index_img || (!vi && F > 0) || !grad_output || (!fused && !grad_v_pix_img) || (fused && !bary_img))
    return DRTK_B200_EINVAL;
  if (H > (1 << 30) || W > (1 << 30)) return DRTK_B200_EUNSUPPORTED;
  if (N > kMaxBatchPerLaunch) {  // batch index rides on gridDim.y/z: slices of the batch (the workspace is reused)
    for (int64_t n0 = 0; n0 < N; n0 += kMaxBatchPerLaunch) {
      const int64_t nn = (N - n0 < kMaxBatchPerLaunch) ? N - n0 : kMaxBatchPerLaunch;
      const int rc = edge_launch(v_pix + n0 * v_strides[0], v_strides, img + n0 * img_strides[0], img_strides,
                                 index_img + n0 * index_strides[0], index_strides, vi + n0 * vi_strides[0], vi_strides,
                                 grad_output + n0 * grad_output_strides[0], grad_output_strides, nn, V, F, C, H, W,
                                 max_dp_dr, grad_v_pix_img ? grad_v_pix_img + n0 * 3 * H * W : nullptr,
                                 bary_img ? bary_img + n0 * bary_strides[0] : nullptr, bary_strides,
                                 grad_v_pix ? grad_v_pix + n0 * V * 3 : nullptr, workspace, workspace_bytes, stream_);
      if (rc) return rc;
    }
    return 0;
  }
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
  const auto al16 = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
  if (fused && F > 0 && W % 8 == 0 && H > 1 && a.is.s2 == 1 && a.is.s1 % 4 == 0 && a.is.s0 % 4 == 0 && al16(index_img) &&
      F <= 65535LL * 256) {
    const size_t tb = (size_t)(N * F) * 32;
    if (!workspace || workspace_bytes < tb + 64) return DRTK_B200_EWORKSPACE;
    float4* table = reinterpret_cast<float4*>((reinterpret_cast<uintptr_t>(workspace) + 31) & ~uintptr_t(31));
    unsigned long long* work_counter = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(table) + tb);
    xy_table_kernel<<<dim3((unsigned)((F + 255) / 256), (unsigned)N), 256, 0, stream>>>(a, table, grad_v_pix, N * V * 3, work_counter);
    const int strips_per_row = (int)((W + kStripPx - 1) / kStripPx);
    const int row_blocks = (int)((H - 1 + kStripRows - 1) / kStripRows);
    const int64_t num_items = N * (int64_t)row_blocks * strips_per_row;
    const int64_t need = (num_items + kStripWarps - 1) / kStripWarps;
    const int64_t cap = (int64_t)num_sms() * 4;  // one resident wave; the warps pull items from a counter
    edge_grad_strip_kernel<<<(unsigned)(need < cap ? need : cap), kStripWarps * 32, 0, stream>>>(a, fz, table, strips_per_row,
                                                                                          row_blocks, num_items, work_counter);
    DRTK_CHECK_LAUNCH();
    return 0;
  }
  if (const int rc = zero_grad_v()) return rc;
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
                     nullptr, 0, stream_);
}

extern "C" int drtk_b200_edge_grad_backward_fused(
    const float* v_pix, const int64_t* v_strides, const float* img, const int64_t* img_strides,
    const int32_t* index_img, const int64_t* index_strides, const int32_t* vi, const int64_t* vi_strides,
    const float* grad_output, const int64_t* grad_output_strides, const float* bary_img,
    const int64_t* bary_strides, int64_t N, int64_t V, int64_t F, int64_t C, int64_t H, int64_t W,
    float max_dp_dr, float* grad_v_pix, void* workspace, size_t workspace_bytes, void* stream_) {
  if (!grad_v_pix && N * V > 0) return DRTK_B200_EINVAL;
  if (N * V == 0) return 0;
  return edge_launch(v_pix, v_strides, img, img_strides, index_img, index_strides, vi, vi_strides, grad_output,
                     grad_output_strides, N, V, F, C, H, W, max_dp_dr, nullptr, bary_img, bary_strides, grad_v_pix,
                     workspace, workspace_bytes, stream_);
}

extern "C" size_t drtk_b200_edge_grad_backward_fused_workspace_bytes(int64_t N, int64_t F) {
  return (N <= 0 || F <= 0) ? 0 : (size_t)(N * F) * 32 + 64;  // + slack to align the table to 32 B (256-bit loads) + item counter
}
