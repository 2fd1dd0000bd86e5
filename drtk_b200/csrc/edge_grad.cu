// edge_grad.cu -- backward of edge_grad_estimator: image-space gradients at visibility
// discontinuities -> per-pixel dL/d(v_pix_img) [N,3,H,W].
//
// Semantics: src/edge_grad/edge_grad_kernel.cu:217-449 (helpers :18-215) of the reference
// (paper: "Rasterized Edge Gradients: Handling Discontinuities Differentiably").
//
// The reference is a SCATTER: every pixel (x < W-1, y < H-1) acts as "centre", classifies the
// (centre,right) and (centre,down) pairs and atomically adds 9 values into a zero-initialised
// output (12 B/px memset + 9 REDG per pixel, interior pixels included).  Here the same sums are
// formed as a GATHER: the thread that owns output pixel p evaluates the (at most) four pairs p
// takes part in -- (p,right), (p,down) as centre and (left,p), (up,p) as neighbour -- and writes
// its three output values exactly once.  No memset, no atomics, deterministic; pairs whose two
// pixels show the same triangle (the vast majority) cost two integer compares.
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

// Contribution of the pair (centre=(xc,yc) showing triangle ci, neighbour=(xn,yn) showing ni) to
// ONE of its two pixels.  axis: 0 = horizontal pair (x gradient), 1 = vertical pair (y gradient).
// want_centre selects which side's (axis, z) contribution is returned.
__device__ __forceinline__ float2 pair_term(const EdgeArgs& a, int n, int ci, int ni, int xc, int yc,
                                            int xn, int yn, int axis, bool want_centre) {
  if (ci == ni) return make_float2(0.f, 0.f);  // lr_diff / ud_diff false (:304-306)
  const bool cv = ci >= 0, nv = ni >= 0;      // (:290-292)
  bool c_in_n = false, n_in_c = false;
  Tri2 tc, tn;
  if (cv && nv) {                              // (:320-325)
    fetch_tri(a, n, ci, tc);
    fetch_tri(a, n, ni, tn);
    c_in_n = pix_in_tri(tn, xc, yc);
    n_in_c = pix_in_tri(tc, xn, yn);
  }
  const bool inter = c_in_n && n_in_c;                    // (:334-335)
  if (!inter) {
    const bool adj = cv && nv && !c_in_n && !n_in_c;      // (:338-341)
    const bool c_over = c_in_n && !n_in_c;                // l_over_r / u_over_d (:328-331)
    const bool n_over = n_in_c && !c_in_n;                // r_over_l / d_over_u
    const bool zero = want_centre ? (!cv || n_over || adj) : (!nv || c_over || adj);  // (:392-393, :409-410)
    if (zero) return make_float2(0.f, 0.f);
    return make_float2(grad_dot(a, n, xc, yc, xn, yn), 0.f);
  }
  // intersection: both triangles valid (:394-406, :411-423)
  const float g = grad_dot(a, n, xc, yc, xn, yn);
  const float3 nc = tri_normal(a, n, tc), nn = tri_normal(a, n, tn);
  const float nca = axis == 0 ? nc.x : nc.y, nna = axis == 0 ? nn.x : nn.y;
  const float2 d = want_centre ? dp_dr(nca, nc.z, nna, nn.z, a.max_dp_dr)
                               : dp_dr(nna, nn.z, nca, nc.z, a.max_dp_dr);
  return make_float2(g * d.x, g * d.y);
}

__global__ void __launch_bounds__(256) edge_grad_bwd_kernel(EdgeArgs a, float* __restrict__ out) {
  const int64_t HW = (int64_t)a.H * a.W;
  const int64_t npix = (int64_t)a.N * HW;
  for (int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pix < npix;
       pix += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(pix / HW);
    const int64_t rem = pix - (int64_t)n * HW;
    const int y = (int)(rem / a.W), x = (int)(rem - (int64_t)y * a.W);
    const int32_t* ip = a.index_img + (int64_t)n * a.is.s0;
    const int id = ip[(int64_t)y * a.is.s1 + (int64_t)x * a.is.s2];
    const bool has_r = x < a.W - 1, has_d = y < a.H - 1, has_l = x > 0, has_u = y > 0;
    float gx = 0.f, gy = 0.f, gz_c = 0.f, gz_r = 0.f, gz_d = 0.f;
    // p as centre: only pixels with x < W-1 && y < H-1 act as centre (:270)
    if (has_r && has_d) {
      const int ir = ip[(int64_t)y * a.is.s1 + (int64_t)(x + 1) * a.is.s2];
      const int idn = ip[(int64_t)(y + 1) * a.is.s1 + (int64_t)x * a.is.s2];
      const float2 tx = pair_term(a, n, id, ir, x, y, x + 1, y, 0, true);
      const float2 ty = pair_term(a, n, id, idn, x, y, x, y + 1, 1, true);
      gx += tx.x; gy += ty.x; gz_c = tx.y + ty.y;
    }
    // p as right neighbour of (x-1, y): that centre must satisfy y < H-1
    if (has_l && has_d) {
      const int il = ip[(int64_t)y * a.is.s1 + (int64_t)(x - 1) * a.is.s2];
      const float2 t = pair_term(a, n, il, id, x - 1, y, x, y, 0, false);
      gx += t.x; gz_r = t.y;
    }
    // p as down neighbour of (x, y-1): that centre must satisfy x < W-1
    if (has_u && has_r) {
      const int iu = ip[(int64_t)(y - 1) * a.is.s1 + (int64_t)x * a.is.s2];
      const float2 t = pair_term(a, n, iu, id, x, y - 1, x, y, 1, false);
      gy += t.x; gz_d = t.y;
    }
    float* o = out + (int64_t)n * 3 * HW + rem;
    o[0] = -gx; o[HW] = -gy; o[2 * HW] = -(gz_c + gz_r + gz_d);  // negated sums (:431-445)
  }
}

}  // namespace
}  // namespace drtk

using namespace drtk;

extern "C" int drtk_b200_edge_grad_backward(const float* v_pix, const int64_t* v_strides, const float* img,
                                            const int64_t* img_strides, const int32_t* index_img,
                                            const int64_t* index_strides, const int32_t* vi,
                                            const int64_t* vi_strides, const float* grad_output,
                                            const int64_t* grad_output_strides, int64_t N, int64_t V,
                                            int64_t F, int64_t C, int64_t H, int64_t W, float max_dp_dr,
                                            float* grad_v_pix_img, void* stream_) {
  if (N < 0 || C < 0 || H < 0 || W < 0) return DRTK_B200_EINVAL;
  const int64_t npix = N * H * W;
  if (npix == 0) return 0;
  if (!v_pix || !img || !index_img || !vi || !grad_output || !grad_v_pix_img) return DRTK_B200_EINVAL;
  if (H > (1 << 30) || W > (1 << 30) || N > (1 << 30)) return DRTK_B200_EUNSUPPORTED;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  EdgeArgs a;
  a.v = v_pix; a.vs = make3(v_strides); a.img = img; a.ims = make4(img_strides);
  a.index_img = index_img; a.is = make3(index_strides); a.vi = vi; a.vis = make3(vi_strides);
  a.go = grad_output; a.gs = make4(grad_output_strides);
  a.N = (int)N; a.V = (int)V; a.F = (int)F; a.C = (int)C; a.H = (int)H; a.W = (int)W;
  a.max_dp_dr = max_dp_dr;
  const int64_t need = (npix + 255) / 256;
  const int64_t cap = (int64_t)kNumSMs * 8 * 8;
  edge_grad_bwd_kernel<<<(unsigned)(need < cap ? need : cap), 256, 0, stream>>>(a, grad_v_pix_img);
  DRTK_CHECK_LAUNCH();
  return 0;
}
