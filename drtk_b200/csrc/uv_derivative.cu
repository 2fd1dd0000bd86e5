// uv_derivative.cu -- screen_space_uv_derivative in one kernel: the per-pixel 2x2 Jacobian d(uv)/d(pixel) that
// mipmap_grid_sample takes as `vt_dxdy_img`.
//
// The reference composes it from stock ops (drtk/screen_space_uv_derivative.py:16-80): face_dpdt
// (drtk/utils/geometry.py:18-82) -> two `interpolate` calls over a 3F-vertex "discontinuous" mesh (6 + 3 channels
// written and re-read) -> project_points_grad (drtk/utils/projection.py:649-706) -> 2x2 inverse -> mask; its own
// comment (:37-38) asks for a CUDA kernel.  Here one thread owns a pixel: 17 B read (index, bary, mask), 16 B
// written, the triangle's vertex / uv rows come from L2.
//
//   (dp/dt)^T = ((dt/db)^T)^-1 (dp/db)^T                 per triangle, 2x3
//   p         = sum_k b_k p_k ;  (dp/dt)^T scaled by sum_k b_k   (what interpolating per-face constants yields)
//   M[i][j]   = d pix_j / d t_i = focal * quotient rule of (R dp_i, R (p - c))        pinhole camera only
//   out       = M^-1, zero where mask is false or no triangle covers the pixel
// Compiled without --use_fast_math (compared with torch's IEEE float ops).
#include "common.cuh"

namespace drtk {
namespace {

struct UvArgs {
  const float* v; Strides3 vs;        // [N,V,3]
  const float* vt; Strides3 ts;       // [N,T,2]
  const int32_t* vi; int64_t vis[2];  // [F,3]
  const int32_t* vti; int64_t vtis[2];
  const int32_t* index; Strides3 xs;  // [N,H,W]
  const float* bary; Strides4 bs;     // [N,3,H,W]
  const uint8_t* mask; Strides3 ms;   // [N,H,W] bool
  const float* cam;                   // [N,16]: campos 3, camrot 9, focal 4
  int N, H, W;
};

__global__ void __launch_bounds__(256) uv_derivative_kernel(UvArgs a, float4* __restrict__ out) {
  const int64_t HW = int64_t(a.H) * a.W, idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (idx >= a.N * HW) return;
  const int n = int(idx / HW), h = int((idx % HW) / a.W), w = int(idx % a.W);
  float4 res = make_float4(0.f, 0.f, 0.f, 0.f);
  const int t = a.index[n * a.xs.s0 + h * a.xs.s1 + w * a.xs.s2];
  if (t >= 0 && a.mask[n * a.ms.s0 + h * a.ms.s1 + w * a.ms.s2]) {
    float p[3][3], uv[3][2];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float* pv = a.v + n * a.vs.s0 + int64_t(a.vi[t * a.vis[0] + k * a.vis[1]]) * a.vs.s1;
      const float* pt = a.vt + n * a.ts.s0 + int64_t(a.vti[t * a.vtis[0] + k * a.vtis[1]]) * a.ts.s1;
      p[k][0] = __ldg(pv); p[k][1] = __ldg(pv + a.vs.s2); p[k][2] = __ldg(pv + 2 * a.vs.s2);
      uv[k][0] = __ldg(pt); uv[k][1] = __ldg(pt + a.ts.s2);
    }
    // (dt/db)^T = [[a00 a01],[a10 a11]], rows = uv1-uv0, uv2-uv0 ; inverse by the adjugate
    const float a00 = uv[1][0] - uv[0][0], a01 = uv[1][1] - uv[0][1], a10 = uv[2][0] - uv[0][0], a11 = uv[2][1] - uv[0][1];
    const float idet = 1.f / (a00 * a11 - a01 * a10);
    const float i00 = a11 * idet, i01 = -a01 * idet, i10 = -a10 * idet, i11 = a00 * idet;
    const float* pb = a.bary + n * a.bs.s0 + h * a.bs.s2 + w * a.bs.s3;
    const float b0 = __ldg(pb), b1 = __ldg(pb + a.bs.s1), b2 = __ldg(pb + 2 * a.bs.s1);
    const float bsum = b0 + b1 + b2;
    const float* cam = a.cam + n * 16;
    float d[2][3], pos[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float e1 = p[1][j] - p[0][j], e2 = p[2][j] - p[0][j];
      d[0][j] = (i00 * e1 + i01 * e2) * bsum;
      d[1][j] = (i10 * e1 + i11 * e2) * bsum;
      pos[j] = (p[0][j] * b0 + p[1][j] * b1 + p[2][j] * b2) - __ldg(cam + j);
    }
    const float* R = cam + 3;
    float pc[3], dc[2][3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float r0 = __ldg(R + 3 * r), r1 = __ldg(R + 3 * r + 1), r2 = __ldg(R + 3 * r + 2);
      pc[r] = r0 * pos[0] + r1 * pos[1] + r2 * pos[2];
      dc[0][r] = r0 * d[0][0] + r1 * d[0][1] + r2 * d[0][2];
      dc[1][r] = r0 * d[1][0] + r1 * d[1][1] + r2 * d[1][2];
    }
    const float z = pc[2] < 0.f ? fminf(pc[2], -1e-8f) : fmaxf(pc[2], 1e-8f);
    const float iz2 = 1.f / (z * z);
    const float f00 = __ldg(cam + 12), f01 = __ldg(cam + 13), f10 = __ldg(cam + 14), f11 = __ldg(cam + 15);
    float m[2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float qx = (dc[i][0] * z - pc[0] * dc[i][2]) * iz2, qy = (dc[i][1] * z - pc[1] * dc[i][2]) * iz2;
      m[i][0] = f00 * qx + f01 * qy;
      m[i][1] = f10 * qx + f11 * qy;
    }
    const float im = 1.f / (m[0][0] * m[1][1] - m[0][1] * m[1][0]);
    res = make_float4(m[1][1] * im, -m[0][1] * im, -m[1][0] * im, m[0][0] * im);
  }
  out[idx] = res;
}

}  // namespace
}  // namespace drtk

using namespace drtk;

extern "C" int drtk_b200_screen_space_uv_derivative(
    const float* v, const int64_t* v_strides, const float* vt, const int64_t* vt_strides, const int32_t* vi,
    const int64_t* vi_strides, const int32_t* vti, const int64_t* vti_strides, const int32_t* index_img,
    const int64_t* index_strides, const float* bary_img, const int64_t* bary_strides, const uint8_t* mask,
    const int64_t* mask_strides, const float* cam, int64_t N, int64_t H, int64_t W, float* out, void* stream) {
  if (!v_strides || !vt_strides || !vi_strides || !vti_strides || !index_strides || !bary_strides || !mask_strides ||
      N < 0 || H < 0 || W < 0)
    return DRTK_B200_EINVAL;
  if (N > INT32_MAX || H > INT32_MAX || W > INT32_MAX) return DRTK_B200_EUNSUPPORTED;
  const int64_t total = N * H * W;
  if (total == 0) return 0;
  if (!v || !vt || !vi || !vti || !index_img || !bary_img || !mask || !cam || !out) return DRTK_B200_EINVAL;
  UvArgs a;
  a.v = v; a.vs = make3(v_strides); a.vt = vt; a.ts = make3(vt_strides);
  a.vi = vi; a.vis[0] = vi_strides[0]; a.vis[1] = vi_strides[1];
  a.vti = vti; a.vtis[0] = vti_strides[0]; a.vtis[1] = vti_strides[1];
  a.index = index_img; a.xs = make3(index_strides); a.bary = bary_img; a.bs = make4(bary_strides);
  a.mask = mask; a.ms = make3(mask_strides); a.cam = cam;
  a.N = int(N); a.H = int(H); a.W = int(W);
  uv_derivative_kernel<<<unsigned((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      a, reinterpret_cast<float4*>(out));
  DRTK_CHECK_LAUNCH();
  return 0;
}
