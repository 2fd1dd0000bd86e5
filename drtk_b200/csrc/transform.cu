// transform.cu -- world -> pixel projection of the vertex table, forward and backward, one kernel each.
//
// the step in front of the rasteriser.  The reference does it with ~20 stock torch ops
// forward and twice that backward (drtk/transform.py:68-119 -> drtk/utils/projection.py:486-646); here one
// thread owns a vertex, the per-camera block sits in registers, and the gradients of the camera
// parameters are reduced in-warp and added with 26 REDs per warp.
//
//   v_cam = R (v - c)                                  (projection.py:536)
//   p     = v_cam.xy / zs,  zs = z clamped away from 0 (projection.py:47-50)
//   q     = distort(p)      pinhole | radial-tangential (:56-136) | fisheye (:139-186) | fisheye62 (:189-276)
//   v_pix = (F q + pp, z)                              (projection.py:51, :646)
//
// Camera block per batch item (28 floats, 112 B, built by the host with one torch.cat so that autograd routes
// the block's gradient back to campos / camrot / focal / princpt / distortion_coeff):
//   [0:3] campos  [3:12] camrot (row major)  [12:16] focal (row major)  [16:18] princpt  [18:26] D  [26] fov  [27] pad
// This file is compiled WITHOUT --use_fast_math: the outputs are compared with torch's IEEE float ops.
#include "common.cuh"

namespace drtk {
namespace {

constexpr int kCam = 28;
constexpr int kNumGrad = 26;  // campos 3 + camrot 9 + focal 4 + princpt 2 + D 8

struct Cam {
  float c[3], R[9], F[4], pp[2], D[8], fov;
};

__device__ __forceinline__ Cam load_cam(const float* __restrict__ cam) {
  Cam k;
  const float4* c4 = reinterpret_cast<const float4*>(cam);
  float buf[kCam];
#pragma unroll
  for (int i = 0; i < kCam / 4; ++i) {
    const float4 t = __ldg(c4 + i);
    buf[4 * i] = t.x; buf[4 * i + 1] = t.y; buf[4 * i + 2] = t.z; buf[4 * i + 3] = t.w;
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) k.c[i] = buf[i];
#pragma unroll
  for (int i = 0; i < 9; ++i) k.R[i] = buf[3 + i];
#pragma unroll
  for (int i = 0; i < 4; ++i) k.F[i] = buf[12 + i];
  k.pp[0] = buf[16]; k.pp[1] = buf[17];
#pragma unroll
  for (int i = 0; i < 8; ++i) k.D[i] = buf[18 + i];
  k.fov = buf[26];
  return k;
}

__device__ __forceinline__ float zclamp(float z) { return z < 0.f ? fminf(z, -1e-8f) : fmaxf(z, 1e-8f); }

// everything the backward needs from the forward of one vertex
struct Fwd {
  float d[3];    // v - c
  float vc[3];   // camera space
  float zs;      // clamped z
  float p[2];    // normalised image plane
  float q[2];    // distorted
  bool culled;   // fisheye62: outside the valid radius -> z = -1
};

// ---- distortion models, forward ----------------------------------------------------------
// radial-tangential, D = (k1, k2, p1, p2, k3, k4, k5, k6), missing coefficients = 0 (projection.py:97-134)
__device__ __forceinline__ void rt_fwd(const Cam& k, const float px, const float py, float& qx, float& qy) {
  const float fov = k.fov;
  const float r2 = fminf(px * px + py * py, fov * fov);
  const float xc = fminf(fmaxf(px, -fov), fov), yc = fminf(fmaxf(py, -fov), fov);
  const float r4 = r2 * r2, r6 = r4 * r2;
  const float num = 1.f + k.D[0] * r2 + k.D[1] * r4 + k.D[4] * r6;
  const float den = 1.f + k.D[5] * r2 + k.D[6] * r4 + k.D[7] * r6;
  const float Rr = num / den;
  qx = px * Rr + 2.f * xc * yc * k.D[2] + r2 * k.D[3] + 2.f * k.D[3] * xc * xc;
  qy = py * Rr + 2.f * xc * yc * k.D[3] + r2 * k.D[2] + 2.f * k.D[2] * yc * yc;
}

// theta_d(theta) and its derivative for the two fisheye polynomials (projection.py:171-178, :246-254)
template <int NK>
__device__ __forceinline__ void fisheye_poly(const Cam& k, float th, float& thd, float& dthd) {
  const float t2 = th * th;
  float pw = t2, s = 1.f, ds = 1.f;
#pragma unroll
  for (int i = 0; i < NK; ++i) {
    s += k.D[i] * pw;
    ds += float(2 * i + 3) * k.D[i] * pw;
    pw *= t2;
  }
  thd = th * s;
  dthd = ds;
}

template <int NK>
__device__ __forceinline__ void fisheye_fwd(const Cam& k, float px, float py, float& qx, float& qy) {
  const float r = sqrtf(px * px + py * py);
  const float rc = fminf(fmaxf(r, 1e-8f), k.fov);
  float thd, dthd;
  fisheye_poly<NK>(k, atanf(rc), thd, dthd);
  const float sc = thd / fmaxf(rc, 1e-8f);
  qx = px * sc; qy = py * sc;
  if (NK == 6) {  // fisheye62: clamp, then tangential terms with D[6], D[7] (projection.py:259-270)
    const float xr = fminf(fmaxf(qx, -k.fov), k.fov), yr = fminf(fmaxf(qy, -k.fov), k.fov);
    const float rr2 = xr * xr + yr * yr;
    qx = xr + (2.f * xr * xr + rr2) * k.D[6] + 2.f * xr * yr * k.D[7];
    qy = yr + 2.f * xr * yr * k.D[6] + (2.f * yr * yr + rr2) * k.D[7];
  }
}

__device__ __forceinline__ Fwd forward_vertex(const Cam& k, int mode, int cull, float vx, float vy, float vz) {
  Fwd f;
  f.d[0] = vx - k.c[0]; f.d[1] = vy - k.c[1]; f.d[2] = vz - k.c[2];
#pragma unroll
  for (int i = 0; i < 3; ++i) f.vc[i] = k.R[3 * i] * f.d[0] + k.R[3 * i + 1] * f.d[1] + k.R[3 * i + 2] * f.d[2];
  f.zs = zclamp(f.vc[2]);
  f.p[0] = f.vc[0] / f.zs; f.p[1] = f.vc[1] / f.zs;
  switch (mode) {
    case DRTK_B200_DIST_RADIAL_TANGENTIAL: rt_fwd(k, f.p[0], f.p[1], f.q[0], f.q[1]); break;
    case DRTK_B200_DIST_FISHEYE: fisheye_fwd<4>(k, f.p[0], f.p[1], f.q[0], f.q[1]); break;
    case DRTK_B200_DIST_FISHEYE62: fisheye_fwd<6>(k, f.p[0], f.p[1], f.q[0], f.q[1]); break;
    default: f.q[0] = f.p[0]; f.q[1] = f.p[1];
  }
  // fisheye62 with a user-given fov: rays beyond the valid radius get z = -1 so the rasteriser culls every
  // triangle touching them (projection.py:624-644)
  f.culled = cull && mode == DRTK_B200_DIST_FISHEYE62 && sqrtf(f.p[0] * f.p[0] + f.p[1] * f.p[1]) > k.fov;
  return f;
}

__global__ void __launch_bounds__(256)
transform_fwd_kernel(const float* __restrict__ v, Strides3 vs, const float* __restrict__ cam,
                     const int32_t* __restrict__ modes, int mode, int cull, int V, float* __restrict__ v_pix,
                     float* __restrict__ v_cam) {
  const int n = blockIdx.y;
  const Cam k = load_cam(cam + size_t(n) * kCam);
  if (modes) mode = __ldg(modes + n);
  const float* vn = v + int64_t(n) * vs.s0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < V; i += gridDim.x * blockDim.x) {
    const float* pv = vn + int64_t(i) * vs.s1;
    const Fwd f = forward_vertex(k, mode, cull, __ldg(pv), __ldg(pv + vs.s2), __ldg(pv + 2 * vs.s2));
    const size_t o = (size_t(n) * V + i) * 3;
    v_pix[o] = k.F[0] * f.q[0] + k.F[1] * f.q[1] + k.pp[0];
    v_pix[o + 1] = k.F[2] * f.q[0] + k.F[3] * f.q[1] + k.pp[1];
    v_pix[o + 2] = f.culled ? -1.f : f.vc[2];
    if (v_cam) { v_cam[o] = f.vc[0]; v_cam[o + 1] = f.vc[1]; v_cam[o + 2] = f.vc[2]; }
  }
}

// ---- backward ----------------------------------------------------------------------------
// torch's clamp passes the gradient where lo <= x <= hi (bounds inclusive)
__device__ __forceinline__ float pass(bool c) { return c ? 1.f : 0.f; }

// g = dL/dq in, dL/dp out; gD accumulates dL/dD
__device__ __forceinline__ void rt_bwd(const Cam& k, float px, float py, float gx, float gy, float& gpx,
                                       float& gpy, float (&gD)[8]) {
  const float fov = k.fov, s = px * px + py * py, f2 = fov * fov;
  const float r2 = fminf(s, f2);
  const float xc = fminf(fmaxf(px, -fov), fov), yc = fminf(fmaxf(py, -fov), fov);
  const float mr = pass(s <= f2), mx = pass(px >= -fov && px <= fov), my = pass(py >= -fov && py <= fov);
  const float r4 = r2 * r2, r6 = r4 * r2;
  const float num = 1.f + k.D[0] * r2 + k.D[1] * r4 + k.D[4] * r6;
  const float den = 1.f + k.D[5] * r2 + k.D[6] * r4 + k.D[7] * r6;
  const float iden = 1.f / den, Rr = num * iden;
  const float dnum = k.D[0] + 2.f * k.D[1] * r2 + 3.f * k.D[4] * r4;
  const float dden = k.D[5] + 2.f * k.D[6] * r2 + 3.f * k.D[7] * r4;
  const float gR = gx * px + gy * py;
  const float p1 = k.D[2], p2 = k.D[3];
  const float gr2 = gR * (dnum - Rr * dden) * iden + gx * p2 + gy * p1;
  gpx = gx * Rr + mx * (gx * (2.f * p1 * yc + 4.f * p2 * xc) + gy * (2.f * p2 * yc)) + mr * gr2 * 2.f * px;
  gpy = gy * Rr + my * (gx * (2.f * p1 * xc) + gy * (2.f * p2 * xc + 4.f * p1 * yc)) + mr * gr2 * 2.f * py;
  const float a = gR * iden, b = -gR * Rr * iden;
  gD[0] += a * r2; gD[1] += a * r4; gD[4] += a * r6;
  gD[5] += b * r2; gD[6] += b * r4; gD[7] += b * r6;
  gD[2] += gx * 2.f * xc * yc + gy * (r2 + 2.f * yc * yc);
  gD[3] += gx * (r2 + 2.f * xc * xc) + gy * 2.f * xc * yc;
}

template <int NK>
__device__ __forceinline__ void fisheye_bwd(const Cam& k, float px, float py, float gx, float gy, float& gpx,
                                            float& gpy, float (&gD)[8]) {
  const float r = sqrtf(px * px + py * py);
  const float rc = fminf(fmaxf(r, 1e-8f), k.fov);
  const float th = atanf(rc);
  float thd, dthd;
  fisheye_poly<NK>(k, th, thd, dthd);
  const float rr = fmaxf(rc, 1e-8f);
  const float sc = thd / rr;
  if (NK == 6) {  // back through the tangential terms and the clamp of the scaled point
    const float fov = k.fov, qx = px * sc, qy = py * sc;
    const float xr = fminf(fmaxf(qx, -fov), fov), yr = fminf(fmaxf(qy, -fov), fov);
    const float rr2 = xr * xr + yr * yr, p0 = k.D[6], p1 = k.D[7];
    gD[6] += gx * (2.f * xr * xr + rr2) + gy * (2.f * xr * yr);
    gD[7] += gx * (2.f * xr * yr) + gy * (2.f * yr * yr + rr2);
    const float cross = 2.f * p0 * yr + 2.f * p1 * xr;
    const float hx = gx * (1.f + 6.f * p0 * xr + 2.f * p1 * yr) + gy * cross;
    const float hy = gx * cross + gy * (1.f + 2.f * p0 * xr + 6.f * p1 * yr);
    gx = hx * pass(qx >= -fov && qx <= fov);
    gy = hy * pass(qy >= -fov && qy <= fov);
  }
  const float gsc = gx * px + gy * py;
  const float gthd = gsc / rr;
  const float grr = -gsc * sc / rr;
  const float grc = gthd * dthd / (1.f + rc * rc) + grr * pass(rc >= 1e-8f);
  const float gr = grc * pass(r >= 1e-8f && r <= k.fov);
  const float gs = r > 0.f ? gr / r : 0.f;  // d sqrt(s) = ds / (2 r); the 2 cancels with d(s) = 2 p dp
  gpx = gx * sc + gs * px;
  gpy = gy * sc + gs * py;
  const float t2 = th * th;
  float pw = t2 * th;
#pragma unroll
  for (int i = 0; i < NK; ++i) { gD[i] += gthd * pw; pw *= t2; }
}

__global__ void __launch_bounds__(256)
transform_bwd_kernel(const float* __restrict__ v, Strides3 vs, const float* __restrict__ cam,
                     const int32_t* __restrict__ modes, int mode, int cull, int V,
                     const float* __restrict__ g_pix, Strides3 gps, const float* __restrict__ g_cam, Strides3 gcs,
                     float* __restrict__ grad_v, float* __restrict__ grad_cam) {
  const int n = blockIdx.y;
  const Cam k = load_cam(cam + size_t(n) * kCam);
  if (modes) mode = __ldg(modes + n);
  const float* vn = v + int64_t(n) * vs.s0;
  float acc[kNumGrad];
#pragma unroll
  for (int i = 0; i < kNumGrad; ++i) acc[i] = 0.f;
  float (&gc)[3] = *reinterpret_cast<float (*)[3]>(acc);
  float (&gR)[9] = *reinterpret_cast<float (*)[9]>(acc + 3);
  float (&gF)[4] = *reinterpret_cast<float (*)[4]>(acc + 12);
  float (&gpp)[2] = *reinterpret_cast<float (*)[2]>(acc + 16);
  float (&gD)[8] = *reinterpret_cast<float (*)[8]>(acc + 18);

  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < V; i += gridDim.x * blockDim.x) {
    const float* pv = vn + int64_t(i) * vs.s1;
    const Fwd f = forward_vertex(k, mode, cull, __ldg(pv), __ldg(pv + vs.s2), __ldg(pv + 2 * vs.s2));
    float gx = 0.f, gy = 0.f, gz = 0.f;
    if (g_pix) {
      const float* pg = g_pix + int64_t(n) * gps.s0 + int64_t(i) * gps.s1;
      gx = __ldg(pg); gy = __ldg(pg + gps.s2); gz = __ldg(pg + 2 * gps.s2);
    }
    // pix = F q + pp
    gF[0] += gx * f.q[0]; gF[1] += gx * f.q[1]; gF[2] += gy * f.q[0]; gF[3] += gy * f.q[1];
    gpp[0] += gx; gpp[1] += gy;
    const float gqx = k.F[0] * gx + k.F[2] * gy, gqy = k.F[1] * gx + k.F[3] * gy;
    float gpx, gpy;
    switch (mode) {
      case DRTK_B200_DIST_RADIAL_TANGENTIAL: rt_bwd(k, f.p[0], f.p[1], gqx, gqy, gpx, gpy, gD); break;
      case DRTK_B200_DIST_FISHEYE: fisheye_bwd<4>(k, f.p[0], f.p[1], gqx, gqy, gpx, gpy, gD); break;
      case DRTK_B200_DIST_FISHEYE62: fisheye_bwd<6>(k, f.p[0], f.p[1], gqx, gqy, gpx, gpy, gD); break;
      default: gpx = gqx; gpy = gqy;
    }
    // p = vc.xy / zs ; zs passes the gradient where |z| >= 1e-8
    const float iz = 1.f / f.zs;
    float gvc[3];
    gvc[0] = gpx * iz; gvc[1] = gpy * iz;
    const float z = f.vc[2];
    const float gzs = -(gpx * f.p[0] + gpy * f.p[1]) * iz;
    gvc[2] = gzs * pass(z < 0.f ? z <= -1e-8f : z >= 1e-8f) + (f.culled ? 0.f : gz);
    if (g_cam) {
      const float* pg = g_cam + int64_t(n) * gcs.s0 + int64_t(i) * gcs.s1;
      gvc[0] += __ldg(pg); gvc[1] += __ldg(pg + gcs.s2); gvc[2] += __ldg(pg + 2 * gcs.s2);
    }
    // vc = R d, d = v - c
    float gd[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) gd[j] = k.R[j] * gvc[0] + k.R[3 + j] * gvc[1] + k.R[6 + j] * gvc[2];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) gR[3 * a + b] += gvc[a] * f.d[b];
#pragma unroll
    for (int j = 0; j < 3; ++j) gc[j] -= gd[j];
    if (grad_v) {
      const size_t o = (size_t(n) * V + i) * 3;
      grad_v[o] = gd[0]; grad_v[o + 1] = gd[1]; grad_v[o + 2] = gd[2];
    }
  }
  if (!grad_cam) return;
  // warp butterfly, then one RED per (warp, parameter)
#pragma unroll
  for (int i = 0; i < kNumGrad; ++i) {
    float x = acc[i];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
    acc[i] = x;
  }
  if ((threadIdx.x & 31) == 0) {
    float* out = grad_cam + size_t(n) * kCam;
#pragma unroll
    for (int i = 0; i < kNumGrad; ++i) atomicAdd(out + i, acc[i]);
  }
}

inline dim3 grid_for(int64_t N, int64_t V) {
  // a handful of CTAs per SM over the whole batch; every thread loops over its vertices
  int64_t bx = (V + 255) / 256;
  const int64_t cap = (int64_t(num_sms()) * 8 + N - 1) / N;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  return dim3(unsigned(bx), unsigned(N));
}

inline bool bad_mode(int mode) { return mode < DRTK_B200_DIST_PINHOLE || mode > DRTK_B200_DIST_FISHEYE62; }

}  // namespace
}  // namespace drtk

using namespace drtk;

extern "C" int drtk_b200_transform_forward(const float* v, const int64_t* v_strides, const float* cam,
                                           const int32_t* modes, int mode, int cull_outside_fov, int64_t N,
                                           int64_t V, float* v_pix, float* v_cam, void* stream) {
  if (!v || !v_strides || !cam || !v_pix || N < 0 || V < 0 || bad_mode(mode)) return DRTK_B200_EINVAL;
  if (N > 65535 || V > INT32_MAX) return DRTK_B200_EUNSUPPORTED;
  if (N == 0 || V == 0) return 0;
  transform_fwd_kernel<<<grid_for(N, V), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      v, make3(v_strides), cam, modes, mode, cull_outside_fov, int(V), v_pix, v_cam);
  DRTK_CHECK_LAUNCH();
  return 0;
}

extern "C" int drtk_b200_transform_backward(const float* v, const int64_t* v_strides, const float* cam,
                                            const int32_t* modes, int mode, int cull_outside_fov,
                                            const float* grad_v_pix, const int64_t* grad_v_pix_strides,
                                            const float* grad_v_cam, const int64_t* grad_v_cam_strides, int64_t N,
                                            int64_t V, float* grad_v, float* grad_cam, void* stream) {
  if (!v || !v_strides || !cam || N < 0 || V < 0 || bad_mode(mode)) return DRTK_B200_EINVAL;
  if ((grad_v_pix && !grad_v_pix_strides) || (grad_v_cam && !grad_v_cam_strides)) return DRTK_B200_EINVAL;
  if (N > 65535 || V > INT32_MAX) return DRTK_B200_EUNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (grad_cam && N > 0) DRTK_CUDA(cudaMemsetAsync(grad_cam, 0, size_t(N) * kCam * sizeof(float), st));
  if (N == 0 || V == 0 || (!grad_v && !grad_cam)) return 0;
  const Strides3 z{0, 0, 0};
  transform_bwd_kernel<<<grid_for(N, V), 256, 0, st>>>(
      v, make3(v_strides), cam, modes, mode, cull_outside_fov, int(V), grad_v_pix,
      grad_v_pix ? make3(grad_v_pix_strides) : z, grad_v_cam, grad_v_cam ? make3(grad_v_cam_strides) : z, grad_v,
      grad_cam);
  DRTK_CHECK_LAUNCH();
  return 0;
}
