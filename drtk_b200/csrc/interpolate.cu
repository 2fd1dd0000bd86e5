// interpolate.cu -- barycentric interpolation of vertex attributes, forward and backward.
//
// Semantics: src/interpolate/interpolate_kernel.cu:38-111 (forward) and :113-299 (backward) of
// the reference; fp32 results within 1e-5 relative.
//
// Forward  (16 + 4C B/px, write dominated): one thread owns four adjacent pixels and walks the
// channels four at a time: per 4-channel group 3 x LDG.128 per distinct triangle (attribute rows
// are 16-B aligned when C % 4 == 0) and 4 x STG.128 into the planar output, so every warp store
// is a full 512-B span of one channel plane.
//
// Backward (16 + 4C B/px read, +12 B/px written when bary_img needs grad):
//   bary grad  : direct per-pixel dot products.
//   vertex grad: lanes of a warp are consecutive pixels; runs of equal (image, triangle) are
//                reduced with a segmented shuffle scan, 4 channels x 3 vertices at a time, and the
//                head lane of each run issues one 128-bit vector reduction per vertex
//                (red.global.add.v4.f32) -- no shared memory, no block barriers (the reference
//                needs one barrier and up to three scalar atomics per channel per run).
#include "common.cuh"

namespace drtk {
namespace {

struct InterpArgs {
  const float* attr;
  Strides3 as;
  const int32_t* vi;
  Strides3 vis;
  const int32_t* index_img;
  Strides3 is;
  const float* bary;
  Strides4 bs;
  int N, V, F, C, H, W;
};

__device__ __forceinline__ void load_vi(const InterpArgs& a, int n, int t, int& i0, int& i1, int& i2) {
  const int32_t* vip = a.vi + (int64_t)n * a.vis.s0 + (int64_t)t * a.vis.s1;
  i0 = vip[0]; i1 = vip[a.vis.s2]; i2 = vip[2 * a.vis.s2];
}

__device__ __forceinline__ float sweep_x(int w, int W) { return ((float)w * 2.0f + 1.0f) / (float)W - 1.0f; }

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
// VEC: index/bary images dense & 16-B aligned along W (4 px per thread).
// AVEC: attribute rows 16-B aligned and C % 4 == 0 (float4 gathers).
template <bool VEC, bool AVEC>
__global__ void __launch_bounds__(256) interp_fwd_kernel(InterpArgs a, float* __restrict__ out) {
  const int64_t HW = (int64_t)a.H * a.W;
  constexpr int PX = VEC ? 4 : 1;
  const int64_t ngroups = (int64_t)a.N * HW / PX;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < ngroups;
       q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = q * PX;
    const int n = (int)(pix / HW);
    const int64_t rem = pix - (int64_t)n * HW;
    const int h = (int)(rem / a.W), w = (int)(rem - (int64_t)h * a.W);
    int ids[PX];
    float b0[PX], b1[PX], b2[PX];
    const int32_t* ip = a.index_img + (int64_t)n * a.is.s0 + (int64_t)h * a.is.s1 + (int64_t)w * a.is.s2;
    const float* bp = a.bary + (int64_t)n * a.bs.s0 + (int64_t)h * a.bs.s2 + (int64_t)w * a.bs.s3;
    if (VEC) {
      const int4 id4 = ldg_stream_i4(ip);
      ids[0] = id4.x; ids[PX > 1 ? 1 : 0] = id4.y; ids[PX > 2 ? 2 : 0] = id4.z; ids[PX > 3 ? 3 : 0] = id4.w;
      const float4 x0 = ldg_stream_f4(bp), x1 = ldg_stream_f4(bp + a.bs.s1), x2 = ldg_stream_f4(bp + 2 * a.bs.s1);
      b0[0] = x0.x; b0[PX > 1 ? 1 : 0] = x0.y; b0[PX > 2 ? 2 : 0] = x0.z; b0[PX > 3 ? 3 : 0] = x0.w;
      b1[0] = x1.x; b1[PX > 1 ? 1 : 0] = x1.y; b1[PX > 2 ? 2 : 0] = x1.z; b1[PX > 3 ? 3 : 0] = x1.w;
      b2[0] = x2.x; b2[PX > 1 ? 1 : 0] = x2.y; b2[PX > 2 ? 2 : 0] = x2.z; b2[PX > 3 ? 3 : 0] = x2.w;
    } else {
      ids[0] = ip[0];
      b0[0] = bp[0]; b1[0] = bp[a.bs.s1]; b2[0] = bp[2 * a.bs.s1];
    }
    int i0[PX], i1[PX], i2[PX];
#pragma unroll
    for (int j = 0; j < PX; ++j) {
      i0[j] = i1[j] = i2[j] = 0;
      if (ids[j] != -1) {
        if (j > 0 && ids[j] == ids[j - 1]) { i0[j] = i0[j - 1]; i1[j] = i1[j - 1]; i2[j] = i2[j - 1]; }
        else load_vi(a, n, ids[j], i0[j], i1[j], i2[j]);
      }
    }
    const float sy = sweep_x(h, a.H);
    const float* an = a.attr + (int64_t)n * a.as.s0;
    float* op = out + (int64_t)n * a.C * HW + rem;

    if (AVEC) {
      for (int c = 0; c < a.C; c += 4) {
        float r[4][PX];  // [channel][pixel]
        float4 A0, A1, A2;
#pragma unroll
        for (int j = 0; j < PX; ++j) {
          if (ids[j] != -1) {
            if (!(j > 0 && ids[j] == ids[j - 1])) {
              A0 = *reinterpret_cast<const float4*>(an + (int64_t)i0[j] * a.as.s1 + c);
              A1 = *reinterpret_cast<const float4*>(an + (int64_t)i1[j] * a.as.s1 + c);
              A2 = *reinterpret_cast<const float4*>(an + (int64_t)i2[j] * a.as.s1 + c);
            }
            r[0][j] = A0.x * b0[j] + A1.x * b1[j] + A2.x * b2[j];  // (:102)
            r[1][j] = A0.y * b0[j] + A1.y * b1[j] + A2.y * b2[j];
            r[2][j] = A0.z * b0[j] + A1.z * b1[j] + A2.z * b2[j];
            r[3][j] = A0.w * b0[j] + A1.w * b1[j] + A2.w * b2[j];
          } else {  // coordinate sweep for empty pixels (:104-109): even channel -> x, odd -> y
            const float sx = sweep_x(w + j, a.W);
            r[0][j] = sx; r[1][j] = sy; r[2][j] = sx; r[3][j] = sy;
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float* o = op + (int64_t)(c + k) * HW;
          if (VEC) stg_stream_f4(o, make_float4(r[k][0], r[k][PX > 1 ? 1 : 0], r[k][PX > 2 ? 2 : 0], r[k][PX > 3 ? 3 : 0]));
          else o[0] = r[k][0];
        }
      }
    } else {
      for (int c = 0; c < a.C; ++c) {
        float r[PX];
#pragma unroll
        for (int j = 0; j < PX; ++j) {
          if (ids[j] != -1) {
            const float v0 = an[(int64_t)i0[j] * a.as.s1 + (int64_t)c * a.as.s2];
            const float v1 = an[(int64_t)i1[j] * a.as.s1 + (int64_t)c * a.as.s2];
            const float v2 = an[(int64_t)i2[j] * a.as.s1 + (int64_t)c * a.as.s2];
            r[j] = v0 * b0[j] + v1 * b1[j] + v2 * b2[j];
          } else {
            r[j] = (c & 1) ? sy : sweep_x(w + j, a.W);
          }
        }
        float* o = op + (int64_t)c * HW;
        if (VEC) stg_stream_f4(o, make_float4(r[0], r[PX > 1 ? 1 : 0], r[PX > 2 ? 2 : 0], r[PX > 3 ? 3 : 0]));
        else o[0] = r[0];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------
struct InterpBwdArgs {
  InterpArgs f;
  const float* grad_out;
  Strides4 gs;
};

// One thread per pixel.  NEED_VERT: accumulate vertex-attribute gradients; NEED_BARY: write
// the barycentric gradient image.  RV4: vert_grad rows are 16-B aligned (C % 4 == 0) so run
// heads can use 128-bit vector reductions.
template <bool NEED_VERT, bool NEED_BARY, bool RV4>
__global__ void __launch_bounds__(256) interp_bwd_kernel(InterpBwdArgs b, float* __restrict__ vert_grad,
                                                         float* __restrict__ bary_grad) {
  const InterpArgs& a = b.f;
  const int64_t HW = (int64_t)a.H * a.W;
  const int64_t npix = (int64_t)a.N * HW;
  const int lane = threadIdx.x & 31;
  const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_range = pix < npix;
  int n = 0, h = 0, w = 0, id = -1;
  int64_t rem = 0;
  if (in_range) {
    n = (int)(pix / HW);
    rem = pix - (int64_t)n * HW;
    h = (int)(rem / a.W); w = (int)(rem - (int64_t)h * a.W);
    id = a.index_img[(int64_t)n * a.is.s0 + (int64_t)h * a.is.s1 + (int64_t)w * a.is.s2];
  }
  const bool used = id != -1;
  const bool warp_used = __any_sync(0xffffffffu, used);
  if (!warp_used) {
    if (NEED_BARY && in_range) {  // every pixel of bary_grad is written (:282-297)
      float* g = bary_grad + (int64_t)n * 3 * HW + rem;
      g[0] = 0.f; g[HW] = 0.f; g[2 * HW] = 0.f;
    }
    return;
  }
  int i0 = 0, i1 = 0, i2 = 0;
  float b0 = 0.f, b1 = 0.f, b2 = 0.f;
  if (used) {
    load_vi(a, n, id, i0, i1, i2);
    if (NEED_VERT) {
      const float* bp = a.bary + (int64_t)n * a.bs.s0 + (int64_t)h * a.bs.s2 + (int64_t)w * a.bs.s3;
      b0 = ldg_stream_f(bp); b1 = ldg_stream_f(bp + a.bs.s1); b2 = ldg_stream_f(bp + 2 * a.bs.s1);
    }
  }
  // runs of equal (image, triangle) along the warp
  const int64_t key = used ? (((int64_t)n << 32) | (uint32_t)id) : ((int64_t)-1 - lane);
  const int64_t key_up = __shfl_up_sync(0xffffffffu, key, 1);
  const int64_t key_dn = __shfl_down_sync(0xffffffffu, key, 1);
  const bool head = (lane == 0) || (key_up != key);
  const bool tail = (lane == 31) || (key_dn != key);
  const unsigned tail_mask = __ballot_sync(0xffffffffu, tail);

  const float* gp = b.grad_out + (int64_t)n * b.gs.s0 + (int64_t)h * b.gs.s2 + (int64_t)w * b.gs.s3;
  const float* an = a.attr + (int64_t)n * a.as.s0;
  const float* a0 = an + (int64_t)i0 * a.as.s1;
  const float* a1 = an + (int64_t)i1 * a.as.s1;
  const float* a2 = an + (int64_t)i2 * a.as.s1;
  float* vg = NEED_VERT ? vert_grad + (int64_t)n * a.V * a.C : nullptr;
  float gb0 = 0.f, gb1 = 0.f, gb2 = 0.f;

  for (int c = 0; c < a.C; c += 4) {
    float g[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      g[k] = (in_range && c + k < a.C) ? ldg_stream_f(gp + (int64_t)(c + k) * b.gs.s1) : 0.f;
    if (NEED_BARY && used) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (c + k < a.C) {
          gb0 += g[k] * a0[(int64_t)(c + k) * a.as.s2];  // (:256-258)
          gb1 += g[k] * a1[(int64_t)(c + k) * a.as.s2];
          gb2 += g[k] * a2[(int64_t)(c + k) * a.as.s2];
        }
      }
    }
    if (NEED_VERT) {
      float s[12];
#pragma unroll
      for (int k = 0; k < 4; ++k) { s[k] = g[k] * b0; s[4 + k] = g[k] * b1; s[8 + k] = g[k] * b2; }  // (:262-267)
      seg_reduce_to_head<12>(s, tail_mask, lane);
      if (head && used) {
        float* r0 = vg + (int64_t)i0 * a.C + c;
        float* r1 = vg + (int64_t)i1 * a.C + c;
        float* r2 = vg + (int64_t)i2 * a.C + c;
        if (RV4) {
          red_add_v4(r0, s[0], s[1], s[2], s[3]);
          red_add_v4(r1, s[4], s[5], s[6], s[7]);
          red_add_v4(r2, s[8], s[9], s[10], s[11]);
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (c + k < a.C) { red_add(r0 + k, s[k]); red_add(r1 + k, s[4 + k]); red_add(r2 + k, s[8 + k]); }
        }
      }
    }
  }
  if (NEED_BARY && in_range) {
    float* g = bary_grad + (int64_t)n * 3 * HW + rem;
    g[0] = gb0; g[HW] = gb1; g[2 * HW] = gb2;  // zeros for empty pixels
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

static int fill_args(InterpArgs& a, const float* attr, const int64_t* attr_strides, const int32_t* vi,
                     const int64_t* vi_strides, const int32_t* index_img, const int64_t* index_strides,
                     const float* bary_img, const int64_t* bary_strides, int64_t N, int64_t V, int64_t F,
                     int64_t C, int64_t H, int64_t W) {
  if (!attr || !vi || !index_img || !bary_img || !attr_strides || !vi_strides || !index_strides ||
      !bary_strides)
    return DRTK_B200_EINVAL;
  if (H > (1 << 30) || W > (1 << 30) || N > (1 << 30) || C > (1 << 20)) return DRTK_B200_EUNSUPPORTED;
  a.attr = attr; a.as = make3(attr_strides); a.vi = vi; a.vis = make3(vi_strides);
  a.index_img = index_img; a.is = make3(index_strides); a.bary = bary_img; a.bs = make4(bary_strides);
  a.N = (int)N; a.V = (int)V; a.F = (int)F; a.C = (int)C; a.H = (int)H; a.W = (int)W;
  return 0;
}

extern "C" int drtk_b200_interpolate_forward(const float* vert_attributes, const int64_t* attr_strides,
                                             const int32_t* vi, const int64_t* vi_strides,
                                             const int32_t* index_img, const int64_t* index_strides,
                                             const float* bary_img, const int64_t* bary_strides,
                                             int64_t N, int64_t V, int64_t F, int64_t C, int64_t H,
                                             int64_t W, float* out, void* stream_) {
  if (N < 0 || C < 0 || H < 0 || W < 0) return DRTK_B200_EINVAL;
  if (N * C * H * W == 0) return 0;
  if (!out) return DRTK_B200_EINVAL;
  InterpArgs a;
  const int rc = fill_args(a, vert_attributes, attr_strides, vi, vi_strides, index_img, index_strides,
                           bary_img, bary_strides, N, V, F, C, H, W);
  if (rc) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool vec = VecOk::image(index_img, W, a.is.s2, a.is.s1, a.is.s0) &&
                   VecOk::image(bary_img, W, a.bs.s3, a.bs.s2, a.bs.s1, a.bs.s0);
  const bool avec = (C % 4 == 0) && a.as.s2 == 1 && (a.as.s1 % 4 == 0) && (a.as.s0 % 4 == 0) &&
                    (reinterpret_cast<uintptr_t>(vert_attributes) % 16 == 0);
  const int64_t npix = N * H * W;
  if (vec && avec)
    interp_fwd_kernel<true, true><<<grid_for(npix / 4, 256, 8), 256, 0, stream>>>(a, out);
  else if (vec)
    interp_fwd_kernel<true, false><<<grid_for(npix / 4, 256, 8), 256, 0, stream>>>(a, out);
  else if (avec)
    interp_fwd_kernel<false, true><<<grid_for(npix, 256, 8), 256, 0, stream>>>(a, out);
  else
    interp_fwd_kernel<false, false><<<grid_for(npix, 256, 8), 256, 0, stream>>>(a, out);
  DRTK_CHECK_LAUNCH();
  return 0;
}

extern "C" int drtk_b200_interpolate_backward(
    const float* grad_out, const int64_t* grad_out_strides, const float* vert_attributes,
    const int64_t* attr_strides, const int32_t* vi, const int64_t* vi_strides, const int32_t* index_img,
    const int64_t* index_strides, const float* bary_img, const int64_t* bary_strides, int64_t N,
    int64_t V, int64_t F, int64_t C, int64_t H, int64_t W, float* vert_attributes_grad,
    float* bary_img_grad, void* stream_) {
  if (N < 0 || V < 0 || C < 0 || H < 0 || W < 0) return DRTK_B200_EINVAL;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (vert_attributes_grad && N * V * C > 0)
    DRTK_CUDA(cudaMemsetAsync(vert_attributes_grad, 0, sizeof(float) * (size_t)(N * V * C), stream));  // (:661)
  const int64_t npix = N * H * W;
  if (npix == 0) return 0;
  if (!vert_attributes_grad && !bary_img_grad) return 0;
  if (C == 0) {
    if (bary_img_grad) DRTK_CUDA(cudaMemsetAsync(bary_img_grad, 0, sizeof(float) * (size_t)(npix * 3), stream));
    return 0;
  }
  if (!grad_out || !grad_out_strides) return DRTK_B200_EINVAL;
  InterpBwdArgs b;
  const int rc = fill_args(b.f, vert_attributes, attr_strides, vi, vi_strides, index_img, index_strides,
                           bary_img, bary_strides, N, V, F, C, H, W);
  if (rc) return rc;
  b.grad_out = grad_out; b.gs = make4(grad_out_strides);
  const unsigned blocks = (unsigned)((npix + 255) / 256);
  const bool rv4 = (C % 4 == 0) && (reinterpret_cast<uintptr_t>(vert_attributes_grad) % 16 == 0);
  const bool nv = vert_attributes_grad != nullptr, nb = bary_img_grad != nullptr;
  if (nv && nb) {
    if (rv4) interp_bwd_kernel<true, true, true><<<blocks, 256, 0, stream>>>(b, vert_attributes_grad, bary_img_grad);
    else interp_bwd_kernel<true, true, false><<<blocks, 256, 0, stream>>>(b, vert_attributes_grad, bary_img_grad);
  } else if (nv) {
    if (rv4) interp_bwd_kernel<true, false, true><<<blocks, 256, 0, stream>>>(b, vert_attributes_grad, nullptr);
    else interp_bwd_kernel<true, false, false><<<blocks, 256, 0, stream>>>(b, vert_attributes_grad, nullptr);
  } else {
    interp_bwd_kernel<false, true, false><<<blocks, 256, 0, stream>>>(b, nullptr, bary_img_grad);
  }
  DRTK_CHECK_LAUNCH();
  return 0;
}
