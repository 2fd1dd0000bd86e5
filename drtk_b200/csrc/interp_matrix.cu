// interp_matrix.cu -- sparse interpolation matrices of a fixed rasterisation.
//
// Semantics: src/interpolate/interpolate_kernel.cu:301-452 of the reference (kernels) and :699-900 (launchers):
//   interpolation_matrix           A  [num_valid_pixels, V] CSR: per foreground pixel three entries (columns =
//                                  the triangle's vertex ids sorted ascending, values = the matching barycentrics)
//   interpolation_normal_matrix    values of A^T A on a CSR structure that depends on the topology only; for every
//                                  foreground pixel the nine products b_i * b_j go to the slots pair[tri][i*3+j]
// and their backward passes w.r.t. bary_img.  Dense fp32 / int32 inputs (the host passes contiguous tensors, as
// the reference launchers do with .contiguous()).
//
// The normal-matrix value kernel is the scatter-heavy one (reference: nine atomics per foreground pixel).  Here a
// thread owns EIGHT consecutive pixels, keeps the nine sums of a run of equal triangle ids in registers and
// flushes them with nine fire-and-forget reductions when the id changes (runs are ~4 px on the 100k-triangle
// benchmark mesh: ~4x fewer reductions), the same walker idea as render / interpolate backward.
#include "common.cuh"

namespace drtk {
namespace {

// (:17-36) indices of the three columns in ascending column order (three compare-exchanges)
__device__ __forceinline__ void sorted_corner_order(const int32_t (&cols)[3], int (&order)[3]) {
  order[0] = 0; order[1] = 1; order[2] = 2;
  if (cols[order[1]] < cols[order[0]]) { const int t = order[0]; order[0] = order[1]; order[1] = t; }
  if (cols[order[2]] < cols[order[1]]) { const int t = order[1]; order[1] = order[2]; order[2] = t; }
  if (cols[order[1]] < cols[order[0]]) { const int t = order[0]; order[0] = order[1]; order[1] = t; }
}

// one thread per CSR row (= foreground pixel)
template <bool BACKWARD>
__global__ void __launch_bounds__(256) interp_matrix_kernel(int64_t nrows, const int32_t* __restrict__ vi,
                                                            const int32_t* __restrict__ index_img,
                                                            const float* __restrict__ bary_img,
                                                            const int64_t* __restrict__ row_pixels,
                                                            int64_t* __restrict__ col_indices, float* __restrict__ values,
                                                            const float* __restrict__ grad_values,
                                                            float* __restrict__ bary_grad, int64_t F, int64_t HW) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nrows) return;
  const int64_t flat = row_pixels[row];
  const int64_t n = flat / HW, hw = flat - n * HW;
  const int32_t tri = index_img[flat];
  const int32_t* face = vi + (n * F + tri) * 3;
  const int32_t cols[3] = {face[0], face[1], face[2]};
  int order[3];
  sorted_corner_order(cols, order);
  if (!BACKWARD) {
    const float* bp = bary_img + n * 3 * HW + hw;
    const float b[3] = {bp[0], bp[HW], bp[2 * HW]};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      // order[k] is data dependent: select instead of indexing the register arrays dynamically
      const int c = order[k];
      col_indices[row * 3 + k] = c == 0 ? cols[0] : (c == 1 ? cols[1] : cols[2]);
      values[row * 3 + k] = c == 0 ? b[0] : (c == 1 ? b[1] : b[2]);
    }
  } else {
    float* gp = bary_grad + n * 3 * HW + hw;  // zero-filled by the launcher (:774)
#pragma unroll
    for (int k = 0; k < 3; ++k) gp[(int64_t)order[k] * HW] = grad_values[row * 3 + k];
  }
}

constexpr int kNmPx = 8;

// PX consecutive pixels per thread (PX = 8 needs H*W % 8 == 0 so that a thread never straddles two images)
template <int PX>
__global__ void __launch_bounds__(256) normal_matrix_values_kernel(int64_t npix, const int32_t* __restrict__ pair,
                                                                   const int32_t* __restrict__ index_img,
                                                                   const float* __restrict__ bary_img,
                                                                   float* __restrict__ values, int64_t F, int64_t HW) {
  const int64_t p0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * PX;
  if (p0 >= npix) return;
  const int64_t n = p0 / HW, hw = p0 - n * HW;
  int ids[PX];
  float b0[PX], b1[PX], b2[PX];
  const float* bp = bary_img + n * 3 * HW + hw;
  if (PX == 8) {
    const int4 ia = ldg_stream_i4(index_img + p0), ib = ldg_stream_i4(index_img + p0 + 4);
    ids[0] = ia.x; ids[1 % PX] = ia.y; ids[2 % PX] = ia.z; ids[3 % PX] = ia.w;
    ids[4 % PX] = ib.x; ids[5 % PX] = ib.y; ids[6 % PX] = ib.z; ids[7 % PX] = ib.w;
    if ((ia.x & ia.y & ia.z & ia.w & ib.x & ib.y & ib.z & ib.w) == -1) return;
#pragma unroll
    for (int q = 0; q < PX / 4; ++q) {
      const float4 x = ldg_stream_f4(bp + 4 * q), y = ldg_stream_f4(bp + HW + 4 * q), z = ldg_stream_f4(bp + 2 * HW + 4 * q);
      b0[4 * q] = x.x; b0[(4 * q + 1) % PX] = x.y; b0[(4 * q + 2) % PX] = x.z; b0[(4 * q + 3) % PX] = x.w;
      b1[4 * q] = y.x; b1[(4 * q + 1) % PX] = y.y; b1[(4 * q + 2) % PX] = y.z; b1[(4 * q + 3) % PX] = y.w;
      b2[4 * q] = z.x; b2[(4 * q + 1) % PX] = z.y; b2[(4 * q + 2) % PX] = z.z; b2[(4 * q + 3) % PX] = z.w;
    }
  } else {
    ids[0] = index_img[p0];
    if (ids[0] == -1) return;
    b0[0] = bp[0]; b1[0] = bp[HW]; b2[0] = bp[2 * HW];
  }
  int cur = -1;
  float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // b0b0, b0b1, b0b2, b1b1, b1b2, b2b2 (the product is symmetric; (i,j) and (j,i) are separate slots)
  auto flush = [&]() {
    const int32_t* pr = pair + (n * F + cur) * 9;
    red_add(values + pr[0], acc[0]); red_add(values + pr[1], acc[1]); red_add(values + pr[2], acc[2]);
    red_add(values + pr[3], acc[1]); red_add(values + pr[4], acc[3]); red_add(values + pr[5], acc[4]);
    red_add(values + pr[6], acc[2]); red_add(values + pr[7], acc[4]); red_add(values + pr[8], acc[5]);
  };
#pragma unroll
  for (int j = 0; j < PX; ++j) {
    const int id = ids[j];
    if (id == -1) continue;
    if (id != cur) {
      if (cur != -1) flush();
      cur = id;
#pragma unroll
      for (int k = 0; k < 6; ++k) acc[k] = 0.f;
    }
    acc[0] = fmaf(b0[j], b0[j], acc[0]); acc[1] = fmaf(b0[j], b1[j], acc[1]); acc[2] = fmaf(b0[j], b2[j], acc[2]);
    acc[3] = fmaf(b1[j], b1[j], acc[3]); acc[4] = fmaf(b1[j], b2[j], acc[4]); acc[5] = fmaf(b2[j], b2[j], acc[5]);
  }
  if (cur != -1) flush();
}

// one thread per pixel; every pixel of bary_grad is written (zeros where empty, like the zero-filled tensor of :880)
__global__ void __launch_bounds__(256) normal_matrix_values_bwd_kernel(int64_t npix, const float* __restrict__ gv,
                                                                       const int32_t* __restrict__ pair,
                                                                       const int32_t* __restrict__ index_img,
                                                                       const float* __restrict__ bary_img,
                                                                       float* __restrict__ bary_grad, int64_t F, int64_t HW) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npix) return;
  const int64_t n = p / HW, hw = p - n * HW;
  float* gp = bary_grad + n * 3 * HW + hw;
  const int32_t tri = index_img[p];
  if (tri == -1) { gp[0] = 0.f; gp[HW] = 0.f; gp[2 * HW] = 0.f; return; }
  const int32_t* pr = pair + (n * F + tri) * 9;
  const float* bp = bary_img + n * 3 * HW + hw;
  const float b0 = bp[0], b1 = bp[HW], b2 = bp[2 * HW];
  const float g00 = gv[pr[0]], g01 = gv[pr[1]], g02 = gv[pr[2]], g10 = gv[pr[3]], g11 = gv[pr[4]], g12 = gv[pr[5]],
              g20 = gv[pr[6]], g21 = gv[pr[7]], g22 = gv[pr[8]];
  gp[0] = 2.f * g00 * b0 + (g01 + g10) * b1 + (g02 + g20) * b2;       // (:447-449)
  gp[HW] = (g10 + g01) * b0 + 2.f * g11 * b1 + (g12 + g21) * b2;
  gp[2 * HW] = (g20 + g02) * b0 + (g21 + g12) * b1 + 2.f * g22 * b2;
}

inline unsigned blocks_for(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

}  // namespace
}  // namespace drtk

using namespace drtk;

extern "C" int drtk_b200_interpolation_matrix(const int32_t* vi, const int32_t* index_img, const float* bary_img,
                                              const int64_t* row_pixels, int64_t N, int64_t F, int64_t H, int64_t W,
                                              int64_t nrows, int64_t* col_indices, float* values, void* stream_) {
  if (N < 0 || F < 0 || H < 0 || W < 0 || nrows < 0) return DRTK_B200_EINVAL;
  if (nrows == 0) return 0;
  if (!vi || !index_img || !bary_img || !row_pixels || !col_indices || !values) return DRTK_B200_EINVAL;
  if (nrows > (int64_t)0x7FFFFFFF * 256) return DRTK_B200_EUNSUPPORTED;
  interp_matrix_kernel<false><<<blocks_for(nrows, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      nrows, vi, index_img, bary_img, row_pixels, col_indices, values, nullptr, nullptr, F, H * W);
  DRTK_CHECK_LAUNCH();
  return 0;
}

extern "C" int drtk_b200_interpolation_matrix_backward(const float* grad_values, const int32_t* vi,
                                                       const int32_t* index_img, const int64_t* row_pixels, int64_t N,
                                                       int64_t F, int64_t H, int64_t W, int64_t nrows,
                                                       float* bary_grad, void* stream_) {
  if (N < 0 || F < 0 || H < 0 || W < 0 || nrows < 0) return DRTK_B200_EINVAL;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (N * H * W > 0) {
    if (!bary_grad) return DRTK_B200_EINVAL;
    DRTK_CUDA(cudaMemsetAsync(bary_grad, 0, sizeof(float) * (size_t)(N * 3 * H * W), stream));
  }
  if (nrows == 0) return 0;
  if (!grad_values || !vi || !index_img || !row_pixels) return DRTK_B200_EINVAL;
  interp_matrix_kernel<true><<<blocks_for(nrows, 256), 256, 0, stream>>>(nrows, vi, index_img, nullptr, row_pixels, nullptr,
                                                                         nullptr, grad_values, bary_grad, F, H * W);
  DRTK_CHECK_LAUNCH();
  return 0;
}

extern "C" int drtk_b200_interpolation_normal_matrix_values(const int32_t* pair_indices, const int32_t* index_img,
                                                            const float* bary_img, int64_t N, int64_t F, int64_t H,
                                                            int64_t W, int64_t nnz, float* values, void* stream_) {
  if (N < 0 || F < 0 || H < 0 || W < 0 || nnz < 0) return DRTK_B200_EINVAL;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (nnz > 0) {
    if (!values) return DRTK_B200_EINVAL;
    DRTK_CUDA(cudaMemsetAsync(values, 0, sizeof(float) * (size_t)nnz, stream));  // (:820)
  }
  const int64_t npix = N * H * W;
  if (npix == 0 || F == 0 || nnz == 0) return 0;
  if (!pair_indices || !index_img || !bary_img) return DRTK_B200_EINVAL;
  const auto al16 = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
  if ((H * W) % kNmPx == 0 && al16(index_img) && al16(bary_img))
    normal_matrix_values_kernel<kNmPx><<<blocks_for(npix / kNmPx, 256), 256, 0, stream>>>(npix, pair_indices, index_img, bary_img,
                                                                                         values, F, H * W);
  else
    normal_matrix_values_kernel<1><<<blocks_for(npix, 256), 256, 0, stream>>>(npix, pair_indices, index_img, bary_img, values, F,
                                                                              H * W);
  DRTK_CHECK_LAUNCH();
  return 0;
}

extern "C" int drtk_b200_interpolation_normal_matrix_values_backward(const float* grad_values, const int32_t* pair_indices,
                                                                     const int32_t* index_img, const float* bary_img,
                                                                     int64_t N, int64_t F, int64_t H, int64_t W,
                                                                     float* bary_grad, void* stream_) {
  if (N < 0 || F < 0 || H < 0 || W < 0) return DRTK_B200_EINVAL;
  const int64_t npix = N * H * W;
  if (npix == 0) return 0;
  if (!bary_grad || !index_img) return DRTK_B200_EINVAL;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (F == 0) {  // nothing can be covered
    DRTK_CUDA(cudaMemsetAsync(bary_grad, 0, sizeof(float) * (size_t)(npix * 3), stream));
    return 0;
  }
  if (!grad_values || !pair_indices || !bary_img) return DRTK_B200_EINVAL;
  normal_matrix_values_bwd_kernel<<<blocks_for(npix, 256), 256, 0, stream>>>(npix, grad_values, pair_indices, index_img, bary_img,
                                                                             bary_grad, F, H * W);
  DRTK_CHECK_LAUNCH();
  return 0;
}
