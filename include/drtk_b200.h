/*
 * drtk_b200.h -- C ABI of libdrtk_b200.so: the B200 (sm_100a) rasterisation hot path.
 *
 * This is the drop-in boundary.  Every entry point replaces one host launcher of the
 * reference (facebookresearch/DRTK) that sits behind a TORCH_LIBRARY op; the reference
 * interface each one stands in for is cited as path:line relative to the reference root.
 * The signatures carry only plain pointers, sizes, element strides and a CUDA stream --
 * no torch types -- so the library can be bound from ctypes (drtk_b200/_lib.py), from a
 * TORCH_LIBRARY shim or from any other host.
 *
 * Conventions
 *   - all data pointers are DEVICE pointers on the current CUDA device; fp32 / int32 only
 *   - `*_strides` are HOST arrays of element strides (as torch.Tensor.stride()), so
 *     expanded (stride 0) and non-contiguous inputs are accepted exactly like the
 *     reference kernels accept them (TensorInfo strides, e.g. src/render/render_kernel.cu:36-55)
 *   - outputs are dense, freshly allocated by the caller, fully written by the callee
 *     (no pre-zeroing needed, also not for the gradient accumulators)
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued asynchronously,
 *     nothing synchronises the device
 *   - return value: 0 on success, otherwise a cudaError_t value (>0) or a DRTK_B200_E*
 *     code (<0); drtk_b200_error_string() turns either into text.  No CPU fallback exists.
 */
#ifndef DRTK_B200_H_
#define DRTK_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRTK_B200_ABI_VERSION 2

#define DRTK_B200_EINVAL (-1)     /* bad argument (null pointer, non-positive size, ...) */
#define DRTK_B200_EWORKSPACE (-2) /* workspace too small */
#define DRTK_B200_EUNSUPPORTED (-3)

int drtk_b200_abi_version(void);
const char* drtk_b200_error_string(int code);

/* ---------------------------------------------------------------------------------------
 * rasterize  -- replaces rasterize_cuda (src/rasterize/rasterize_kernel.cu:417-563), i.e.
 * the op `rasterize_ext::rasterize(Tensor v, Tensor vi, int height, int width,
 * bool wireframe) -> Tensor[]` (src/rasterize/rasterize_module.cpp:77-79).
 *
 *   v          [N,V,3] f32, strides v_strides[3]
 *   vi         [N,F,3] i32, strides vi_strides[3] (batch stride 0 for a shared topology)
 *   depth_img  [N,H,W] f32 out (0 where empty);  index_img [N,H,W] i32 out (-1 where empty)
 *   workspace  scratch of at least drtk_b200_rasterize_workspace_bytes(...) bytes
 *   wireframe  0 = filled triangles (rasterize_kernel, :42-168)
 *              1 = wireframe (rasterize_lines_kernel, :261-400): a pixel carries a triangle id when a visible
 *                  edge crosses the diamond |dx|+|dy| = 0.5 around its centre; edge visibility = bits 28-30 of
 *                  vi[..., 0]; interiors still write depth with index -1 (occluders)
 *   algo       0 = tile-binned, shared-memory z-buffer (default)
 *              1 = triangle-parallel 64-bit global atomicMin (validation path; same bits)
 *              (ignored in wireframe mode: one warp per triangle, packed global z-buffer)
 * Bit-exact contract: depth_img / index_img equal the reference CUDA kernels' output, in both modes.
 * ------------------------------------------------------------------------------------- */
size_t drtk_b200_rasterize_workspace_bytes(int64_t N, int64_t F, int64_t H, int64_t W, int algo);

int drtk_b200_rasterize(const float* v, const int64_t* v_strides, const int32_t* vi,
                        const int64_t* vi_strides, int64_t N, int64_t V, int64_t F, int64_t H,
                        int64_t W, int wireframe, int algo, float* depth_img, int32_t* index_img,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * render forward -- replaces render_cuda (src/render/render_kernel.cu:283-380), op
 * `render_ext::render(Tensor v, Tensor vi, Tensor index_img) -> Tensor[]`
 * (src/render/render_module.cpp:89-91).
 *   index_img [N,H,W] i32 (strides index_strides[3])
 *   depth_img [N,H,W] f32 out;  bary_img [N,3,H,W] f32 out (planar)
 * ------------------------------------------------------------------------------------- */
int drtk_b200_render_forward(const float* v, const int64_t* v_strides, const int32_t* vi,
                             const int64_t* vi_strides, const int32_t* index_img,
                             const int64_t* index_strides, int64_t N, int64_t V, int64_t F,
                             int64_t H, int64_t W, float* depth_img, float* bary_img,
                             void* stream);

/* render backward -- replaces render_cuda_backward (src/render/render_kernel.cu:382-436).
 *   grad_depth [N,H,W] (may be NULL = zeros), grad_bary [N,3,H,W] (may be NULL = zeros)
 *   grad_v     [N,V,3] f32 out, dense; every element written by the callee
 *   workspace  scratch of at least drtk_b200_render_backward_workspace_bytes(N,V,F)
 *              bytes (per-triangle setup table + 16-B padded accumulators of the fast path)       */
size_t drtk_b200_render_backward_workspace_bytes(int64_t N, int64_t V, int64_t F);

int drtk_b200_render_backward(const float* v, const int64_t* v_strides, const int32_t* vi,
                              const int64_t* vi_strides, const int32_t* index_img,
                              const int64_t* index_strides, const float* grad_depth,
                              const int64_t* grad_depth_strides, const float* grad_bary,
                              const int64_t* grad_bary_strides, int64_t N, int64_t V, int64_t F,
                              int64_t H, int64_t W, float* grad_v, void* workspace,
                              size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * interpolate forward -- replaces interpolate_cuda
 * (src/interpolate/interpolate_kernel.cu:454-570), op `interpolate_ext::interpolate(Tensor
 * vert_attributes, Tensor vi, Tensor index_img, Tensor bary_img) -> Tensor`
 * (src/interpolate/interpolate_module.cpp:632-634).
 *   vert_attributes [N,V,C] f32;  bary_img [N,3,H,W] f32;  out [N,C,H,W] f32 (planar)
 * ------------------------------------------------------------------------------------- */
int drtk_b200_interpolate_forward(const float* vert_attributes, const int64_t* attr_strides,
                                  const int32_t* vi, const int64_t* vi_strides,
                                  const int32_t* index_img, const int64_t* index_strides,
                                  const float* bary_img, const int64_t* bary_strides, int64_t N,
                                  int64_t V, int64_t F, int64_t C, int64_t H, int64_t W,
                                  float* out, void* stream);

/* interpolate backward -- replaces interpolate_cuda_backward
 * (src/interpolate/interpolate_kernel.cu:642-697).
 *   grad_out [N,C,H,W] (strides grad_out_strides[4])
 *   vert_attributes_grad [N,V,C] out or NULL (zero-filled by the callee, then accumulated)
 *   bary_img_grad        [N,3,H,W] out or NULL (every pixel written)
 *   workspace            device scratch of at least drtk_b200_interpolate_backward_workspace_bytes(N, F,
 *                        vi_strides[0]) bytes (the packed per-triangle vertex-id table of the tiled fast path;
 *                        the library itself never allocates device memory)                              */
size_t drtk_b200_interpolate_backward_workspace_bytes(int64_t N, int64_t F, int64_t vi_batch_stride);

int drtk_b200_interpolate_backward(const float* grad_out, const int64_t* grad_out_strides,
                                   const float* vert_attributes, const int64_t* attr_strides,
                                   const int32_t* vi, const int64_t* vi_strides,
                                   const int32_t* index_img, const int64_t* index_strides,
                                   const float* bary_img, const int64_t* bary_strides, int64_t N,
                                   int64_t V, int64_t F, int64_t C, int64_t H, int64_t W,
                                   float* vert_attributes_grad, float* bary_img_grad,
                                   void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * Sparse interpolation matrices of a fixed rasterisation (dense, contiguous inputs, like the reference
 * launchers which call .contiguous()):
 *   interpolation_matrix -- replaces the kernel launch of interpolation_matrix_cuda
 *     (src/interpolate/interpolate_kernel.cu:699-762; op `interpolate_ext::interpolation_matrix`,
 *     src/interpolate/interpolate_module.cpp:635-636).  row_pixels [nrows] i64 = flattened indices of the
 *     foreground pixels (the host computes them, as the reference does with at::nonzero); outputs
 *     col_indices [3*nrows] i64 (ascending per row) and values [3*nrows] f32.
 *   ..._backward -- replaces interpolation_matrix_cuda_backward (:764-803): bary_grad [N,3,H,W], zero-filled by
 *     the callee, receives grad_values at the foreground pixels.
 *   interpolation_normal_matrix_values -- replaces interpolation_normal_matrix_values_cuda (:805-860; op
 *     `interpolate_ext::interpolation_normal_matrix_values`): values [nnz] (zero-filled by the callee) +=
 *     b_i*b_j at slot pair_indices[n, tri, i*3+j] for every foreground pixel; pair_indices [N,F,9] i32.
 *   ..._backward -- replaces interpolation_normal_matrix_values_cuda_backward (:862-900): bary_grad [N,3,H,W],
 *     every pixel written.
 * ------------------------------------------------------------------------------------- */
int drtk_b200_interpolation_matrix(const int32_t* vi, const int32_t* index_img, const float* bary_img,
                                   const int64_t* row_pixels, int64_t N, int64_t F, int64_t H, int64_t W,
                                   int64_t nrows, int64_t* col_indices, float* values, void* stream);

int drtk_b200_interpolation_matrix_backward(const float* grad_values, const int32_t* vi,
                                            const int32_t* index_img, const int64_t* row_pixels, int64_t N,
                                            int64_t F, int64_t H, int64_t W, int64_t nrows, float* bary_grad,
                                            void* stream);

int drtk_b200_interpolation_normal_matrix_values(const int32_t* pair_indices, const int32_t* index_img,
                                                 const float* bary_img, int64_t N, int64_t F, int64_t H,
                                                 int64_t W, int64_t nnz, float* values, void* stream);

int drtk_b200_interpolation_normal_matrix_values_backward(const float* grad_values, const int32_t* pair_indices,
                                                          const int32_t* index_img, const float* bary_img,
                                                          int64_t N, int64_t F, int64_t H, int64_t W,
                                                          float* bary_grad, void* stream);

/* ---------------------------------------------------------------------------------------
 * edge_grad backward -- replaces edge_grad_estimator_cuda_backward
 * (src/edge_grad/edge_grad_kernel.cu:475-506), the backward of op
 * `edge_grad_ext::edge_grad_estimator(...)` (src/edge_grad/edge_grad_module.cpp:205-208;
 * its forward is the identity on `img`, :118-137, and needs no kernel).
 *   v_pix [N,V,3]; img [N,C,H,W]; grad_output [N,C,H,W]; grad_v_pix_img [N,3,H,W] out
 *   (every pixel written: gather formulation, no zero-fill, no atomics, deterministic)
 * ------------------------------------------------------------------------------------- */
int drtk_b200_edge_grad_backward(const float* v_pix, const int64_t* v_strides, const float* img,
                                 const int64_t* img_strides, const int32_t* index_img,
                                 const int64_t* index_strides, const int32_t* vi,
                                 const int64_t* vi_strides, const float* grad_output,
                                 const int64_t* grad_output_strides, int64_t N, int64_t V,
                                 int64_t F, int64_t C, int64_t H, int64_t W, float max_dp_dr,
                                 float* grad_v_pix_img, void* stream);

/* edge_grad backward fused with the backward of the conduit `interpolate(v_pix, vi, index_img,
 * bary_img.detach())` through which the reference routes dL/d(v_pix_img) to the vertices
 * (drtk/edge_grad_estimator.py:165-180, src/edge_grad/edge_grad_module.cpp:139-169 followed by
 * interpolate_cuda_backward with C = 3, src/interpolate/interpolate_kernel.cu:642-697).  Used when no
 * `v_pix_img_hook` needs to see the [N,3,H,W] image.
 *   bary_img   [N,3,H,W] (strides bary_strides[4])
 *   grad_v_pix [N,V,3] out, dense; zero-filled by the callee, then accumulated
 *   workspace  scratch of at least drtk_b200_edge_grad_backward_fused_workspace_bytes(N,F)
 *              bytes (per-triangle screen-space vertex table of the fast path)                        */
size_t drtk_b200_edge_grad_backward_fused_workspace_bytes(int64_t N, int64_t F);

int drtk_b200_edge_grad_backward_fused(const float* v_pix, const int64_t* v_strides, const float* img,
                                       const int64_t* img_strides, const int32_t* index_img,
                                       const int64_t* index_strides, const int32_t* vi,
                                       const int64_t* vi_strides, const float* grad_output,
                                       const int64_t* grad_output_strides, const float* bary_img,
                                       const int64_t* bary_strides, int64_t N, int64_t V, int64_t F,
                                       int64_t C, int64_t H, int64_t W, float max_dp_dr,
                                       float* grad_v_pix, void* workspace, size_t workspace_bytes,
                                       void* stream);

/* ---------------------------------------------------------------------------------------
 * transform -- the world -> pixel projection in front of the rasteriser.  The reference
 * has no native op here: `drtk.transform` (drtk/transform.py:14-119) runs `project_points`
 * (drtk/utils/projection.py:486-646) as ~20 stock torch kernels forward and as many again backward; these two
 * entry points do each direction in ONE kernel.
 *   v     [N,V,3] f32, strides v_strides[3]
 *   cam   [N,28] f32 dense, one block per batch item:
 *         [0:3] campos  [3:12] camrot row-major  [12:16] focal row-major  [16:18] princpt
 *         [18:26] distortion coefficients (zero-padded)  [26] fov (max normalised radius)  [27] unused
 *   mode  DRTK_B200_DIST_*; `modes` (device, [N] i32) overrides it per batch item when not NULL
 *         PINHOLE            project_pinhole             (projection.py:33-53)
 *         RADIAL_TANGENTIAL  project_pinhole_distort_rt  (:56-136)  D = k1 k2 p1 p2 [k3 [k4 k5 k6]]
 *         FISHEYE            project_fisheye_distort     (:139-186) D = k0..k3
 *         FISHEYE62          project_fisheye_distort_62  (:189-276, without the LUT) D = k0..k5 p0 p1
 *   cull_outside_fov  fisheye62 only: vertices with |p| > fov get z = -1 (projection.py:624-644)
 *   v_pix [N,V,3] out = (x_pix, y_pix, z_cam);  v_cam [N,V,3] out or NULL
 * backward: grad_v_pix / grad_v_cam [N,V,3] (either may be NULL = zeros, any strides);
 *   grad_v [N,V,3] out or NULL (every element written); grad_cam [N,28] out or NULL (zero-filled by the
 *   callee, then accumulated; slots 26, 27 stay 0: fov is not differentiated)
 * ------------------------------------------------------------------------------------- */
#define DRTK_B200_DIST_PINHOLE 0
#define DRTK_B200_DIST_RADIAL_TANGENTIAL 1
#define DRTK_B200_DIST_FISHEYE 2
#define DRTK_B200_DIST_FISHEYE62 3

int drtk_b200_transform_forward(const float* v, const int64_t* v_strides, const float* cam,
                                const int32_t* modes, int mode, int cull_outside_fov, int64_t N, int64_t V,
                                float* v_pix, float* v_cam, void* stream);

int drtk_b200_transform_backward(const float* v, const int64_t* v_strides, const float* cam,
                                 const int32_t* modes, int mode, int cull_outside_fov,
                                 const float* grad_v_pix, const int64_t* grad_v_pix_strides,
                                 const float* grad_v_cam, const int64_t* grad_v_cam_strides, int64_t N,
                                 int64_t V, float* grad_v, float* grad_cam, void* stream);

/* ---------------------------------------------------------------------------------------
 * Texture samplers either side of `interpolate`.  Enumerations as in the reference ops:
 *   padding_mode 0 zeros, 1 border, 2 reflection;  interpolation_mode 0 bilinear, 2 bicubic
 *   (drtk/mipmap_grid_sample.py:100-113, drtk/grid_scatter.py:79-92); coordinates in [-1, 1] as grid_sample.
 *
 * mipmap_grid_sample forward -- replaces mipmap_aniso_grid_sampler_2d_cuda
 * (src/mipmap_grid_sampler/mipmap_grid_sampler_kernel.cu:900-1097), op
 * `mipmap_grid_sampler_ext::mipmap_grid_sampler_2d(Tensor[] input, Tensor grid, Tensor vt_dxdy_img, int max_aniso,
 * int padding_mode, int interpolation_mode, bool align_corners, bool force_max_ansio, bool clip_grad) -> Tensor`
 * (src/mipmap_grid_sampler/mipmap_grid_sampler_module.cpp:252-256).
 *   levels        HOST array of num_levels (1..11) device pointers, level l = [N,C,H_l,W_l] f32
 *   level_hw      HOST [num_levels][2] = (H_l, W_l);  level_strides HOST [num_levels][4] element strides
 *   grid          [N,H,W,2] f32 (strides grid_strides[4]);  vt_dxdy_img [N,H,W,2,2] f32 (strides vt_strides[5])
 *   out           [N,C,H,W] f32 dense, every element written
 * Like the reference kernel (:423) the FORWARD ignores align_corners (always false); the backward honours it.
 * backward -- replaces mipmap_aniso_grid_sampler_2d_cuda_backward (:1099-1249):
 *   grad_levels   HOST array of num_levels device pointers to dense [N,C,H_l,W_l] accumulators (zero-filled by
 *                 the callee), or NULL when no texture needs a gradient
 *   grad_grid     [N,H,W,2] dense, every element written, or NULL.   vt_dxdy_img receives no gradient.
 *
 * grid_scatter forward -- replaces grid_scatter_2d_cuda (src/grid_scatter/grid_scatter_kernel.cu:624-729), op
 * `grid_scatter_ext::grid_scatter_2d(Tensor input, Tensor grid, int output_height, int output_width,
 * int padding_mode, int interpolation_mode, bool align_corners) -> Tensor` (src/grid_scatter/grid_scatter_module.cpp:137-140):
 *   input [N,C,H,W], grid [N,H,W,2] -> out [N,C,out_H,out_W] dense (zero-filled by the callee, then accumulated)
 * backward -- replaces grid_scatter_2d_cuda_backward (:731-788): grad_input [N,C,H,W] dense or NULL,
 *   grad_grid [N,H,W,2] dense or NULL; every element written.
 * ------------------------------------------------------------------------------------- */
int drtk_b200_mipmap_grid_sample_forward(const float* const* levels, const int64_t* level_hw,
                                         const int64_t* level_strides, int num_levels, const float* grid,
                                         const int64_t* grid_strides, const float* vt_dxdy_img,
                                         const int64_t* vt_strides, int64_t N, int64_t C, int64_t H, int64_t W,
                                         int max_aniso, int padding_mode, int interpolation_mode, int align_corners,
                                         int force_max_aniso, int clip_grad, float* out, void* stream);

int drtk_b200_mipmap_grid_sample_backward(const float* grad_out, const int64_t* grad_out_strides,
                                          const float* const* levels, const int64_t* level_hw,
                                          const int64_t* level_strides, int num_levels, const float* grid,
                                          const int64_t* grid_strides, const float* vt_dxdy_img,
                                          const int64_t* vt_strides, int64_t N, int64_t C, int64_t H, int64_t W,
                                          int max_aniso, int padding_mode, int interpolation_mode, int align_corners,
                                          int force_max_aniso, int clip_grad, float* const* grad_levels,
                                          float* grad_grid, void* stream);

int drtk_b200_grid_scatter_forward(const float* input, const int64_t* input_strides, const float* grid,
                                   const int64_t* grid_strides, int64_t N, int64_t C, int64_t H, int64_t W,
                                   int64_t out_H, int64_t out_W, int padding_mode, int interpolation_mode,
                                   int align_corners, float* out, void* stream);

int drtk_b200_grid_scatter_backward(const float* grad_out, const int64_t* grad_out_strides, const float* input,
                                    const int64_t* input_strides, const float* grid, const int64_t* grid_strides,
                                    int64_t N, int64_t C, int64_t H, int64_t W, int64_t out_H, int64_t out_W,
                                    int padding_mode, int interpolation_mode, int align_corners, float* grad_input,
                                    float* grad_grid, void* stream);

/* ---------------------------------------------------------------------------------------
 * Multi-GPU support: the batch shards over ranks with no data-path collective; the only
 * exchange is the gradient of parameters SHARED by all batch items.  The reference has no counterpart (its
 * users get the per-rank batch sum from autograd's expand-backward and the cross-rank sum from DDP / NCCL).
 *
 * batch_sum: out[m] = sum_n x[n * batch_stride + m], m < M  -- the local batch reduction, written straight into the
 *   communication bucket.
 * batch_sum_allreduce: the same followed, IN THE SAME KERNEL, by the sum over all ranks through NVSwitch multicast
 *   memory (multimem.red.add.f32 into every rank's copy of the bucket, one flag barrier over peer memory); no NCCL
 *   call.  The bucket is double buffered: acc_multicast = multicast address of the half that receives this call's sums
 *   (every rank's copy of it must be zero: it was zero_local of that rank's previous call, or freshly zeroed memory),
 *   zero_local = THIS rank's copy of the other half, zero-filled by this call for the next one (M floats each).
 *   peer_flags[r]: rank r's flag array (world * drtk_b200_batch_sum_allreduce_grid() zero-initialised uint32) as mapped
 *   into this process; epoch: number of earlier calls on these flags.  Every rank must make the same sequence of
 *   calls.  On return (stream order) this rank's copy of the accumulated half holds the sum over ranks; consume it
 *   before the call after next (which zero-fills it).  timeout_flag (device int, may be NULL) is set to 1 when a peer
 *   did not arrive within ~2 s (the wait then gives up instead of hanging the GPU; the bucket is invalid).
 *   max_ctas: 0 = one full co-resident wave (drtk_b200_batch_sum_allreduce_grid() CTAs; for an exchange on the critical
 *   path); > 0 caps the grid for an exchange that overlaps other kernels (its CTAs wait at the barrier for the slowest
 *   rank).  Every rank must pass the same value.
 * ------------------------------------------------------------------------------------- */
int drtk_b200_batch_sum(const float* x, int64_t N, int64_t M, int64_t batch_stride, float* out, void* stream);

int drtk_b200_batch_sum_allreduce_grid(void);

int drtk_b200_batch_sum_allreduce(const float* x, int64_t N, int64_t M, int64_t batch_stride, float* zero_local,
                                  float* acc_multicast, void* const* peer_flags, int rank, int world,
                                  uint32_t epoch, int* timeout_flag, int max_ctas, void* stream);

/* ---------------------------------------------------------------------------------------
 * float64 dispatch of the six hot-path launchers (the reference instantiates every kernel for float and double,
 * src/include/kernel_utils.h:47-57).  Same arguments and contracts as the float entry points above, with double
 * data; no workspace except the packed 64-bit z-buffer of the rasteriser; `rasterize_f64` still writes a float
 * depth_img (src/rasterize/rasterize_kernel.cu:481); wireframe mode included (:492-535).  Plain kernels (thread per
 * pixel, atomicAdd(double)): fp64 is for gradient checks, not throughput.
 * ------------------------------------------------------------------------------------- */
size_t drtk_b200_rasterize_f64_workspace_bytes(int64_t N, int64_t H, int64_t W);

int drtk_b200_rasterize_f64(const double* v, const int64_t* v_strides, const int32_t* vi, const int64_t* vi_strides,
                            int64_t N, int64_t V, int64_t F, int64_t H, int64_t W, int wireframe, float* depth_img,
                            int32_t* index_img, void* workspace, size_t workspace_bytes, void* stream);

int drtk_b200_render_forward_f64(const double* v, const int64_t* v_strides, const int32_t* vi,
                                 const int64_t* vi_strides, const int32_t* index_img, const int64_t* index_strides,
                                 int64_t N, int64_t V, int64_t F, int64_t H, int64_t W, double* depth_img,
                                 double* bary_img, void* stream);

int drtk_b200_render_backward_f64(const double* v, const int64_t* v_strides, const int32_t* vi,
                                  const int64_t* vi_strides, const int32_t* index_img, const int64_t* index_strides,
                                  const double* grad_depth, const int64_t* grad_depth_strides, const double* grad_bary,
                                  const int64_t* grad_bary_strides, int64_t N, int64_t V, int64_t F, int64_t H,
                                  int64_t W, double* grad_v, void* stream);

int drtk_b200_interpolate_forward_f64(const double* vert_attributes, const int64_t* attr_strides, const int32_t* vi,
                                      const int64_t* vi_strides, const int32_t* index_img, const int64_t* index_strides,
                                      const double* bary_img, const int64_t* bary_strides, int64_t N, int64_t V,
                                      int64_t F, int64_t C, int64_t H, int64_t W, double* out, void* stream);

int drtk_b200_interpolate_backward_f64(const double* grad_out, const int64_t* grad_out_strides,
                                       const double* vert_attributes, const int64_t* attr_strides, const int32_t* vi,
                                       const int64_t* vi_strides, const int32_t* index_img, const int64_t* index_strides,
                                       const double* bary_img, const int64_t* bary_strides, int64_t N, int64_t V,
                                       int64_t F, int64_t C, int64_t H, int64_t W, double* vert_attributes_grad,
                                       double* bary_img_grad, void* stream);

int drtk_b200_edge_grad_backward_f64(const double* v_pix, const int64_t* v_strides, const double* img,
                                     const int64_t* img_strides, const int32_t* index_img, const int64_t* index_strides,
                                     const int32_t* vi, const int64_t* vi_strides, const double* grad_output,
                                     const int64_t* grad_output_strides, int64_t N, int64_t V, int64_t F, int64_t C,
                                     int64_t H, int64_t W, double max_dp_dr, double* grad_v_pix_img, void* stream);

/* ---------------------------------------------------------------------------------------
 * screen_space_uv_derivative -- the producer of `vt_dxdy_img` for mipmap_grid_sample, one kernel instead of the
 * reference's composition of face_dpdt + 2 x interpolate + project_points_grad + 2x2 inverse
 * (drtk/screen_space_uv_derivative.py:16-80; pinhole camera, as in the reference).
 *   v [N,V,3], vt [N,T,2] f32; vi, vti [F,3] i32 (shared topology; element strides *_strides[2])
 *   index_img [N,H,W] i32; bary_img [N,3,H,W] f32; mask [N,H,W] bool (1 byte per pixel)
 *   cam [N,16] f32 dense: campos 3, camrot 9 (row major), focal 4 (row major)
 *   out [N,H,W,2,2] f32 dense = [[du/dx, dv/dx], [du/dy, dv/dy]]; zero where mask is false or the pixel is empty
 * Forward only (no gradient); the Python host falls back to the differentiable composition when one is needed.
 * ------------------------------------------------------------------------------------- */
int drtk_b200_screen_space_uv_derivative(const float* v, const int64_t* v_strides, const float* vt,
                                         const int64_t* vt_strides, const int32_t* vi, const int64_t* vi_strides,
                                         const int32_t* vti, const int64_t* vti_strides, const int32_t* index_img,
                                         const int64_t* index_strides, const float* bary_img,
                                         const int64_t* bary_strides, const uint8_t* mask, const int64_t* mask_strides,
                                         const float* cam, int64_t N, int64_t H, int64_t W, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DRTK_B200_H_ */
