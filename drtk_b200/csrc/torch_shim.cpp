// torch_shim.cpp -- the PyTorch-dispatcher face of libdrtk_b200.so.
//
// The reference exposes its kernels as dispatcher ops registered from shared libraries:
//   TORCH_LIBRARY(rasterize_ext)   rasterize(Tensor v, Tensor vi, int height, int width, bool wireframe) -> Tensor[]
//                                  (src/rasterize/rasterize_module.cpp:77-95)
//   TORCH_LIBRARY(render_ext)      render(Tensor v, Tensor vi, Tensor index_img) -> Tensor[]
//                                  (src/render/render_module.cpp:89-107)
//   TORCH_LIBRARY(interpolate_ext) interpolate(Tensor vert_attributes, Tensor vi, Tensor index_img, Tensor bary_img) -> Tensor
//                                  (src/interpolate/interpolate_module.cpp:632-668)
//   TORCH_LIBRARY(edge_grad_ext)   edge_grad_estimator(Tensor v_pix, Tensor v_pix_img, Tensor vi, Tensor img,
//                                  Tensor index_img, float max_dp_dr=1e4) -> Tensor
//                                  (src/edge_grad/edge_grad_module.cpp:205-224)
// each with Autograd (custom Function), Autocast (cast reduced-precision floats to fp32) and CUDA implementations.
// This file registers the SAME schemas under drtk_b200_<name>_ext (one process cannot hold two definitions of
// rasterize_ext) with the same three dispatch keys; the CUDA implementations call the C ABI of
// include/drtk_b200.h.  The reference's Python layer binds to it by changing one string per op.
// There is no CPU implementation: the CPU key maps to the same launchers, whose first TORCH_CHECK rejects a CPU tensor
// with the reference's own wording ("...expected all inputs to be on same cuda device").
//
// Extra op (no counterpart in the reference): drtk_b200_edge_grad_ext::edge_grad_estimator_fused, the estimator
// with its C = 3 conduit folded in (no [N,3,H,W] gradient image; used when no v_pix_img hook is registered).
//
// Built by __graft_entry__.build() / drtk_b200.build_torch_ops() with torch.utils.cpp_extension (host C++ only,
// links libdrtk_b200.so, $ORIGIN rpath).
#include <ATen/autocast_mode.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/autograd.h>
#include <torch/library.h>

#include <tuple>
#include <vector>

#include "../../include/drtk_b200.h"

namespace {

using at::Tensor;
using torch::autograd::AutogradContext;
using torch::autograd::tensor_list;

// ---- plumbing ----------------------------------------------------------------------------------------------
struct Strides {
  int64_t s[4];
  explicit Strides(const Tensor& t) {
    for (int i = 0; i < 4; ++i) s[i] = i < t.dim() ? t.stride(i) : 0;
  }
  operator const int64_t*() const { return s; }
};

inline void* stream_of(const Tensor& t) { return at::cuda::getCurrentCUDAStream(t.get_device()).stream(); }

inline void ok(int rc, const char* who) { TORCH_CHECK(rc == 0, who, ": ", drtk_b200_error_string(rc)); }

inline Tensor scratch(size_t bytes, const Tensor& like) {
  return at::empty({(int64_t)(bytes > 16 ? bytes : 16)}, like.options().dtype(at::kByte));
}

template <class T>
inline const T* cptr(const Tensor& t) { return t.defined() ? t.data_ptr<T>() : nullptr; }

inline bool is_f64(const Tensor& t) { return t.scalar_type() == at::kDouble; }

// fp32 and fp64 are served natively; anything else must have been cast by the Autocast kernel
inline void need_real(const Tensor& t, const char* who, const char* name) {
  TORCH_CHECK(t.is_floating_point(), who, "(): expected ", name, " to have floating point type, but ", name, " has ", t.dtype());
  TORCH_CHECK(t.scalar_type() == at::kFloat || t.scalar_type() == at::kDouble, who,
              "(): drtk_b200 computes in float32 only (float64: the plain double kernels), but ", name, " has ", t.dtype(),
              "; cast it to float32 or run under torch.autocast");
}

// ---- rasterize (checks: src/rasterize/rasterize_kernel.cu:423-468) ----------------------------------------------
std::vector<Tensor> rasterize_cuda(const Tensor& v, const Tensor& vi, int64_t height, int64_t width, bool wireframe) {
  TORCH_CHECK(v.defined() && vi.defined(), "rasterize(): expected all inputs to be defined");
  TORCH_CHECK(v.device() == vi.device() && v.is_cuda(), "rasterize(): expected all inputs to be on same cuda device");
  need_real(v, "rasterize", "v");
  TORCH_CHECK(vi.scalar_type() == at::kInt, "rasterize(): expected vi to have int32 type, but vi has ", vi.dtype());
  TORCH_CHECK(v.layout() == at::kStrided && vi.layout() == at::kStrided, "rasterize(): expected all inputs to have torch.strided layout");
  TORCH_CHECK(v.dim() == 3 && vi.dim() == 3, "rasterize(): expected v.ndim == 3, vi.ndim == 3, but got v with sizes ", v.sizes(),
              " and vi with sizes ", vi.sizes());
  TORCH_CHECK(v.size(2) == 3 && vi.size(2) == 3,
              "rasterize(): expected third dim of v to be of size 3, and last dim of vi to be of size 3, but got ", v.size(2),
              " in the third dim of v, and ", vi.size(2), " in the last dim of vi");
  TORCH_CHECK(vi.size(0) == v.size(0), "rasterize(): expected first dim of vi to match first dim of v, but got ", v.size(0),
              " in first dim of v, and ", vi.size(0), " in the first dim of vi");
  TORCH_CHECK(v.size(1) < 0x10000000LL, "rasterize(): expected second dim of v to be less or eual to 268435456, but got ", v.size(1));
  TORCH_CHECK(height > 0 && width > 0, "rasterize(): both height and width have to be greater than zero, but got height: ", height,
              ", and width: ", width);
  const c10::cuda::CUDAGuard guard(v.device());
  const int64_t N = v.size(0), V = v.size(1), F = vi.size(1);
  Tensor depth = at::empty({N, height, width}, v.options().dtype(at::kFloat));  // float even for double v (:481)
  Tensor index = at::empty({N, height, width}, v.options().dtype(at::kInt));
  if (is_f64(v)) {
    Tensor ws = scratch(drtk_b200_rasterize_f64_workspace_bytes(N, height, width), v);
    ok(drtk_b200_rasterize_f64(v.data_ptr<double>(), Strides(v), cptr<int32_t>(vi), Strides(vi), N, V, F, height, width,
                               wireframe ? 1 : 0, depth.data_ptr<float>(), index.data_ptr<int32_t>(), ws.data_ptr(),
                               (size_t)ws.numel(), stream_of(v)), "rasterize()");
  } else {
    Tensor ws = scratch(drtk_b200_rasterize_workspace_bytes(N, F, height, width, 0), v);
    ok(drtk_b200_rasterize(v.data_ptr<float>(), Strides(v), cptr<int32_t>(vi), Strides(vi), N, V, F, height, width,
                           wireframe ? 1 : 0, 0, depth.data_ptr<float>(), index.data_ptr<int32_t>(), ws.data_ptr(),
                           (size_t)ws.numel(), stream_of(v)), "rasterize()");
  }
  return {depth, index};
}

// ---- render (checks: src/render/render_kernel.cu:285-336) ---------------------------------------------------------
void check_render(const Tensor& v, const Tensor& vi, const Tensor& index_img) {
  TORCH_CHECK(v.defined() && vi.defined() && index_img.defined(), "render(): expected all inputs to be defined");
  TORCH_CHECK(v.device() == vi.device() && v.device() == index_img.device() && v.is_cuda(),
              "render(): expected all inputs to be on same cuda device");
  need_real(v, "render", "v");
  TORCH_CHECK(vi.scalar_type() == at::kInt, "render(): expected vi to have int32 type, but vi has ", vi.dtype());
  TORCH_CHECK(index_img.scalar_type() == at::kInt, "render(): expected index_img to have int32 type, but index_img has ", index_img.dtype());
  TORCH_CHECK(v.dim() == 3 && vi.dim() == 3 && index_img.dim() == 3,
              "render(): expected v.ndim == 3, vi.ndim == 3, index_img.ndim == 3, but got v with sizes ", v.sizes(),
              " and vi with sizes ", vi.sizes(), " and index_img with sizes ", index_img.sizes());
  TORCH_CHECK(v.size(0) == index_img.size(0), "render(): expected v and index_img to have same batch size, but got v with sizes ",
              v.sizes(), " and index_img with sizes ", index_img.sizes());
  TORCH_CHECK(vi.size(0) == v.size(0), "rasterize(): expected first dim of vi to match first dim of v but got ", v.size(0),
              " in first dim of v, and ", vi.size(0), " in the first dim of vi");
  TORCH_CHECK(v.size(2) == 3 && vi.size(2) == 3,
              "render(): expected third dim of v to be of size 3, and third dim of vi to be of size 3, but got ", v.size(2),
              " in the third dim of v, and ", vi.size(2), " in the third dim of vi");
}

std::vector<Tensor> render_cuda(const Tensor& v, const Tensor& vi, const Tensor& index_img) {
  check_render(v, vi, index_img);
  const c10::cuda::CUDAGuard guard(v.device());
  const int64_t N = v.size(0), V = v.size(1), F = vi.size(1), H = index_img.size(1), W = index_img.size(2);
  Tensor depth = at::empty({N, H, W}, v.options());
  Tensor bary = at::empty({N, 3, H, W}, v.options());
  if (is_f64(v))
    ok(drtk_b200_render_forward_f64(v.data_ptr<double>(), Strides(v), cptr<int32_t>(vi), Strides(vi), cptr<int32_t>(index_img),
                                    Strides(index_img), N, V, F, H, W, depth.data_ptr<double>(), bary.data_ptr<double>(),
                                    stream_of(v)), "render()");
  else
    ok(drtk_b200_render_forward(v.data_ptr<float>(), Strides(v), cptr<int32_t>(vi), Strides(vi), cptr<int32_t>(index_img),
                                Strides(index_img), N, V, F, H, W, depth.data_ptr<float>(), bary.data_ptr<float>(),
                                stream_of(v)), "render()");
  return {depth, bary};
}

Tensor render_cuda_backward(const Tensor& v, const Tensor& vi, const Tensor& index_img, const Tensor& grad_depth_in,
                            const Tensor& grad_bary_in) {
  const c10::cuda::CUDAGuard guard(v.device());  // autograd engine thread: set the device ourselves (:388)
  const int64_t N = v.size(0), V = v.size(1), F = vi.size(1), H = index_img.size(1), W = index_img.size(2);
  Tensor grad_v = at::empty({N, V, 3}, v.options());
  const Tensor gd = grad_depth_in.defined() ? grad_depth_in.to(v.scalar_type()) : grad_depth_in;
  const Tensor gb = grad_bary_in.defined() ? grad_bary_in.to(v.scalar_type()) : grad_bary_in;
  const Strides gds = gd.defined() ? Strides(gd) : Strides(v), gbs = gb.defined() ? Strides(gb) : Strides(v);
  if (is_f64(v)) {
    ok(drtk_b200_render_backward_f64(v.data_ptr<double>(), Strides(v), cptr<int32_t>(vi), Strides(vi), cptr<int32_t>(index_img),
                                     Strides(index_img), cptr<double>(gd), gds, cptr<double>(gb), gbs, N, V, F, H, W,
                                     grad_v.data_ptr<double>(), stream_of(v)), "render() backward");
  } else {
    Tensor ws = scratch(drtk_b200_render_backward_workspace_bytes(N, V, F), v);
    ok(drtk_b200_render_backward(v.data_ptr<float>(), Strides(v), cptr<int32_t>(vi), Strides(vi), cptr<int32_t>(index_img),
                                 Strides(index_img), cptr<float>(gd), gds, cptr<float>(gb), gbs, N, V, F, H, W,
                                 grad_v.data_ptr<float>(), ws.data_ptr(), (size_t)ws.numel(), stream_of(v)), "render() backward");
  }
  return grad_v;
}

// ---- interpolate (checks: src/interpolate/interpolate_kernel.cu:459-526) --------------------------------------------
void check_interpolate(const Tensor& attr, const Tensor& vi, const Tensor& index_img, const Tensor& bary_img) {
  TORCH_CHECK(attr.defined() && vi.defined() && index_img.defined() && bary_img.defined(), "interpolate(): expected all inputs to be defined");
  TORCH_CHECK(attr.device() == vi.device() && attr.device() == index_img.device() && attr.device() == bary_img.device(),
              "interpolate(): expected all inputs to be on same device");
  TORCH_CHECK(attr.is_cuda(), "interpolate(): drtk_b200 has no CPU path; expected all inputs to be on a cuda device");
  TORCH_CHECK(attr.is_floating_point(), "interpolate(): expected vert_attributes to have floating point type, but v has ", attr.dtype());
  TORCH_CHECK(attr.dtype() == bary_img.dtype(), "interpolate(): expected vert_attributes and bary_img to have same dtype, but vert_attributes has ",
              attr.dtype(), " and bary_img has ", bary_img.dtype());
  need_real(attr, "interpolate", "vert_attributes");
  TORCH_CHECK(vi.scalar_type() == at::kInt, "interpolate(): expected vi to have int32 type, but vi has ", vi.dtype());
  TORCH_CHECK(index_img.scalar_type() == at::kInt, "interpolate(): expected index_img to have int32 type, but index_img has ", index_img.dtype());
  TORCH_CHECK(attr.dim() == 3 && vi.dim() == 3 && index_img.dim() == 3 && bary_img.dim() == 4,
              "interpolate(): expected vert_attributes.ndim == 3, vi.ndim == 3, index_img.ndim == 3, bary_img.ndim == 4, but got vert_attributes with sizes ",
              attr.sizes(), " and vi with sizes ", vi.sizes(), " and index_img with sizes ", index_img.sizes(),
              " and bary_img with sizes ", bary_img.sizes());
  TORCH_CHECK(attr.size(0) == index_img.size(0) && attr.size(0) == bary_img.size(0),
              "interpolate(): expected vert_attributes, index_img and bary_img to have same batch size, but got vert_attributes with sizes ",
              attr.sizes(), " and index_img with sizes ", index_img.sizes(), " and bary_img with sizes ", bary_img.sizes());
  TORCH_CHECK(vi.size(2) == 3 && bary_img.size(1) == 3,
              "interpolate(): expected last dim of vi to be of size 3, and second dim of bary_img to be of size 3, but got ", vi.size(2),
              " in the last dim of vi, and ", bary_img.size(1), " in the second dim of bary_img");
  TORCH_CHECK(vi.size(0) == attr.size(0), "interpolate(): expected vi to have same first dimension as vert_atrributes, but got ",
              vi.size(0), " in the first dim of vi, and ", attr.size(0), " in the first dim of vert_attributes");
  TORCH_CHECK(index_img.size(1) == bary_img.size(2) && index_img.size(2) == bary_img.size(3),
              "interpolate(): expected H and W dims of index_img and bary_img to match");
}

Tensor interpolate_cuda(const Tensor& attr, const Tensor& vi, const Tensor& index_img, const Tensor& bary_img) {
  check_interpolate(attr, vi, index_img, bary_img);
  const c10::cuda::CUDAGuard guard(attr.device());
  const int64_t N = attr.size(0), V = attr.size(1), C = attr.size(2), F = vi.size(1), H = bary_img.size(2), W = bary_img.size(3);
  Tensor out = at::empty({N, C, H, W}, attr.options());
  if (is_f64(attr))
    ok(drtk_b200_interpolate_forward_f64(attr.data_ptr<double>(), Strides(attr), cptr<int32_t>(vi), Strides(vi),
                                         cptr<int32_t>(index_img), Strides(index_img), bary_img.data_ptr<double>(),
                                         Strides(bary_img), N, V, F, C, H, W, out.data_ptr<double>(), stream_of(attr)), "interpolate()");
  else
    ok(drtk_b200_interpolate_forward(attr.data_ptr<float>(), Strides(attr), cptr<int32_t>(vi), Strides(vi),
                                     cptr<int32_t>(index_img), Strides(index_img), bary_img.data_ptr<float>(),
                                     Strides(bary_img), N, V, F, C, H, W, out.data_ptr<float>(), stream_of(attr)), "interpolate()");
  return out;
}

// only the requested gradients are computed (src/interpolate/interpolate_kernel.cu:610-639)
std::tuple<Tensor, Tensor> interpolate_cuda_backward(const Tensor& grad_out_in, const Tensor& attr, const Tensor& vi,
                                                     const Tensor& index_img, const Tensor& bary_img, bool need_attr,
                                                     bool need_bary) {
  const c10::cuda::CUDAGuard guard(attr.device());
  const int64_t N = attr.size(0), V = attr.size(1), C = attr.size(2), F = vi.size(1), H = bary_img.size(2), W = bary_img.size(3);
  const Tensor g = grad_out_in.to(attr.scalar_type());
  Tensor ga = need_attr ? at::empty({N, V, C}, attr.options()) : Tensor();
  Tensor gb = need_bary ? at::empty({N, 3, H, W}, attr.options()) : Tensor();
  if (is_f64(attr)) {
    ok(drtk_b200_interpolate_backward_f64(g.data_ptr<double>(), Strides(g), attr.data_ptr<double>(), Strides(attr), cptr<int32_t>(vi),
                                          Strides(vi), cptr<int32_t>(index_img), Strides(index_img), bary_img.data_ptr<double>(),
                                          Strides(bary_img), N, V, F, C, H, W, need_attr ? ga.data_ptr<double>() : nullptr,
                                          need_bary ? gb.data_ptr<double>() : nullptr, stream_of(attr)), "interpolate() backward");
  } else {
    Tensor ws = scratch(drtk_b200_interpolate_backward_workspace_bytes(N, F, vi.stride(0)), attr);
    ok(drtk_b200_interpolate_backward(g.data_ptr<float>(), Strides(g), attr.data_ptr<float>(), Strides(attr), cptr<int32_t>(vi),
                                      Strides(vi), cptr<int32_t>(index_img), Strides(index_img), bary_img.data_ptr<float>(),
                                      Strides(bary_img), N, V, F, C, H, W, need_attr ? ga.data_ptr<float>() : nullptr,
                                      need_bary ? gb.data_ptr<float>() : nullptr, ws.data_ptr(), (size_t)ws.numel(),
                                      stream_of(attr)), "interpolate() backward");
  }
  return {ga, gb};
}

// ---- edge_grad_estimator (checks: src/edge_grad/edge_grad_module.cpp:30-112) -----------------------------------------
void check_edge_grad(const Tensor& v_pix, const Tensor& v_pix_img, const Tensor& vi, const Tensor& img, const Tensor& index_img) {
  const char* who = "edge_grad_estimator()";
  TORCH_CHECK(v_pix.defined() && v_pix_img.defined() && vi.defined() && img.defined() && index_img.defined(), who,
              ": expected all inputs to be defined");
  TORCH_CHECK(v_pix.device() == v_pix_img.device() && v_pix.device() == vi.device() && v_pix.device() == img.device() &&
                  v_pix.device() == index_img.device() && v_pix.is_cuda(), who, ": expected all inputs to be on same cuda device");
  TORCH_CHECK(v_pix.is_floating_point() && v_pix_img.is_floating_point() && img.is_floating_point(), who,
              ": expected v_pix, v_pix_img, and img to have floating point type, but v_pix has ", v_pix.dtype(), " v_pix has ",
              v_pix_img.dtype(), " img has ", img.dtype());
  TORCH_CHECK(vi.scalar_type() == at::kInt, who, ": expected vi to have int32 type, but vi has ", vi.dtype());
  TORCH_CHECK(index_img.scalar_type() == at::kInt, who, ": expected index_img to have int32 type, but index_img has ", index_img.dtype());
  TORCH_CHECK(v_pix.dim() == 3 && v_pix_img.dim() == 4 && vi.dim() == 3 && img.dim() == 4 && index_img.dim() == 3, who,
              ": expected v_pix.ndim == 3, v_pix_img.ndim == 4, vi.ndim == 3, img.ndim == 4, index_img.ndim == 3, but got v_pix with sizes ",
              v_pix.sizes(), " and v_pix_img with sizes ", v_pix_img.sizes(), " and vi with sizes ", vi.sizes(), " and img with sizes ",
              img.sizes(), " and index_img with sizes ", index_img.sizes());
  TORCH_CHECK(v_pix.size(0) == v_pix_img.size(0) && v_pix.size(0) == img.size(0) && v_pix.size(0) == index_img.size(0), who,
              ": expected v and index_img to have same batch size, but got v_pix with sizes ", v_pix.sizes(), ", v_pix_img with sizes ",
              v_pix_img.sizes(), ", img with sizes ", img.sizes(), " and index_img with sizes ", index_img.sizes());
  TORCH_CHECK(v_pix.size(2) == 3 && v_pix_img.size(1) == 3 && vi.size(2) == 3, who,
              ": expected third dim of v_pix to be of size 3, and third dim of vi to be of size 3, but got ", v_pix.size(2),
              " in the third dim of v_pix, and ", v_pix_img.size(1), " in the second dim of v_pix_img, and ", vi.size(2),
              " in the third dim of vi");
  TORCH_CHECK(v_pix_img.size(3) == img.size(3) && v_pix_img.size(3) == index_img.size(2) && v_pix_img.size(2) == img.size(2) &&
                  v_pix_img.size(2) == index_img.size(1), who,
              ": expected width and height of v_pix_img, img, and index_img to match, but got size of v_pix_img: ", v_pix_img.sizes(),
              ", size of img: ", img.sizes(), ", size of index_img: ", index_img.sizes());
}

// the forward is the identity on img (:118-137): this is what runs below autograd
Tensor edge_grad_estimator_cuda_fwd(const Tensor& v_pix, const Tensor& v_pix_img, const Tensor& vi, const Tensor& img,
                                    const Tensor& index_img, double /*max_dp_dr*/) {
  check_edge_grad(v_pix, v_pix_img, vi, img, index_img);
  return img;
}

Tensor edge_grad_cuda_backward(const Tensor& v_pix, const Tensor& img_in, const Tensor& index_img, const Tensor& vi,
                               const Tensor& grad_in, double max_dp_dr) {
  const c10::cuda::CUDAGuard guard(v_pix.device());
  need_real(v_pix, "edge_grad_estimator", "v_pix");
  const Tensor img = img_in.to(v_pix.scalar_type()), g = grad_in.to(v_pix.scalar_type());
  const int64_t N = v_pix.size(0), V = v_pix.size(1), F = vi.size(1), C = img.size(1), H = img.size(2), W = img.size(3);
  Tensor out = at::empty({N, 3, H, W}, v_pix.options());
  if (is_f64(v_pix))
    ok(drtk_b200_edge_grad_backward_f64(v_pix.data_ptr<double>(), Strides(v_pix), img.data_ptr<double>(), Strides(img),
                                        cptr<int32_t>(index_img), Strides(index_img), cptr<int32_t>(vi), Strides(vi),
                                        g.data_ptr<double>(), Strides(g), N, V, F, C, H, W, max_dp_dr, out.data_ptr<double>(),
                                        stream_of(v_pix)), "edge_grad_estimator() backward");
  else
    ok(drtk_b200_edge_grad_backward(v_pix.data_ptr<float>(), Strides(v_pix), img.data_ptr<float>(), Strides(img),
                                    cptr<int32_t>(index_img), Strides(index_img), cptr<int32_t>(vi), Strides(vi),
                                    g.data_ptr<float>(), Strides(g), N, V, F, C, H, W, (float)max_dp_dr, out.data_ptr<float>(),
                                    stream_of(v_pix)), "edge_grad_estimator() backward");
  return out;
}

// estimator backward + the C = 3 conduit backward in one kernel -> grad_v_pix [N,V,3]
Tensor edge_grad_fused_cuda_backward(const Tensor& v_pix, const Tensor& img, const Tensor& index_img, const Tensor& vi,
                                     const Tensor& grad, const Tensor& bary_img, double max_dp_dr) {
  const c10::cuda::CUDAGuard guard(v_pix.device());
  if (is_f64(v_pix)) {  // fp64: the two plain kernels back to back
    const Tensor g_img = edge_grad_cuda_backward(v_pix, img, index_img, vi, grad, max_dp_dr);
    return std::get<0>(interpolate_cuda_backward(g_img, v_pix, vi, index_img, bary_img.to(at::kDouble), true, false));
  }
  const Tensor im = img.to(at::kFloat), g = grad.to(at::kFloat), bary = bary_img.to(at::kFloat);
  const int64_t N = v_pix.size(0), V = v_pix.size(1), F = vi.size(1), C = im.size(1), H = im.size(2), W = im.size(3);
  Tensor out = at::empty({N, V, 3}, v_pix.options());
  Tensor ws = scratch(drtk_b200_edge_grad_backward_fused_workspace_bytes(N, F), v_pix);
  ok(drtk_b200_edge_grad_backward_fused(v_pix.data_ptr<float>(), Strides(v_pix), im.data_ptr<float>(), Strides(im),
                                        cptr<int32_t>(index_img), Strides(index_img), cptr<int32_t>(vi), Strides(vi),
                                        g.data_ptr<float>(), Strides(g), bary.data_ptr<float>(), Strides(bary), N, V, F, C, H, W,
                                        (float)max_dp_dr, out.data_ptr<float>(), ws.data_ptr(), (size_t)ws.numel(),
                                        stream_of(v_pix)), "edge_grad_estimator() backward");
  return out;
}

// ---- dispatcher entry points (re-dispatch below Autograd / Autocast) ------------------------------------------------
std::vector<Tensor> rasterize_op(const Tensor& v, const Tensor& vi, int64_t height, int64_t width, bool wireframe) {
  static auto op = c10::Dispatcher::singleton().findSchemaOrThrow("drtk_b200_rasterize_ext::rasterize", "").typed<decltype(rasterize_op)>();
  return op.call(v, vi, height, width, wireframe);
}
std::vector<Tensor> render_op(const Tensor& v, const Tensor& vi, const Tensor& index_img) {
  static auto op = c10::Dispatcher::singleton().findSchemaOrThrow("drtk_b200_render_ext::render", "").typed<decltype(render_op)>();
  return op.call(v, vi, index_img);
}
Tensor interpolate_op(const Tensor& attr, const Tensor& vi, const Tensor& index_img, const Tensor& bary_img) {
  static auto op = c10::Dispatcher::singleton().findSchemaOrThrow("drtk_b200_interpolate_ext::interpolate", "").typed<decltype(interpolate_op)>();
  return op.call(attr, vi, index_img, bary_img);
}
Tensor edge_grad_op(const Tensor& v_pix, const Tensor& v_pix_img, const Tensor& vi, const Tensor& img, const Tensor& index_img, double max_dp_dr) {
  static auto op = c10::Dispatcher::singleton().findSchemaOrThrow("drtk_b200_edge_grad_ext::edge_grad_estimator", "").typed<decltype(edge_grad_op)>();
  return op.call(v_pix, v_pix_img, vi, img, index_img, max_dp_dr);
}
Tensor edge_grad_fused_op(const Tensor& v_pix, const Tensor& vi, const Tensor& bary_img, const Tensor& img, const Tensor& index_img, double max_dp_dr) {
  static auto op = c10::Dispatcher::singleton().findSchemaOrThrow("drtk_b200_edge_grad_ext::edge_grad_estimator_fused", "").typed<decltype(edge_grad_fused_op)>();
  return op.call(v_pix, vi, bary_img, img, index_img, max_dp_dr);
}

// ---- autograd ------------------------------------------------------------------------------------------------
// rasterize: outputs are discrete, nothing flows back (src/rasterize/rasterize_module.cpp:31-52)
struct RasterizeFn : torch::autograd::Function<RasterizeFn> {
  static tensor_list forward(AutogradContext* ctx, const Tensor& v, const Tensor& vi, int64_t height, int64_t width, bool wireframe) {
    ctx->set_materialize_grads(false);
    at::AutoDispatchBelowADInplaceOrView below;
    auto outs = rasterize_op(v, vi, height, width, wireframe);
    ctx->mark_non_differentiable(outs);
    return outs;
  }
  static tensor_list backward(AutogradContext*, const tensor_list&) { return tensor_list(5); }
};

// render: gradient to v only, and only when v required grad at forward time (src/render/render_module.cpp:27-72)
struct RenderFn : torch::autograd::Function<RenderFn> {
  static tensor_list forward(AutogradContext* ctx, const Tensor& v, const Tensor& vi, const Tensor& index_img) {
    // an output nobody differentiates (usually depth_img) arrives as an UNDEFINED gradient and is passed to the
    // launcher as a null pointer; the reference lets autograd materialise it (render_module.cpp:46-64): a zero-fill of
    // [N,H,W] plus a kernel that reads those zeros -- same result, 134 MB + 134 MB of traffic per step at config 4
    ctx->set_materialize_grads(false);
    ctx->save_for_backward({v, vi, index_img});
    ctx->saved_data["v_requires_grad"] = v.requires_grad();
    at::AutoDispatchBelowADInplaceOrView below;
    return render_op(v, vi, index_img);
  }
  static tensor_list backward(AutogradContext* ctx, const tensor_list& grads) {
    if (!ctx->saved_data["v_requires_grad"].toBool()) return tensor_list(3);
    const auto saved = ctx->get_saved_variables();
    if (!grads[0].defined() && !grads[1].defined()) return {at::zeros_like(saved[0]), Tensor(), Tensor()};
    return {render_cuda_backward(saved[0], saved[1], saved[2], grads[0], grads[1]), Tensor(), Tensor()};
  }
};

// interpolate: gradients per requires_grad; an undefined incoming gradient means "nothing" (src/interpolate/interpolate_module.cpp:378-425)
struct InterpolateFn : torch::autograd::Function<InterpolateFn> {
  static tensor_list forward(AutogradContext* ctx, const Tensor& attr, const Tensor& vi, const Tensor& index_img, const Tensor& bary_img) {
    ctx->set_materialize_grads(false);
    ctx->save_for_backward({attr, vi, index_img, bary_img});
    at::AutoDispatchBelowADInplaceOrView below;
    return {interpolate_op(attr, vi, index_img, bary_img)};
  }
  static tensor_list backward(AutogradContext* ctx, const tensor_list& grads) {
    const auto saved = ctx->get_saved_variables();
    const bool need_attr = saved[0].requires_grad(), need_bary = saved[3].requires_grad();
    if ((!need_attr && !need_bary) || !grads[0].defined()) return tensor_list(4);
    auto g = interpolate_cuda_backward(grads[0], saved[0], saved[1], saved[2], saved[3], need_attr, need_bary);
    return {std::get<0>(g), Tensor(), Tensor(), std::get<1>(g)};
  }
};

// edge_grad_estimator: identity on img; dL/d(v_pix_img) from the estimator kernel; skipped when v_pix_img needs no grad
// (src/edge_grad/edge_grad_module.cpp:116-170)
struct EdgeGradFn : torch::autograd::Function<EdgeGradFn> {
  static tensor_list forward(AutogradContext* ctx, const Tensor& v_pix, const Tensor& v_pix_img, const Tensor& vi,
                             const Tensor& img, const Tensor& index_img, double max_dp_dr) {
    ctx->set_materialize_grads(false);
    ctx->save_for_backward({v_pix, img, index_img, vi});
    ctx->saved_data["conduit_requires_grad"] = v_pix_img.requires_grad();
    ctx->saved_data["max_dp_dr"] = max_dp_dr;
    at::AutoDispatchBelowADInplaceOrView below;
    return {edge_grad_op(v_pix, v_pix_img, vi, img, index_img, max_dp_dr)};
  }
  static tensor_list backward(AutogradContext* ctx, const tensor_list& grads) {
    tensor_list out(6);
    out[3] = grads[0];
    if (!ctx->saved_data["conduit_requires_grad"].toBool() || !grads[0].defined()) return out;
    const auto saved = ctx->get_saved_variables();
    out[1] = edge_grad_cuda_backward(saved[0], saved[1], saved[2], saved[3], grads[0], ctx->saved_data["max_dp_dr"].toDouble());
    return out;
  }
};

// fused variant: the gradient goes straight to v_pix [N,V,3]
struct EdgeGradFusedFn : torch::autograd::Function<EdgeGradFusedFn> {
  static tensor_list forward(AutogradContext* ctx, const Tensor& v_pix, const Tensor& vi, const Tensor& bary_img,
                             const Tensor& img, const Tensor& index_img, double max_dp_dr) {
    ctx->set_materialize_grads(false);
    ctx->save_for_backward({v_pix, img, index_img, vi, bary_img});
    ctx->saved_data["v_requires_grad"] = v_pix.requires_grad();
    ctx->saved_data["max_dp_dr"] = max_dp_dr;
    at::AutoDispatchBelowADInplaceOrView below;
    return {edge_grad_fused_op(v_pix, vi, bary_img, img, index_img, max_dp_dr)};
  }
  static tensor_list backward(AutogradContext* ctx, const tensor_list& grads) {
    tensor_list out(6);
    out[3] = grads[0];
    if (!ctx->saved_data["v_requires_grad"].toBool() || !grads[0].defined()) return out;
    const auto saved = ctx->get_saved_variables();
    out[0] = edge_grad_fused_cuda_backward(saved[0], saved[1], saved[2], saved[3], grads[0], saved[4],
                                           ctx->saved_data["max_dp_dr"].toDouble());
    return out;
  }
};

Tensor edge_grad_fused_cuda_fwd(const Tensor& v_pix, const Tensor& vi, const Tensor& bary_img, const Tensor& img,
                                const Tensor& index_img, double /*max_dp_dr*/) {
  check_edge_grad(v_pix, bary_img, vi, img, index_img);  // bary_img has the conduit image's shape [N,3,H,W]
  return img;
}

std::vector<Tensor> rasterize_autograd(const Tensor& v, const Tensor& vi, int64_t h, int64_t w, bool wf) { return RasterizeFn::apply(v, vi, h, w, wf); }
std::vector<Tensor> render_autograd(const Tensor& v, const Tensor& vi, const Tensor& index_img) { return RenderFn::apply(v, vi, index_img); }
Tensor interpolate_autograd(const Tensor& a, const Tensor& vi, const Tensor& i, const Tensor& b) { return InterpolateFn::apply(a, vi, i, b)[0]; }
Tensor edge_grad_autograd(const Tensor& v, const Tensor& vimg, const Tensor& vi, const Tensor& img, const Tensor& i, double m) {
  return EdgeGradFn::apply(v, vimg, vi, img, i, m)[0];
}
Tensor edge_grad_fused_autograd(const Tensor& v, const Tensor& vi, const Tensor& b, const Tensor& img, const Tensor& i, double m) {
  return EdgeGradFusedFn::apply(v, vi, b, img, i, m)[0];
}

// ---- autocast: reduced-precision floats -> fp32 (cached_cast leaves float64 alone), then re-dispatch ------------------
inline Tensor f32(const Tensor& t) { return at::autocast::cached_cast(at::kFloat, t); }
std::vector<Tensor> rasterize_autocast(const Tensor& v, const Tensor& vi, int64_t h, int64_t w, bool wf) {
  c10::impl::ExcludeDispatchKeyGuard no_autocast(c10::DispatchKey::Autocast);
  return rasterize_op(f32(v), vi, h, w, wf);
}
std::vector<Tensor> render_autocast(const Tensor& v, const Tensor& vi, const Tensor& index_img) {
  c10::impl::ExcludeDispatchKeyGuard no_autocast(c10::DispatchKey::Autocast);
  return render_op(f32(v), vi, index_img);
}
Tensor interpolate_autocast(const Tensor& a, const Tensor& vi, const Tensor& i, const Tensor& b) {
  c10::impl::ExcludeDispatchKeyGuard no_autocast(c10::DispatchKey::Autocast);
  return interpolate_op(f32(a), vi, i, f32(b));
}
Tensor edge_grad_autocast(const Tensor& v, const Tensor& vimg, const Tensor& vi, const Tensor& img, const Tensor& i, double m) {
  c10::impl::ExcludeDispatchKeyGuard no_autocast(c10::DispatchKey::Autocast);
  return edge_grad_op(f32(v), f32(vimg), vi, f32(img), i, m);
}
Tensor edge_grad_fused_autocast(const Tensor& v, const Tensor& vi, const Tensor& b, const Tensor& img, const Tensor& i, double m) {
  c10::impl::ExcludeDispatchKeyGuard no_autocast(c10::DispatchKey::Autocast);
  return edge_grad_fused_op(f32(v), vi, f32(b), f32(img), i, m);
}

}  // namespace

TORCH_LIBRARY(drtk_b200_rasterize_ext, m) {
  m.def("rasterize(Tensor v, Tensor vi, int height, int width, bool wireframe) -> Tensor[]");
}
TORCH_LIBRARY_IMPL(drtk_b200_rasterize_ext, Autograd, m) { m.impl("rasterize", &rasterize_autograd); }
TORCH_LIBRARY_IMPL(drtk_b200_rasterize_ext, Autocast, m) { m.impl("rasterize", &rasterize_autocast); }
TORCH_LIBRARY_IMPL(drtk_b200_rasterize_ext, CUDA, m) { m.impl("rasterize", &rasterize_cuda); }
TORCH_LIBRARY_IMPL(drtk_b200_rasterize_ext, CPU, m) { m.impl("rasterize", &rasterize_cuda); }  // raises: no CPU path

TORCH_LIBRARY(drtk_b200_render_ext, m) { m.def("render(Tensor v, Tensor vi, Tensor index_img) -> Tensor[]"); }
TORCH_LIBRARY_IMPL(drtk_b200_render_ext, Autograd, m) { m.impl("render", &render_autograd); }
TORCH_LIBRARY_IMPL(drtk_b200_render_ext, Autocast, m) { m.impl("render", &render_autocast); }
TORCH_LIBRARY_IMPL(drtk_b200_render_ext, CUDA, m) { m.impl("render", &render_cuda); }
TORCH_LIBRARY_IMPL(drtk_b200_render_ext, CPU, m) { m.impl("render", &render_cuda); }  // raises: no CPU path

TORCH_LIBRARY(drtk_b200_interpolate_ext, m) {
  m.def("interpolate(Tensor vert_attributes, Tensor vi, Tensor index_img, Tensor bary_img) -> Tensor");
}
TORCH_LIBRARY_IMPL(drtk_b200_interpolate_ext, Autograd, m) { m.impl("interpolate", &interpolate_autograd); }
TORCH_LIBRARY_IMPL(drtk_b200_interpolate_ext, Autocast, m) { m.impl("interpolate", &interpolate_autocast); }
TORCH_LIBRARY_IMPL(drtk_b200_interpolate_ext, CUDA, m) { m.impl("interpolate", &interpolate_cuda); }
TORCH_LIBRARY_IMPL(drtk_b200_interpolate_ext, CPU, m) { m.impl("interpolate", &interpolate_cuda); }  // raises: no CPU path

TORCH_LIBRARY(drtk_b200_edge_grad_ext, m) {
  m.def("edge_grad_estimator(Tensor v_pix, Tensor v_pix_img, Tensor vi, Tensor img, Tensor index_img, float max_dp_dr=1e4) -> Tensor");
  m.def("edge_grad_estimator_fused(Tensor v_pix, Tensor vi, Tensor bary_img, Tensor img, Tensor index_img, float max_dp_dr=1e4) -> Tensor");
}
TORCH_LIBRARY_IMPL(drtk_b200_edge_grad_ext, Autograd, m) {
  m.impl("edge_grad_estimator", &edge_grad_autograd);
  m.impl("edge_grad_estimator_fused", &edge_grad_fused_autograd);
}
TORCH_LIBRARY_IMPL(drtk_b200_edge_grad_ext, Autocast, m) {
  m.impl("edge_grad_estimator", &edge_grad_autocast);
  m.impl("edge_grad_estimator_fused", &edge_grad_fused_autocast);
}
TORCH_LIBRARY_IMPL(drtk_b200_edge_grad_ext, CUDA, m) {
  m.impl("edge_grad_estimator", &edge_grad_estimator_cuda_fwd);
  m.impl("edge_grad_estimator_fused", &edge_grad_fused_cuda_fwd);
}
TORCH_LIBRARY_IMPL(drtk_b200_edge_grad_ext, CPU, m) {  // raise: no CPU path
  m.impl("edge_grad_estimator", &edge_grad_estimator_cuda_fwd);
  m.impl("edge_grad_estimator_fused", &edge_grad_fused_cuda_fwd);
}
