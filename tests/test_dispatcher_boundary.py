"""GPU suite: the dispatcher boundary (csrc/torch_shim.cpp).  The reference's Python layer calls
`th.ops.<name>_ext.<op>(...)` (`drtk/rasterize.py:61-65`, `drtk/render.py:35-39`, `drtk/interpolate.py:47-50`,
`drtk/edge_grad_estimator.py:165-180`); binding it to this library is a change of the namespace string.  The class
below IS that layer with `drtk_b200_` in front of the namespaces -- nothing else of this package's Python host is
involved -- and must give the reference's results (its CUDA kernels from oracle/_ref when they travelled, else this
package's ctypes host)."""
import pytest
import torch as th

import drtk_b200
from drtk_b200 import scenes, torch_ops
from oracle import ref as R
from tests.util import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class RefPythonLayer:
    """The reference's Python layer, verbatim in structure, over `torch.ops.<prefix><name>_ext`."""

    def __init__(self, prefix):
        self.ops = lambda name: getattr(th.ops, prefix + name)

    @staticmethod
    def _exp(vi, n):
        return vi[None].expand(n, -1, -1) if vi.ndim == 2 else vi

    def rasterize(self, v, vi, height, width, wireframe=False):
        return self.ops("rasterize_ext").rasterize(v, self._exp(vi, v.shape[0]), height, width, wireframe)[1]

    def render(self, v, vi, index_img):
        depth_img, bary_img = self.ops("render_ext").render(v, self._exp(vi, v.shape[0]), index_img)
        return depth_img, bary_img

    def interpolate(self, vert_attributes, vi, index_img, bary_img):
        return self.ops("interpolate_ext").interpolate(vert_attributes, self._exp(vi, vert_attributes.shape[0]), index_img, bary_img)

    def edge_grad_estimator(self, v_pix, vi, bary_img, img, index_img, v_pix_img_hook=None, max_dp_dr=1e4):
        vi = self._exp(vi, v_pix.shape[0])
        v_pix_img = self.interpolate(v_pix, vi, index_img, bary_img.detach())  # (drtk/edge_grad_estimator.py:172)
        out = self.ops("edge_grad_ext").edge_grad_estimator(v_pix, v_pix_img, vi, img, index_img, max_dp_dr)
        if v_pix_img_hook is not None:
            v_pix_img.register_hook(v_pix_img_hook)
        return out


def run(api, v, vi, attr, w, H, W, hook):
    vv, aa = v.clone().requires_grad_(True), attr.clone().requires_grad_(True)
    cap = {}
    index = api.rasterize(vv, vi, H, W)
    depth, bary = api.render(vv, vi, index)
    img = api.interpolate(aa, vi, index, bary)
    out = api.edge_grad_estimator(vv, vi, bary, img, index, v_pix_img_hook=(lambda g: cap.__setitem__("g", g.clone())) if hook else None)
    ((out * w).sum() + depth.sum()).backward()
    return dict(index=index, depth=depth.detach(), bary=bary.detach(), img=img.detach(), gv=vv.grad, ga=aa.grad, gpix=cap.get("g"))


@pytest.mark.parametrize("overdraw,hook", [(False, False), (True, True)])
def test_reference_python_layer_on_the_dispatcher_ops(overdraw, hook):
    torch_ops.load()
    v, vi, H, W = scenes.config_mesh(3, N=2, overdraw=overdraw, device=DEV)
    attr = scenes.vertex_attributes(2, v.shape[1], 16, seed=77, device=DEV)
    w = th.rand((2, 16, H, W), device=DEV, generator=th.Generator(device=DEV).manual_seed(5))
    new = run(RefPythonLayer("drtk_b200_"), v, vi, attr, w, H, W, hook)
    ref = run(R if R.available() else drtk_b200, v, vi, attr, w, H, W, hook)
    assert th.equal(new["index"], ref["index"])
    for k, tol in (("depth", 1e-5), ("bary", 1e-5), ("img", 1e-5), ("ga", 5e-5), ("gv", 5e-5)):
        assert_close(new[k].cpu().numpy(), ref[k].cpu().numpy(), rtol=tol, what=k)
    if hook:
        assert_close(new["gpix"].cpu().numpy(), ref["gpix"].cpu().numpy(), rtol=1e-5, scale_rtol=1e-5, what="grad_v_pix_img")


def test_dispatcher_ops_semantics():
    """Autograd contract through torch.ops: rasterize outputs carry no grad, render gives grad to v only when v required
    it, interpolate per requires_grad, the estimator returns img requiring grad; float64 goes to the double kernels."""
    torch_ops.load()
    v, vi, H, W = scenes.config_mesh(3, N=1, device=DEV)
    vi3 = vi[None]
    depth, index = th.ops.drtk_b200_rasterize_ext.rasterize(v.clone().requires_grad_(True), vi3, H, W, False)
    assert not depth.requires_grad and not index.requires_grad and index.dtype == th.int32
    d, b = th.ops.drtk_b200_render_ext.render(v, vi3, index)
    assert not b.requires_grad
    vv = v.clone().requires_grad_(True)
    d, b = th.ops.drtk_b200_render_ext.render(vv, vi3, index)
    assert b.requires_grad
    attr = scenes.vertex_attributes(1, v.shape[1], 4, seed=1, device=DEV)
    img = th.ops.drtk_b200_interpolate_ext.interpolate(attr, vi3, index, b.detach())
    assert not img.requires_grad
    out = th.ops.drtk_b200_edge_grad_ext.edge_grad_estimator_fused(vv, vi3, b.detach(), img, index, 1e4)
    assert out.requires_grad and th.equal(out, img)
    out.sum().backward()
    assert vv.grad is not None and bool(th.isfinite(vv.grad).all())
    d64, i64 = th.ops.drtk_b200_rasterize_ext.rasterize(v.double(), vi3, H, W, False)
    assert d64.dtype == th.float32 and int((i64 != index).sum()) <= 2
    _, b64 = th.ops.drtk_b200_render_ext.render(v.double(), vi3, index)
    assert b64.dtype == th.float64
    with pytest.raises(RuntimeError, match="int32"):
        th.ops.drtk_b200_render_ext.render(v, vi3.long(), index)
