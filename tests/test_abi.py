"""CPU suite: the C-ABI library loads and exports every symbol include/drtk_b200.h declares.
No compute calls (no GPU here) -- only loading, symbol resolution and argument-error paths."""
import ctypes
import os
import re

import pytest

from tests.util import ROOT

HEADER = os.path.join(ROOT, "include", "drtk_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(drtk_b200_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_path():
    names = declared_functions()
    for need in ("drtk_b200_rasterize", "drtk_b200_render_forward", "drtk_b200_render_backward",
                 "drtk_b200_interpolate_forward", "drtk_b200_interpolate_backward",
                 "drtk_b200_edge_grad_backward"):
        assert need in names


def test_library_exports_every_declared_symbol():
    import drtk_b200
    from drtk_b200 import _lib
    path = drtk_b200.native_library_path()
    assert os.path.exists(path), "build the library first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(path)
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in include/drtk_b200.h but not exported"
        assert name in _lib.PROTOTYPES, f"{name} has no ctypes prototype in drtk_b200/_lib.py"
    assert set(_lib.PROTOTYPES) == set(declared_functions())


def test_abi_version_and_error_strings():
    from drtk_b200 import _lib
    lib = _lib.load()
    assert lib.drtk_b200_abi_version() == _lib.ABI_VERSION
    assert b"workspace" in lib.drtk_b200_error_string(-2)
    assert lib.drtk_b200_error_string(0) == b"success"
    assert lib.drtk_b200_rasterize_workspace_bytes(8, 100352, 2048, 2048, 0) > 4 * 8 * 100352 * 4
    assert lib.drtk_b200_rasterize_workspace_bytes(1, 10, 64, 64, 1) >= 64 * 64 * 8


def test_argument_errors_do_not_need_a_gpu():
    from drtk_b200 import _lib
    lib = _lib.load()
    s3 = (ctypes.c_int64 * 3)(0, 3, 1)
    # null outputs / negative sizes are invalid; both return before any CUDA call (wireframe or not)
    assert lib.drtk_b200_rasterize(None, s3, None, s3, 1, 1, 1, 8, 8, 1, 0, None, None, None, 0, None) == -1
    assert lib.drtk_b200_rasterize(None, s3, None, s3, 1, 1, 1, -8, 8, 0, 0, None, None, None, 0, None) == -1
    # empty problems are a successful no-op
    assert lib.drtk_b200_render_forward(None, s3, None, s3, None, s3, 0, 0, 0, 0, 0, None, None, None) == 0


def test_dispatcher_boundary_registers_the_reference_schemas():
    """csrc/torch_shim.cpp: the reference's op schemas under drtk_b200_*_ext, with Autograd / Autocast / CUDA kernels
    and a CPU key that only raises (no CPU path, the reference's wording)."""
    import torch
    from drtk_b200 import torch_ops
    torch_ops.build()
    torch_ops.load()
    want = {
        "drtk_b200_rasterize_ext::rasterize": "(Tensor v, Tensor vi, int height, int width, bool wireframe) -> Tensor[]",
        "drtk_b200_render_ext::render": "(Tensor v, Tensor vi, Tensor index_img) -> Tensor[]",
        "drtk_b200_interpolate_ext::interpolate": "(Tensor vert_attributes, Tensor vi, Tensor index_img, Tensor bary_img) -> Tensor",
        "drtk_b200_edge_grad_ext::edge_grad_estimator":
            "(Tensor v_pix, Tensor v_pix_img, Tensor vi, Tensor img, Tensor index_img, float max_dp_dr=10000.) -> Tensor",
    }
    for name, sig in want.items():
        ns, op = name.split("::")
        schema = str(getattr(getattr(torch.ops, ns), op).default._schema)
        assert schema == name + sig, schema
        for key in ("CUDA", "Autograd", "AutocastCUDA"):
            assert torch._C._dispatch_has_kernel_for_dispatch_key(name, key), (name, key)
    # a CPU tensor fails loudly (no CPU fallback), with the reference's message
    with pytest.raises(RuntimeError, match="same cuda device"):
        torch.ops.drtk_b200_render_ext.render(torch.zeros(1, 3, 3), torch.zeros(1, 1, 3, dtype=torch.int32),
                                              torch.zeros(1, 4, 4, dtype=torch.int32))


def test_rasterize_workspace_covers_its_lists():
    """No GPU: the workspace query must cover every list the bin kernels may fill in the worst case (records of small
    triangles: 4 per triangle x 80 B; medium entries: 4 per triangle x 20 B; large entries: 1 per triangle x 20 B; two
    uint32 per tile and per 256-px super-tile) and the packed 64-bit image of the wireframe / validation paths, and grow
    with the batch."""
    from drtk_b200 import _lib
    lib = _lib.load()
    for N, F, H, W in ((1, 2, 512, 512), (8, 100352, 2048, 2048), (3, 7, 33, 65), (1, 0, 16, 16)):
        tiles = ((H + 31) // 32) * ((W + 31) // 32)
        supers = ((H + 255) // 256) * ((W + 255) // 256)
        need = N * F * (4 * 80 + 4 * 20 + 20) + N * (tiles + supers) * 8 + N * 4
        got = lib.drtk_b200_rasterize_workspace_bytes(N, F, H, W, 0)
        assert got >= need, (N, F, H, W, got, need)
        assert got >= 8 * N * H * W  # wireframe mode shares the buffer
        assert lib.drtk_b200_rasterize_workspace_bytes(N, F, H, W, 1) >= 8 * N * H * W
        assert lib.drtk_b200_rasterize_workspace_bytes(N + 1, F, H, W, 0) >= got
    assert lib.drtk_b200_rasterize_workspace_bytes(0, 5, 64, 64, 0) == 0
