#!/usr/bin/env python
"""Build the UNMODIFIED reference DRTK hot-path extensions into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (drtk_b200/) may import or load
anything produced here; only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs do.

What it does
------------
Compiles the reference's own sources *where they lie* under /root/reference (nothing is
copied into this repository) with the reference's own flags (setup.py:22-24, :36-62:
host `-O3 --fast-math -std=c++17`, device `-O3 --use_fast_math -std=c++20`) for sm_100,
one shared object per extension:

    oracle/_ref/rasterize_ext.so   <- src/rasterize/{rasterize_module.cpp,rasterize_kernel.cu,rasterize_kernel_cpu.cpp}
    oracle/_ref/render_ext.so      <- src/render/...
    oracle/_ref/interpolate_ext.so <- src/interpolate/...
    oracle/_ref/edge_grad_ext.so   <- src/edge_grad/...
    oracle/_ref/grid_scatter_ext.so, mipmap_grid_sampler_ext.so <- src/grid_scatter/..., src/mipmap_grid_sampler/...

The reference's setup.py is NOT run (it builds four unrelated extensions too and wants a
writable source tree); ninja + nvcc/g++ are driven through torch.utils.cpp_extension.load
with an explicit build directory.  `-DNO_PYBIND` drops the empty pybind module
(src/*/..._module.cpp `#ifndef NO_PYBIND`) so the result is a plain TORCH_LIBRARY .so that
`torch.ops.load_library` can open on the GPU box, where /root/reference does not exist.

The .so files contain both the reference CUDA kernels (bit-exact oracle on the B200) and
the reference CPU twins (logic oracle in the CPU container, and the `--impl reference`
CPU arm of bench.py).

Usage:  python oracle/build_ref.py [--ref /root/reference] [--only rasterize,render]
"""
import argparse
import os
import shutil
import sys
import time

EXTS = {
    "rasterize": ["rasterize_module.cpp", "rasterize_kernel.cu", "rasterize_kernel_cpu.cpp"],
    "render": ["render_module.cpp", "render_kernel.cu", "render_kernel_cpu.cpp"],
    "interpolate": ["interpolate_module.cpp", "interpolate_kernel.cu", "interpolate_kernel_cpu.cpp"],
    "edge_grad": ["edge_grad_module.cpp", "edge_grad_kernel.cu", "edge_grad_kernel_cpu.cpp"],
    # the samplers either side of interpolate in real pipelines (CUDA only, no CPU twins)
    "grid_scatter": ["grid_scatter_module.cpp", "grid_scatter_kernel.cu"],
    "mipmap_grid_sampler": ["mipmap_grid_sampler_module.cpp", "mipmap_grid_sampler_kernel.cu"],
}

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")


def build(ref_root: str, only=None, verbose: bool = False) -> None:
    if not os.path.isdir(os.path.join(ref_root, "src", "rasterize")):
        raise SystemExit(f"reference sources not found under {ref_root}")
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ.setdefault("MAX_JOBS", str(os.cpu_count() or 4))
    from torch.utils import cpp_extension

    os.makedirs(OUT, exist_ok=True)
    inc = os.path.join(ref_root, "src", "include")

    def build_one(name):
        files = EXTS[name]
        so_final = os.path.join(OUT, f"{name}_ext.so")
        if os.path.exists(so_final):
            print(f"[build_ref] {so_final} exists, skipping")
            return
        t0 = time.time()
        bdir = os.path.join(OUT, "build", name)
        os.makedirs(bdir, exist_ok=True)
        srcs = [os.path.join(ref_root, "src", name, f) for f in files]
        cpp_extension.load(
            name=f"{name}_ext",
            sources=srcs,
            extra_include_paths=[inc, os.path.join(ref_root, "src", name)],
            extra_cflags=["-O3", "--fast-math", "-std=c++17", "-DNO_PYBIND", "-w"],
            extra_cuda_cflags=["-O3", "--use_fast_math", "-std=c++20", "-DNO_PYBIND", "-w"],
            build_directory=bdir,
            is_python_module=False,
            verbose=verbose,
        )
        shutil.copy2(os.path.join(bdir, f"{name}_ext.so"), so_final)
        print(f"[build_ref] built {so_final} in {time.time() - t0:.0f}s")

    # every extension has only two or three translation units, so one ninja run cannot fill the cores: the
    # extensions are compiled side by side (each in its own build directory)
    todo = [n for n in EXTS if not only or n in only]
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(len(todo), 4) or 1) as pool:
        for f in [pool.submit(build_one, n) for n in todo]:
            f.result()
    # the intermediate objects are large and not needed on the GPU box
    shutil.rmtree(os.path.join(OUT, "build"), ignore_errors=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--only", default="")
    ap.add_argument("-v", action="store_true")
    a = ap.parse_args()
    build(a.ref, set(a.only.split(",")) if a.only else None, a.v)
    sys.exit(0)
