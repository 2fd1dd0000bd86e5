"""Deterministic synthetic scenes for tests and bench.py (BASELINE.json configs).

Pure PyTorch on the CPU generator (so the same seed gives the same scene everywhere); no
dependency on the native library or on the oracle.
"""
import math
from typing import Tuple

import torch as th

# (nx, ny, H, W, N) of the BASELINE.json configs that are jittered grid meshes
CONFIGS = {
    3: dict(nx=51, ny=51, H=1024, W=1024, N=8),      # 5 000 triangles
    4: dict(nx=225, ny=225, H=2048, W=2048, N=8),    # 100 352 triangles  (north star)
    5: dict(nx=709, ny=709, H=4096, W=4096, N=8),    # 1 002 528 triangles, per GPU
}


def grid_topology(nx: int, ny: int, offset: int = 0) -> th.Tensor:
    """Two triangles per grid cell: (a,b,c), (b,d,c) with a=(i,j), b=(i,j+1), c=(i+1,j), d=(i+1,j+1)."""
    i = th.arange(ny - 1).view(-1, 1)
    j = th.arange(nx - 1).view(1, -1)
    a = (i * nx + j).reshape(-1)
    b = a + 1
    c = a + nx
    d = c + 1
    tris = th.stack((th.stack((a, b, c), 1), th.stack((b, d, c), 1)), 1).reshape(-1, 3)
    return (tris + offset).to(th.int32)


def _sheet(nx, ny, H, W, gen, zlo, zhi, angle_deg=0.0):
    cx, cy = (W - 1) / 2.0, (H - 1) / 2.0
    sx, sy = 0.9 * W / (nx - 1), 0.9 * H / (ny - 1)
    gx = (th.arange(nx, dtype=th.float64) - (nx - 1) / 2.0) * sx
    gy = (th.arange(ny, dtype=th.float64) - (ny - 1) / 2.0) * sy
    X = gx.view(1, -1).expand(ny, nx).clone()
    Y = gy.view(-1, 1).expand(ny, nx).clone()
    X += (th.rand((ny, nx), generator=gen, dtype=th.float64) * 0.6 - 0.3) * sx
    Y += (th.rand((ny, nx), generator=gen, dtype=th.float64) * 0.6 - 0.3) * sy
    if angle_deg:
        c, s = math.cos(math.radians(angle_deg)), math.sin(math.radians(angle_deg))
        X, Y = c * X - s * Y, s * X + c * Y
    Z = zlo + (zhi - zlo) * th.rand((ny, nx), generator=gen, dtype=th.float64)
    return th.stack((X + cx, Y + cy, Z), -1).reshape(-1, 3).to(th.float32)


def grid_mesh(nx: int, ny: int, H: int, W: int, N: int, seed: int, overdraw: bool = False,
              device="cpu") -> Tuple[th.Tensor, th.Tensor]:
    """Jittered regular grid mesh: v [N,V,3] float32 (pixel-space xy, camera-space z), vi [F,3] int32.

    Vertices span the central 90 % of the canvas, xy jitter U(-0.3,0.3) cell, z ~ U[2,3); batch
    item b uses generator seed `seed + b`.  overdraw=True adds a second sheet rotated by 7 degrees
    with z ~ U[1.5,3.5) so that occlusion and genuine intersections occur.
    """
    vs = []
    for b in range(N):
        gen = th.Generator().manual_seed(seed + b)
        v = _sheet(nx, ny, H, W, gen, 2.0, 3.0)
        if overdraw:
            v = th.cat((v, _sheet(nx, ny, H, W, gen, 1.5, 3.5, angle_deg=7.0)), 0)
        vs.append(v)
    vi = grid_topology(nx, ny)
    if overdraw:
        vi = th.cat((vi, grid_topology(nx, ny, offset=nx * ny)), 0)
    return th.stack(vs).to(device), vi.to(device)


def config_mesh(config: int, N=None, overdraw=False, device="cpu"):
    c = CONFIGS[config]
    n = c["N"] if N is None else N
    v, vi = grid_mesh(c["nx"], c["ny"], c["H"], c["W"], n, seed=1000 * config, overdraw=overdraw, device=device)
    return v, vi, c["H"], c["W"]


def vertex_attributes(N: int, V: int, C: int, seed: int, device="cpu") -> th.Tensor:
    gen = th.Generator().manual_seed(seed)
    return th.rand((N, V, C), generator=gen, dtype=th.float32).to(device)


def two_triangles(device="cpu"):
    """BASELINE config 2: the six literal vertices of the reference demo (test/two_triangles.py:21-32)."""
    v = th.tensor([[10, 200, 100], [300, 50, 100], [400, 500, 100],
                   [50, 400, 200], [400, 50, 50], [300, 500, 200]], dtype=th.float32)[None]
    vi = th.arange(6, dtype=th.int32).view(2, 3)
    return v.to(device), vi.to(device), 512, 512


def hello_triangle(device="cpu"):
    """BASELINE config 1: the README triangle (README.md:35) on a 512x512 canvas."""
    v = th.tensor([[0, 511, 1], [255, 0, 1], [511, 511, 1]], dtype=th.float32)[None]
    vi = th.tensor([[0, 1, 2]], dtype=th.int32)
    return v.to(device), vi.to(device), 512, 512
