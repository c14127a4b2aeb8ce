#!/usr/bin/env python
"""One forward+backward of the kernels that are new in round 2, at training size, for `ncu` captures:
marching cubes (288^3), the mesh rasteriser (800x800 on that mesh: rasterize / interpolate / antialias, both ways),
the Laplacian, and the densify gather (200k Gaussians)."""
import ctypes
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dg-mesh_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import _dgm_lib  # noqa: E402
import meshrast as dr  # noqa: E402
from diso import DiffMC  # noqa: E402
from nvdiffrast_utils.regularizer import laplace_regularizer_const  # noqa: E402

dev = torch.device("cuda")
G = 288
ax = torch.linspace(-1, 1, G, device=dev)
x, y, z = torch.meshgrid(ax, ax, ax, indexing="ij")
phi0 = ((x * x + y * y + z * z).sqrt() - 0.55 + 0.03 * torch.sin(9 * x) * torch.sin(7 * y)).contiguous()
mc = DiffMC(dtype=torch.float32).to(dev)
W = H = 800
f = 1.0 / math.tan(0.6911 / 2)
P = torch.tensor([[f, 0, 0, 0], [0, f, 0, 0], [0, 0, -1.01, -0.2], [0, 0, -1, 0]], device=dev)
Vm = torch.eye(4, device=dev)
Vm[2, 3] = -4.0


def once():
    phi = phi0.clone().requires_grad_(True)
    verts, faces = mc(phi, deform=None, isovalue=0.0)
    v = verts * 2 - 1
    pos = torch.cat([v, torch.ones_like(v[:, :1])], 1) @ (P @ Vm).t()
    tri = faces.int()
    rast, _ = dr.rasterize(None, pos[None], tri, resolution=[H, W])
    topo = dr.edge_opposites(tri, v.shape[0])
    col, _ = dr.interpolate((v * 0.5 + 0.5)[None], rast, tri)
    img = dr.antialias(col, rast, pos[None], tri, topology_hash=topo)
    ones, _ = dr.interpolate(torch.ones_like(v)[None], rast, tri)
    mask = dr.antialias(ones, rast, pos[None], tri, topology_hash=topo)
    (img.mean() + mask.mean() + 100 * laplace_regularizer_const(v, faces)).backward()
    torch.cuda.synchronize()
    return verts.shape[0], faces.shape[0]


def densify():
    lib = _dgm_lib.lib()
    n = 200_000
    g = torch.Generator().manual_seed(0)
    r = lambda *s: torch.randn(*s, generator=g).to(dev)  # noqa: E731
    widths = (3, 3, 45, 1, 3, 4, 3)
    src = [r(n, w) for w in widths]
    m1 = [r(n, w) for w in widths]
    m2 = [r(n, w).abs() for w in widths]
    src[4] = (math.log(0.037) + 1.2 * r(n, 3)).contiguous()
    src[3] = (r(n, 1) * 2 - 3).contiguous()
    nb = _dgm_lib.c_size_t()
    lib.dgd_workspace_size(n, ctypes.byref(nb))
    ws = torch.empty(nb.value, dtype=torch.uint8, device=dev)
    counts = torch.empty(4, dtype=torch.int32, device=dev)
    acc, den = (torch.rand(n, generator=g) * 4e-4).to(dev), torch.randint(0, 3, (n,), generator=g).float().to(dev)
    st = _dgm_lib.stream_ptr()
    lib.dgd_plan(n, acc.data_ptr(), den.data_ptr(), src[4].data_ptr(), src[3].data_ptr(), 0.0002, 0.005, 3.7, 0.01, 1,
                 20.0, ws.data_ptr(), nb.value, counts.data_ptr(), st)
    k, c, s, ch = counts.tolist()
    n_out = k + c + 2 * ch
    stds = torch.empty(2 * s, 3, device=dev)
    lib.dgd_split_stds(n, src[4].data_ptr(), ws.data_ptr(), nb.value, stds.data_ptr(), st)
    samples = torch.normal(torch.zeros_like(stds), stds)
    fields = (_dgm_lib.DgdField * 7)()
    keep = []
    for i, w in enumerate(widths):
        d, d1, d2 = (torch.empty(n_out, w, device=dev) for _ in range(3))
        keep += [d, d1, d2]
        fl = fields[i]
        fl.src, fl.m1_src, fl.m2_src = src[i].data_ptr(), m1[i].data_ptr(), m2[i].data_ptr()
        fl.dst, fl.m1_dst, fl.m2_dst, fl.width, fl.role = d.data_ptr(), d1.data_ptr(), d2.data_ptr(), w, (1, 0, 0, 0, 2, 0, 0)[i]
    lib.dgd_apply(n, 7, fields, src[5].data_ptr(), samples.data_ptr(), ws.data_ptr(), nb.value, st)
    torch.cuda.synchronize()
    return n, n_out


def nearest():
    """anchor_mesh's cross-set nearest neighbour at training size: 200k Gaussians against ~110k face centroids"""
    import anchor
    g = torch.Generator().manual_seed(1)
    q = torch.randn(200_000, 3, generator=g).to(dev)
    r = torch.randn(110_000, 3, generator=g).to(dev)
    d2, idx = anchor.nearest(q, r)
    torch.cuda.synchronize()
    return float(d2.mean())


for _ in range(2):
    print(once(), densify(), nearest())
