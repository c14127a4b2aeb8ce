#!/usr/bin/env python
"""One forward+backward of each non-rasterizer path at its training size, for `ncu` captures:
DPSR (G = 288, 200k points), marching cubes on that field, distCUDA2 (100k points), the fused image loss
(800x800) and densify_and_prune-sized gathers are exercised once after one warm-up round.
   ncu --set full --clock-control none --kernel-name-base demangled -k regex:dgm:: -s <warm-up launches> ...
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dg-mesh_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)
from simple_knn._C import distCUDA2  # noqa: E402
from nvdiffrast_utils.dpsr import DPSR  # noqa: E402
from diso import DiffMC  # noqa: E402
from utils.loss_utils import image_loss  # noqa: E402

G, n = 288, 200_000
d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=1)
V = (0.5 + 0.19 * d + 0.003 * torch.randn(n, 3, generator=g)).clamp(1e-4, 1 - 1e-4).to(dev)
N = d.to(dev)
pts = (torch.randn(100_000, 3, generator=g) * 0.5).to(dev)
img = torch.rand(3, 800, 800, generator=g).to(dev)
gt = torch.rand(3, 800, 800, generator=g).to(dev)
dpsr = DPSR(res=(G, G, G), sig=3.0)
mc = DiffMC(dtype=torch.float32).to(dev)
thres = torch.zeros((), device=dev, requires_grad=True)


def once():
    Va, Na = V.clone().requires_grad_(True), N.clone().requires_grad_(True)
    phi = dpsr.forward_signed(Va[None], Na[None], thres)
    v, f = mc(phi, deform=None, isovalue=0.0)
    (v.sum()).backward()
    distCUDA2(pts)
    x = img.clone().requires_grad_(True)
    image_loss(x, gt, 0.2).backward()
    torch.cuda.synchronize()
    return v.shape[0], f.shape[0]


rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for _ in range(rounds):
    print(once())
