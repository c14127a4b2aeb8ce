#!/usr/bin/env python
"""Where does a frame batch spend its time?  Times the forward and backward halves of the bench's
8-frame batch separately (CUDA events, synchronised between the halves) next to the sum of the
stand-alone kernel durations."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dg-mesh_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import bench  # noqa: E402
import diff_gaussian_rasterization as dgr  # noqa: E402

dev = torch.device("cuda")
sc, cams, dpix = bench.make_inputs(dev)
leaves, gflat = bench.flat_params(sc, dev)
bg = torch.ones(3, device=dev)
frames = list(range(bench.FRAMES))
settings = bench.batch_settings(dgr, cams, bg, frames)
dp = torch.stack([d.to(dev) for d in dpix]).contiguous()


def ev():
    return torch.cuda.Event(enable_timing=True)


fw, bw, tot = [], [], []
for it in range(12):
    gflat.zero_()
    torch.cuda.synchronize()
    e0, e1, e2 = ev(), ev(), ev()
    e0.record()
    color, radii = dgr.BatchGaussianRasterizer(settings)(
        means3D=leaves["means3D"], means2D=None, opacities=leaves["opacities"], shs=leaves["shs"],
        scales=leaves["scales"], rotations=leaves["rotations"])
    e1.record()
    torch.cuda.synchronize()
    e1b = ev()
    e1b.record()
    color.backward(dp)
    e2.record()
    torch.cuda.synchronize()
    if it >= 4:
        fw.append(e0.elapsed_time(e1)); bw.append(e1b.elapsed_time(e2))
print(f"streams={os.environ.get('DGMESH_B200_STREAMS', 'default')}  forward {sum(fw)/len(fw):.3f} ms   backward {sum(bw)/len(bw):.3f} ms  "
      f"(8 frames; stand-alone kernel sums: fwd 8x(0.115 binning + 0.096 blend), bwd 8x(0.198 + 0.027))")
