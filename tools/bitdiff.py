#!/usr/bin/env python
"""Bit-level agreement of every intermediate with the UNMODIFIED reference on config C2
(100k Gaussians, 800x800): fraction of bit-identical entries per array."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dg-mesh_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import synth  # noqa: E402
from test_raster_gpu import _cam_cuda, _cuda, run_ours, run_ref  # noqa: E402

sc = _cuda(synth.gaussian_scene(n=100_000, seed=0))
for az in (30.0, 135.0):
    cam = _cam_cuda(synth.look_at_camera(azimuth_deg=az, width=800, height=800))
    bg = torch.ones(3, device="cuda")
    dpix = torch.randn(3, 800, 800, generator=torch.Generator().manual_seed(1)).cuda()
    a, b = run_ours(sc, cam, bg, 3, False, False, dpix), run_ref(sc, cam, bg, 3, False, False, dpix)
    vis = b["radii"] > 0
    print(f"azimuth {az}: R ours {a['R']} ref {b['R']}")
    for k in ("depths", "means2D", "cov3D", "conic_opacity", "rgb"):
        x, y = a[k][vis].contiguous().view(torch.int32), b[k][vis].contiguous().view(torch.int32)
        print(f"  {k:14s} bit-identical {float((x == y).float().mean()):.6f}")
    for k in ("radii", "tiles_touched", "point_list", "point_list_keys", "ranges", "n_contrib"):
        same = float((a[k] == b[k]).float().mean()) if a[k].shape == b[k].shape else -1.0
        print(f"  {k:14s} equal         {same:.6f}")
    for k in ("final_T", "color"):
        x, y = a[k].contiguous().view(torch.int32), b[k].contiguous().view(torch.int32)
        print(f"  {k:14s} bit-identical {float((x == y).float().mean()):.6f}  max abs diff {float((a[k]-b[k]).abs().max()):.3e}")
    for k, g in b["grads"].items():
        if g is not None:
            d = (a["grads"][k] - g).abs().max() / (g.abs().max() + 1e-30)
            print(f"  grad {k:10s} max-norm rel err {float(d):.3e}")
