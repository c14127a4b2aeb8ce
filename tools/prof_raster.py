#!/usr/bin/env python
"""Small driver for ncu: a few forward+backward frames of config C2 (100k Gaussians, 800x800)
through the public API.  Usage (on the GPU box):
  ncu --set full --clock-control none --import-source on -s 14 -c 14 -o gpurun_out/prof python tools/prof_raster.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dg-mesh_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import bench  # noqa: E402
import synth  # noqa: E402
import diff_gaussian_rasterization as dgr  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda", 0)
sc, cams, dpix = bench.make_inputs(dev)
dpix = [d.to(dev) for d in dpix]
leaves, gflat, _ = bench.flat_params(sc, dev)
bg = torch.ones(3, device=dev)
bench.run_frames(dgr, synth, leaves, cams, dpix, bg, list(range(frames)))
torch.cuda.synchronize()
print("done", frames, "frames")
