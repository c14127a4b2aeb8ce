#!/usr/bin/env python
"""Host-side cost of one frame through the reference API (GaussianRasterizer forward + backward): wall time per
call with the GPU running asynchronously, and a cProfile of 300 calls.  `--batch1` profiles the F = 1 batch call
(what one rank of an 8-GPU step runs)."""
import cProfile
import io
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dg-mesh_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import bench  # noqa: E402
import synth  # noqa: E402
import diff_gaussian_rasterization as dgr  # noqa: E402

dev = torch.device("cuda", 0)
sc, cams, dpix = bench.make_inputs(dev)
dpix = [d.to(dev) for d in dpix]
leaves, gflat, _ = bench.flat_params(sc, dev)
bg = torch.ones(3, device=dev)
deltas = bench.make_deltas(sc, dev)
batch1 = "--batch1" in sys.argv
if batch1:
    settings = bench.batch_settings(dgr, cams, bg, [0])
    d1 = {n: v[[0]].contiguous() for n, v in deltas.items()}
    dp1 = torch.stack([dpix[0]]).contiguous()


def call(k):
    if batch1:
        gflat.zero_()
        bench.run_batch(dgr, leaves, settings, dp1, deltas=d1)
    else:
        bench.run_frames(dgr, synth, leaves, cams, dpix, bg, [k % 8], deltas=deltas)


for k in range(20):
    call(k)
torch.cuda.synchronize()
n = 300
t0 = time.perf_counter()
for k in range(n):
    call(k)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e6 * (t1 - t0) / n:.1f} us/call, with final sync {1e6 * (t2 - t0) / n:.1f} us/call")
pr = cProfile.Profile()
pr.enable()
for k in range(n):
    call(k)
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print(s.getvalue()[:6000])
