#!/usr/bin/env python
"""Host side of the full C3 train step (tools/train_step.py, this repo's arm): cProfile of 10 steps sorted by own
time, to find blocking calls and Python overhead between the ~500 kernel launches of a step."""
import cProfile
import io
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402
import train_step as ts  # noqa: E402

S = ts.build("ours", 200_000, 288, torch.device("cuda"))
for _ in range(4):
    ts.full_step(S)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    ts.full_step(S)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host {1e3 * (t1 - t0) / 10:.2f} ms/step, with final sync {1e3 * (t2 - t0) / 10:.2f} ms/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    ts.full_step(S)
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(45)
print(s.getvalue()[:9000])
