#!/usr/bin/env python
"""Timings of the non-rasterizer rows of SURVEY.md 8(a) beside the reference on the same GPU:
distCUDA2 (a12), DPSR forward+backward (a13), marching cubes forward+backward (a14, no reference:
diso is absent).  Prints one JSON object."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dg-mesh_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import util  # noqa: E402


def timeit(fn, steps=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


out = {}
ref = util.load_reference_pymodules()
dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)

# ---- a12 distCUDA2
from simple_knn._C import distCUDA2  # noqa: E402
import test_knn  # noqa: E402  (the reference extension loader lives with its test)
ref_knn = test_knn._ref_knn()
for n in (100_000, 500_000):
    pts = (torch.randn(n, 3, generator=g) * 0.5).to(dev)
    row = {"ours_ms": timeit(lambda: distCUDA2(pts))}
    row["algorithmic_GBps"] = (12 + 4 + 16) * n / (row["ours_ms"] * 1e-3) / 1e9
    if ref_knn is not None:
        row["ref_ms"] = timeit(lambda: ref_knn.distCUDA2(pts))
        row["speedup"] = row["ref_ms"] / row["ours_ms"]
    out[f"knn_{n}"] = row

# ---- a13 DPSR fwd+bwd at the training resolution
from nvdiffrast_utils.dpsr import DPSR  # noqa: E402
G, n = 288, 200_000
d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=1)
V = (0.5 + 0.19 * d + 0.003 * torch.randn(n, 3, generator=g)).clamp(1e-4, 1 - 1e-4).to(dev)
N = d.to(dev)
gout = torch.randn(G, G, G, generator=g).to(dev)


def dpsr_step(mod):
    Va, Na = V.clone().requires_grad_(True), N.clone().requires_grad_(True)
    phi = mod(Va[None], Na[None])
    (phi[0] * gout).sum().backward()


mine = DPSR(res=(G, G, G), sig=3.0)
row = {"ours_fwd_bwd_ms": timeit(lambda: dpsr_step(mine), steps=5)}
with torch.no_grad():
    row["ours_fwd_ms"] = timeit(lambda: mine(V[None], N[None]), steps=5)
row["algorithmic_GB_fwd"] = 2.7
row["fwd_GBps"] = 2.7 / (row["ours_fwd_ms"] * 1e-3)
if ref is not None:
    theirs = ref.dpsr.DPSR(res=(G, G, G), sig=3.0).to(dev)
    row["ref_fwd_bwd_ms"] = timeit(lambda: dpsr_step(theirs), steps=3, warmup=1)
    with torch.no_grad():
        row["ref_fwd_ms"] = timeit(lambda: theirs(V[None], N[None]), steps=3, warmup=1)
    row["speedup_fwd_bwd"] = row["ref_fwd_bwd_ms"] / row["ours_fwd_bwd_ms"]
out["dpsr_288_200k"] = row

# ---- a14 marching cubes on that field
from diso import DiffMC  # noqa: E402
mc = DiffMC(dtype=torch.float32).to(dev)
with torch.no_grad():
    phi = mine.forward_signed(V[None], N[None], torch.zeros((), device=dev))


def mc_step():
    p = phi.clone().requires_grad_(True)
    v, f = mc(p, deform=None, isovalue=0.0)
    v.sum().backward()
    return v.shape[0], f.shape[0]


nv, nf = mc_step()
row = {"verts": nv, "faces": nf, "fwd_bwd_ms": timeit(mc_step, steps=5)}
with torch.no_grad():
    row["fwd_ms"] = timeit(lambda: mc(phi, deform=None, isovalue=0.0), steps=5)
row["algorithmic_GBps_fwd"] = (4 * G ** 3 + 12 * nv + 12 * nf) / (row["fwd_ms"] * 1e-3) / 1e9


def mc_kernels_only():
    """count + scan + emit + resolve through the C-ABI with fixed capacities and no host wait: the device time
    of the forward (the API call above adds the host's wait for {V, F} and its allocations)"""
    import _dgm_lib
    lib = _dgm_lib.lib()
    nbytes = _dgm_lib.c_size_t()
    lib.dgmc_workspace_size(G, nbytes)
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
    totals = torch.empty(2, dtype=torch.int32, device=dev)
    vc, fc = int(nv * 1.25) + 1024, int(nf * 1.25) + 1024
    v = torch.empty(vc, 3, device=dev)
    f = torch.empty(fc, 3, dtype=torch.int32, device=dev)
    g = phi.contiguous().float().reshape(G, G, G)
    st = _dgm_lib.stream_ptr()

    def run():
        lib.dgmc_count(G, g.data_ptr(), 0.0, ws.data_ptr(), nbytes.value, totals.data_ptr(), None, None, st)
        lib.dgmc_emit(G, g.data_ptr(), 0.0, ws.data_ptr(), nbytes.value, v.data_ptr(), vc, f.data_ptr(), fc, st)
    ms = timeit(run, steps=20)
    assert totals.tolist() == [nv, nf], (totals.tolist(), nv, nf)
    return ms


row["fwd_kernels_only_ms"] = mc_kernels_only()
row["algorithmic_GBps_kernels_only"] = (4 * G ** 3 + 12 * nv + 12 * nf) / (row["fwd_kernels_only_ms"] * 1e-3) / 1e9
out["mc_288"] = row
print(json.dumps(out))
