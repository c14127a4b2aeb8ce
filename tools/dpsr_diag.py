#!/usr/bin/env python
"""DPSR diagnostics: pairwise agreement of ours / reference-GPU / reference-CPU / numpy oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dg-mesh_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import util  # noqa: E402
import test_dpsr  # noqa: E402
from nvdiffrast_utils.dpsr import DPSR  # noqa: E402
from oracle.oracle import dpsr_forward_np  # noqa: E402

ref = util.load_reference_pymodules()
for G, n in ((32, 3000), (64, 20000)):
    V, N = test_dpsr._points(n, G + 1)
    V[:5] = torch.tensor([0.25, 0.5, 0.75])
    gout = torch.randn(G, G, G, generator=torch.Generator().manual_seed(3))
    res = {}
    for name, dev in (("ours", "cuda"), ("ref_gpu", "cuda"), ("ref_cpu", "cpu")):
        Va, Na = V.to(dev).requires_grad_(True), N.to(dev).requires_grad_(True)
        mod = DPSR(res=(G, G, G), sig=3.0) if name == "ours" else ref.dpsr.DPSR(res=(G, G, G), sig=3.0).to(dev)
        phi = mod(Va[None], Na[None])[0]
        (phi * gout.to(dev)).sum().backward()
        res[name] = (phi.detach().cpu(), Na.grad.cpu(), Va.grad.cpu())
    res["numpy"] = (torch.from_numpy(dpsr_forward_np(V.detach().numpy(), N.detach().numpy(), G, 3.0)), None, None)
    names = list(res)
    print(f"G={G} n={n}")
    for i in range(len(names)):
        for j in range(i + 1, len(names)):
            a, b = res[names[i]], res[names[j]]
            line = f"  {names[i]:8s} vs {names[j]:8s} phi {util.rel_err(a[0], b[0]):.2e}"
            ok, info = test_dpsr.affine_close(a[0], b[0])
            line += f"  affine(alpha-1={info[0]-1:+.2e}, beta={info[1]:+.2e}, resid={info[2]:.2e})"
            if a[1] is not None and b[1] is not None:
                line += f"  dN {util.rel_err(a[1], b[1]):.2e} dV {util.rel_err(a[2], b[2]):.2e}"
            print(line)
