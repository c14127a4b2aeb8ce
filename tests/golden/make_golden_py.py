#!/usr/bin/env python
"""Golden vectors from the reference's pure-PyTorch modules, generated in the build container by
importing them from /root/reference (through oracle/_ref/refpy, the verbatim copies made by
oracle/build_ref.py) and running them on the CPU in fp32:

    python tests/golden/make_golden_py.py          # writes tests/golden/{mlp,dpsr}_*.npz

* mlp_<Class>_<blender>.npz : seed -> default-initialised module (parameters are NOT stored: the
  drop-in modules construct their layers in the same order, so the same seed gives the same
  parameters; a checksum guards that), 257 points, outputs and the gradients of sum(outputs * g).
* dpsr_g32.npz : 600 oriented points (some exactly on grid nodes), phi, gradients of sum(phi * g).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "dg-mesh_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import util  # noqa: E402

MLP_CASES = [("DeformNetworkNormal", True, 11), ("DeformNetworkNormal", False, 12), ("DeformNetwork", True, 13),
             ("DeformNetworkNormalSep", True, 14), ("AppearanceNetwork", True, 15)]


def mlp_inputs(n, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 3, generator=g) * 0.6
    t = torch.rand(n, 1, generator=g)
    gout = torch.randn(n, 16, generator=g)
    return x, t, gout


def param_checksum(mod):
    return float(sum((p.detach().double() * (i + 1)).sum() for i, p in enumerate(mod.parameters())))


def main():
    ref = util.load_reference_pymodules()
    assert ref is not None, "run oracle/build_ref.py first (copies the reference modules to oracle/_ref/refpy)"
    torch.set_num_threads(4)
    for cls, blender, seed in MLP_CASES:
        torch.manual_seed(seed)
        net = getattr(ref.time_utils, cls)(is_blender=blender)
        if cls == "DeformNetworkNormalSep":      # zero-initialised head (time_utils.py:247-249): make it do work
            torch.manual_seed(seed + 100)
            torch.nn.init.normal_(net.gaussian_normal.weight, std=0.05)
        x, t, gout = mlp_inputs(257, seed)
        xa = x.clone().requires_grad_(True)
        out = net(xa, t)
        y = torch.cat(out if isinstance(out, tuple) else (out,), -1)
        (y * gout[:, :y.shape[1]]).sum().backward()
        grads = {k.replace(".", "__"): p.grad.numpy() for k, p in net.named_parameters()}
        np.savez_compressed(os.path.join(HERE, f"mlp_{cls}_{int(blender)}.npz"), seed=seed, y=y.detach().numpy(),
                            dx=xa.grad.numpy(), checksum=param_checksum(net),
                            **{f"g_{k}": v for k, v in grads.items() if v.size <= 4096})   # biases + small heads
    # DPSR
    G, n = 32, 600
    g = torch.Generator().manual_seed(21)
    V = (0.5 + 0.18 * torch.randn(n, 3, generator=g)).clamp(1e-6, 1 - 1e-6)
    V[:6] = torch.tensor([0.25, 0.5, 0.75])
    N = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=1)
    gout = torch.randn(G, G, G, generator=g)
    Va, Na = V.clone().requires_grad_(True), N.clone().requires_grad_(True)
    phi = ref.dpsr.DPSR(res=(G, G, G), sig=3.0)(Va[None], Na[None])[0]
    (phi * gout).sum().backward()
    np.savez_compressed(os.path.join(HERE, "dpsr_g32.npz"), V=V.numpy(), N=N.numpy(), gout=gout.numpy(),
                        phi=phi.detach().numpy(), dV=Va.grad.numpy(), dN=Na.grad.numpy(), sig=3.0)
    print("written to", HERE)


if __name__ == "__main__":
    main()
