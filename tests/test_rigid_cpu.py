"""`utils.rigid_utils` (the is_6dof SE(3) helpers) against the reference's file (oracle/_ref/refpy copy):
same names, same results; float64 so the comparison is about the formulas, not rounding."""
import importlib.util
import os

import pytest
import torch

import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


REF = os.path.join(util.REF_DIR, "refpy", "rigid_utils.py")


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref not built")
def test_rigid_utils_match_reference():
    ours = _load(os.path.join(ROOT, "dg-mesh_b200", "utils", "rigid_utils.py"), "ours_rigid")
    theirs = _load(REF, "ref_rigid")
    for name in ("skew", "rp_to_se3", "exp_so3", "exp_se3", "to_homogenous", "from_homogenous"):
        assert callable(getattr(ours, name)) and callable(getattr(theirs, name))
    g = torch.Generator().manual_seed(0)
    n = 200
    w = torch.nn.functional.normalize(torch.randn(n, 3, generator=g, dtype=torch.float64), dim=-1)
    v = torch.randn(n, 3, generator=g, dtype=torch.float64)
    theta = torch.rand(n, 1, generator=g, dtype=torch.float64) * 6 - 3
    assert torch.equal(ours.skew(w), theirs.skew(w))
    assert (ours.exp_so3(w, theta) - theirs.exp_so3(w, theta)).abs().max() < 1e-14
    # the reference builds its constants in fp32 (torch.eye / torch.tensor defaults): compare in fp32 too
    S = torch.cat([w, v], -1).float()
    a, b = ours.exp_se3(S, theta.float()), theirs.exp_se3(S, theta.float())
    assert a.shape == b.shape == (n, 4, 4) and (a - b).abs().max() < 5e-6
    p = torch.randn(n, 3, generator=g)
    assert torch.equal(ours.to_homogenous(p), theirs.to_homogenous(p))
    h = torch.randn(n, 4, generator=g)
    assert torch.equal(ours.from_homogenous(h), theirs.from_homogenous(h))
    # gradients flow through theta and the axis
    S.requires_grad_(True)
    ours.exp_se3(S, theta.float()).sum().backward()
    assert torch.isfinite(S.grad).all()
