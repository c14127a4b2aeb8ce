"""simple-knn parity: CUDA distCUDA2 vs the unmodified reference extension and vs a CPU k-d tree."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

import util
from oracle.oracle import knn_mean_dist2


def _ref_knn():
    path = os.path.join(util.REF_DIR, "simple_knn", "_C.so")
    if not os.path.exists(path):
        return None
    if "ref_simple_knn._C" in sys.modules:
        return sys.modules["ref_simple_knn._C"]
    spec = importlib.util.spec_from_file_location("ref_simple_knn._C", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules["ref_simple_knn._C"] = mod
    return mod


def test_oracle_small_cases():
    pts = np.array([[0, 0, 0], [1, 0, 0], [0, 2, 0], [0, 0, 3], [5, 5, 5]], np.float32)
    d = knn_mean_dist2(pts)
    assert np.allclose(d[0], (1 + 4 + 9) / 3) and np.allclose(d[1], (1 + 5 + 10) / 3)
    assert knn_mean_dist2(pts[:2])[0] > 1e37     # fewer than 3 neighbours: FLT_MAX slots remain


@pytest.mark.gpu
@pytest.mark.parametrize("n,kind", [(5, "cube"), (300, "cube"), (4097, "blob"), (100_000, "blob"), (50_000, "dups"),
                                    (20_000, "plane")])
def test_distcuda2_matches_reference_and_kdtree(n, kind):
    from simple_knn._C import distCUDA2
    g = torch.Generator().manual_seed(n)
    if kind == "cube":
        pts = torch.rand(n, 3, generator=g) * 2 - 1
    elif kind == "blob":
        pts = torch.randn(n, 3, generator=g) * torch.tensor([0.5, 0.2, 1.0])
    elif kind == "plane":
        pts = torch.rand(n, 3, generator=g)
        pts[:, 2] = 0.25                      # degenerate extent on one axis
    else:
        base = torch.randn(n // 2, 3, generator=g)
        pts = torch.cat([base, base])          # exact duplicates: distance 0 counts
    ours = distCUDA2(pts.cuda()).cpu().numpy()
    orc = knn_mean_dist2(pts.numpy())
    assert ours.shape == (n,)
    assert np.allclose(ours, orc, rtol=2e-5, atol=1e-12), float(np.abs(ours - orc).max())
    ref = _ref_knn()
    if ref is not None and n >= 4:
        r = ref.distCUDA2(pts.cuda()).cpu().numpy()
        assert np.allclose(ours, r, rtol=1e-6, atol=0), float(np.abs(ours - r).max())


@pytest.mark.gpu
def test_distcuda2_edge_cases():
    from simple_knn._C import distCUDA2
    assert distCUDA2(torch.zeros(0, 3, device="cuda")).shape == (0,)
    one = distCUDA2(torch.zeros(1, 3, device="cuda"))
    assert one.shape == (1,) and float(one[0]) > 1e37
    with pytest.raises(ValueError):
        distCUDA2(torch.zeros(4, 3))
