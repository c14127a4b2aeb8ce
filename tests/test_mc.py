"""Marching cubes: `diso` (the reference's third-party MC) is unavailable, so parity is defined
geometrically: closed, consistently oriented 2-manifolds with the right Euler
characteristic, vertices exactly on the iso-level of the edge-interpolated field; the CUDA path must
produce the same vertex set and the same triangles (as sets) as the numpy restatement, and its
backward must match finite differences."""
import numpy as np
import pytest
import torch

from oracle.oracle import marching_cubes_np, mesh_topology


def field(G, kind):
    ax = np.linspace(-1, 1, G, dtype=np.float32)
    x, y, z = np.meshgrid(ax, ax, ax, indexing="ij")
    if kind == "sphere":
        return (np.sqrt(x * x + y * y + z * z) - 0.62).astype(np.float32)
    if kind == "torus":
        return (np.sqrt((np.sqrt(x * x + y * y) - 0.55) ** 2 + z * z) - 0.23).astype(np.float32)
    if kind == "two":
        a = np.sqrt((x - 0.4) ** 2 + y * y + z * z) - 0.3
        b = np.sqrt((x + 0.4) ** 2 + (y - 0.1) ** 2 + z * z) - 0.25
        return np.minimum(a, b).astype(np.float32)
    rng = np.random.default_rng(0)           # smooth noise: exercises ambiguous faces
    f = rng.standard_normal((G, G, G)).astype(np.float32)
    for _ in range(2):
        f = (f + np.roll(f, 1, 0) + np.roll(f, 1, 1) + np.roll(f, 1, 2)) / 4
    f[0], f[-1], f[:, 0], f[:, -1], f[:, :, 0], f[:, :, -1] = 1, 1, 1, 1, 1, 1   # close the surface
    return f


@pytest.mark.parametrize("kind,chi", [("sphere", 2), ("torus", 0), ("two", 4), ("noise", None)])
def test_oracle_surfaces_are_closed_oriented_manifolds(kind, chi):
    phi = field(20, kind)
    v, f = marching_cubes_np(phi, 0.0)
    assert len(v) and len(f)
    euler, manifold, oriented = mesh_topology(v, f)
    assert manifold and oriented
    if chi is not None:
        assert euler == chi
    assert v.min() >= 0 and v.max() <= 1
    # orientation: normals point from inside (phi < iso) to outside -> positive volume for the sphere
    if kind == "sphere":
        p = v[f]
        vol = np.einsum("ij,ij->i", p[:, 0], np.cross(p[:, 1], p[:, 2])).sum() / 6
        assert vol > 0 and abs(vol - 4 / 3 * np.pi * (0.62 / 2) ** 3) < 0.02 * vol


def test_generated_table_is_committed():
    import subprocess
    import sys
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    assert subprocess.call([sys.executable, os.path.join(root, "tools", "gen_mc_tables.py"), "--check"]) == 0


def _canon(verts, faces):
    """Order-independent description: sorted quantised vertices, triangles as sorted coordinate triples."""
    q = np.round(np.asarray(verts, np.float64) * 1e5).astype(np.int64)
    tri = q[np.asarray(faces)]                                  # [F,3,3]
    # rotate each triangle so its lexicographically smallest vertex comes first (keeps orientation)
    keys = tri[..., 0] * 4_000_000_000_000 + tri[..., 1] * 2_000_000 + tri[..., 2]
    r = keys.argmin(1)
    idx = (r[:, None] + np.arange(3)[None]) % 3
    tri = np.take_along_axis(tri, idx[..., None], 1).reshape(len(tri), 9)
    return np.unique(q, axis=0), tri[np.lexsort(tri.T[::-1])]


@pytest.mark.gpu
@pytest.mark.parametrize("kind,G", [("sphere", 24), ("torus", 32), ("noise", 20)])
def test_cuda_mc_equals_oracle_as_sets(kind, G):
    from diso import DiffMC
    phi = field(G, kind)
    v, f = DiffMC(dtype=torch.float32).cuda()(torch.from_numpy(phi).cuda(), deform=None, isovalue=0.0)
    assert f.dtype == torch.int64 and v.dtype == torch.float32
    vo, fo = marching_cubes_np(phi, 0.0)
    assert v.shape == vo.shape and f.shape == fo.shape
    a, b = _canon(v.cpu().numpy(), f.cpu().numpy()), _canon(vo, fo)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    euler, manifold, oriented = mesh_topology(v.cpu().numpy(), f.cpu().numpy())
    assert manifold and oriented


@pytest.mark.gpu
def test_cuda_mc_backward_matches_finite_differences_and_full_size():
    from diso import DiffMC
    mc = DiffMC(dtype=torch.float32).cuda()
    G = 16
    phi = torch.from_numpy(field(G, "sphere")).cuda().double()
    w = torch.randn(4096, 3, generator=torch.Generator().manual_seed(0)).cuda()

    def loss(p):
        v, _ = mc(p.float(), isovalue=0.0)
        return (v.double() * w[:v.shape[0]].double()).sum()

    p = phi.clone().float().requires_grad_(True)
    loss(p).backward()
    g = p.grad.double()
    idx = torch.nonzero(g.abs() > 1e-6)[:40]
    for (i, j, k) in idx.tolist():
        e = 1e-3
        pp, pm = phi.clone(), phi.clone()
        pp[i, j, k] += e
        pm[i, j, k] -= e
        fd = float((loss(pp) - loss(pm)) / (2 * e))
        assert abs(fd - float(g[i, j, k])) < 2e-2 * max(1.0, abs(fd)), (i, j, k, fd, float(g[i, j, k]))
    # BASELINE grid size: 288^3 sphere, closed surface, V - E + F = 2
    G = 288
    ax = torch.linspace(-1, 1, G, device="cuda")
    x, y, z = torch.meshgrid(ax, ax, ax, indexing="ij")
    big = (x * x + y * y + z * z).sqrt() - 0.7
    v, f = mc(big, isovalue=0.0)
    euler, manifold, oriented = mesh_topology(v.cpu().numpy(), f.cpu().numpy())
    assert euler == 2 and manifold and oriented and v.shape[0] > 100_000
    r = ((v * 2 - 1) ** 2).sum(1).sqrt()
    assert float((r - 0.7).abs().max()) < 2e-4
