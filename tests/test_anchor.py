"""anchor_mesh on the device (dg-mesh_b200/anchor.py, csrc/knn.cu nearest_kernel) against the reference's own
`GaussianModelDPSRDynamicAnchor.anchor_mesh` (oracle/_ref/dgmesh/scene/gaussian_model_dpsr_dynamic_anchor.py:736-828,
imported unmodified; its third-party calls -- trimesh, pytorch3d.knn_points / axis_angle_to_quaternion -- are the
restatements in tools/harness_stubs.py) under the same seed: same survivors, same new Gaussians, same Adam
state, same loss and the same gradient into the deformation network."""
import math

import pytest
import torch

import util
from test_densify import ATTRS, _build, _reference_model_class



def _sphere(n_lat=40, n_lon=64, radius=0.5, device="cuda"):
    th = torch.linspace(0.05, math.pi - 0.05, n_lat)
    ph = torch.linspace(0, 2 * math.pi, n_lon + 1)[:-1]
    T, Ph = torch.meshgrid(th, ph, indexing="ij")
    v = radius * torch.stack([T.sin() * Ph.cos(), T.sin() * Ph.sin(), T.cos()], -1).reshape(-1, 3)
    idx = torch.arange(n_lat * n_lon).reshape(n_lat, n_lon)
    a, b = idx[:-1], idx[1:]
    ar, br = a.roll(-1, 1), b.roll(-1, 1)
    f = torch.cat([torch.stack([a, b, ar], -1).reshape(-1, 3), torch.stack([ar, b, br], -1).reshape(-1, 3)])
    return v.to(device), f.to(device)


@pytest.mark.gpu
def test_nearest_kernel_matches_brute_force():
    import anchor
    g = torch.Generator().manual_seed(0)
    for Q, R in ((1, 1), (1000, 7), (5000, 3001), (257, 1024), (3, 5000)):
        q, r = torch.randn(Q, 3, generator=g).cuda(), torch.randn(R, 3, generator=g).cuda()
        r[R // 2] = r[0]                                           # an exact tie: the lower index must win
        d2, idx = anchor.nearest(q, r)
        diff = q[:, None, :].double() - r[None, :, :].double()
        want = (diff * diff).sum(-1)
        wd, wi = want.min(1)
        assert torch.allclose(d2.double(), wd, rtol=1e-5, atol=1e-12)
        near_tie = (want.topk(2, dim=1, largest=False).values.diff(dim=1).squeeze(1).abs() < 1e-6 * (1 + wd)) if R > 1 \
            else torch.zeros(Q, dtype=torch.bool, device="cuda")
        assert torch.equal(idx[~near_tie], wi[~near_tie])
        assert not bool((idx == R // 2).any()) or R == 1           # duplicates resolve to index 0
    with pytest.raises(ValueError):
        anchor.nearest(torch.zeros(4, 3), torch.zeros(4, 3))         # CPU tensors: no fallback


@pytest.mark.gpu
@pytest.mark.parametrize("P,bs,increase_bs", [(20000, 256, 1024), (3000, 16, 50)])
def test_anchor_mesh_equals_reference(P, bs, increase_bs):
    cls = _reference_model_class()
    if cls is None:
        pytest.skip("oracle/_ref/dgmesh missing")
    import anchor
    import scene
    assert cls.anchor_mesh is anchor.anchor_mesh                         # the launcher swapped it in
    verts, faces = _sphere()
    cent = verts[faces].mean(1)
    torch.manual_seed(11)
    deform = scene.DeformModelNormal(is_blender=True)
    deform_back = scene.DeformModelNormal(is_blender=True, model_name="deform_back")
    for d in (deform, deform_back):                                       # small but non-zero offsets
        with torch.no_grad():
            for h in (d.deform.gaussian_warp, d.deform.gaussian_scaling, d.deform.gaussian_rotation, d.deform.gaussian_normal):
                h.weight.mul_(0.02)
                h.bias.mul_(0.02)
    models = []
    for _ in range(2):
        m = _build(cls, P, 3, 3.7)
        g = torch.Generator().manual_seed(21)
        pick = torch.randint(0, cent.shape[0] // 2, (P,), generator=g).cuda()       # half of the faces stay empty
        pos = cent[pick] + 0.004 * torch.randn(P, 3, generator=g).cuda()
        pos[: P // 10] = 3.0 * torch.randn(P // 10, 3, generator=g).cuda()           # far away: pruned as invalid
        with torch.no_grad():
            m._xyz.copy_(pos)
        m.gaussian_scale = 1.0
        models.append(m)
    a, b = models
    losses = []
    for m, fn in ((a, cls._reference_anchor_mesh), (b, cls.anchor_mesh)):
        torch.manual_seed(5)
        for p in deform.deform.parameters():
            p.grad = None
        loss = fn(m, verts, faces, deform, deform_back, 0.3, search_radius=0.0005, topn=2, bs=bs, increase_bs=increase_bs)
        loss.backward()
        losses.append((float(loss), {k: p.grad.clone() for k, p in deform.deform.named_parameters() if p.grad is not None}))
    n = a._xyz.shape[0]
    assert n != P and b._xyz.shape[0] == n, (P, n, b._xyz.shape[0])
    assert math.isfinite(losses[0][0]) and abs(losses[0][0] - losses[1][0]) <= 1e-5 * abs(losses[0][0])
    assert losses[0][1].keys() == losses[1][1].keys() and losses[0][1]
    for k in losses[0][1]:
        assert util.rel_l2(losses[1][1][k], losses[0][1][k]) < 1e-3, k
    for x in ATTRS:
        pa, pb = getattr(a, x), getattr(b, x)
        assert pa.shape == pb.shape and isinstance(pb, torch.nn.Parameter) and pb.requires_grad, x
        assert torch.allclose(pa, pb, rtol=1e-5, atol=1e-6), (x, float((pa - pb).abs().max()))
    ga = {g["name"]: g for g in a.optimizer.param_groups}
    gb = {g["name"]: g for g in b.optimizer.param_groups}
    for name, attr in zip(("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation", "normal"), ATTRS):
        assert gb[name]["params"][0] is getattr(b, attr)
        sa, sb = a.optimizer.state[ga[name]["params"][0]], b.optimizer.state[gb[name]["params"][0]]
        assert torch.equal(sa["exp_avg"], sb["exp_avg"]) and torch.equal(sa["exp_avg_sq"], sb["exp_avg_sq"]), name
    for s in ("xyz_gradient_accum", "denom", "max_radii2D"):
        assert getattr(a, s).shape == getattr(b, s).shape


def test_face_geometry_matches_the_trimesh_restatement_cpu():
    import os
    import sys
    sys.path.insert(0, os.path.join(util.ROOT, "tools"))
    import harness_stubs
    harness_stubs.install()
    import anchor
    import trimesh                                    # the functional stand-in of tools/harness_stubs.py, or the real one
    verts, faces = _sphere(12, 20, device="cpu")
    c, nrm, e = anchor.face_geometry(verts, faces)
    mesh = trimesh.Trimesh(vertices=verts.cpu().numpy(), faces=faces.cpu().numpy())
    assert torch.allclose(c.cpu(), torch.tensor(mesh.triangles_center, dtype=torch.float), atol=1e-7)
    assert torch.allclose(nrm.cpu(), torch.tensor(mesh.face_normals, dtype=torch.float), atol=1e-6)
    assert abs(float(e) - float(mesh.edges_unique_length.mean())) < 1e-9


@pytest.mark.parametrize("topn", [2, 3])
def test_select_topn_equals_the_match_mask_formulation_cpu(topn):
    """anchor.select_topn (one stable sort) against the reference's formulation of the same selection
    (gaussian_model_dpsr_dynamic_anchor.py:787-803): an [X, G] match mask, a row-wise running count, the first
    `topn` matches of every row kept, the remaining matches deleted -- evaluated here with plain loops."""
    import anchor
    g = torch.Generator().manual_seed(3 + topn)
    n_faces, G = 40, 600
    face_indices = torch.randint(0, n_faces, (G,), generator=g)
    counts = torch.bincount(face_indices, minlength=n_faces)
    eligible = torch.nonzero(counts >= topn).squeeze(1)
    drawn = eligible[torch.randperm(eligible.shape[0], generator=g)[:7]]          # distinct faces, random order
    to_delete, members = anchor.select_topn(face_indices, drawn, n_faces, topn)
    want_delete = torch.zeros(G, dtype=torch.bool)
    want_members = []
    for f in drawn.tolist():
        idx = [i for i in range(G) if int(face_indices[i]) == f]
        want_members.append(idx[:topn])
        for i in idx[topn:]:
            want_delete[i] = True
    assert torch.equal(to_delete, want_delete)
    assert members.tolist() == want_members
    # Gaussians of faces that were not drawn are untouched
    assert not bool(to_delete[~torch.isin(face_indices, drawn)].any())
    # a drawn face with too few Gaussians is an error, as in the reference
    few = torch.nonzero((counts > 0) & (counts < topn)).squeeze(1)
    if few.numel():
        with pytest.raises(RuntimeError):
            anchor.select_topn(face_indices, few[:1], n_faces, topn)
    # nothing drawn: nothing happens
    td, mm = anchor.select_topn(face_indices, drawn[:0], n_faces, topn)
    assert not bool(td.any()) and tuple(mm.shape) == (0, topn)
