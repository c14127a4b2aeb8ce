"""Adaptive density control: the fused device version (dg-mesh_b200/densify.py, csrc/densify.cu)
against the reference's own GaussianModelDPSRDynamicAnchor.densify_and_prune
(oracle/_ref/dgmesh/scene/gaussian_model_dpsr_dynamic_anchor.py:540-551, imported unmodified) under the same seed:
identical survivors in the identical order, copied tensors bit-equal, computed ones (child positions / scales)
to 1e-6, identical Adam state, and an optimiser that keeps working on the new parameters."""
import argparse
import os
import sys

import numpy as np
import pytest
import torch

import util

REF_TREE = os.path.join(util.ROOT, "oracle", "_ref", "dgmesh")
ATTRS = ("_xyz", "_features_dc", "_features_rest", "_opacity", "_scaling", "_rotation", "_normal")


def _reference_model_class():
    if not os.path.isdir(REF_TREE):
        return None
    sys.path.insert(0, os.path.join(util.ROOT, "tools"))
    import harness_stubs
    harness_stubs.install()
    import launch
    launch.install(REF_TREE)
    import scene
    return scene.GaussianModelDPSRDynamicAnchor


def _build(cls, P, seed, extent):
    g = torch.Generator().manual_seed(seed)
    m = cls(3, 32, 0.0, 3.0)
    r = lambda *s: torch.randn(*s, generator=g)      # noqa: E731
    m._xyz = torch.nn.Parameter(r(P, 3).cuda())
    m._features_dc = torch.nn.Parameter(r(P, 1, 3).cuda())
    m._features_rest = torch.nn.Parameter((0.1 * r(P, 15, 3)).cuda())
    m._opacity = torch.nn.Parameter((r(P, 1) * 2.0 - 3.0).cuda())                 # some below min_opacity = 0.005
    # scales around percent_dense * extent so that clone AND split both happen, a few above 0.1 * extent
    m._scaling = torch.nn.Parameter((torch.log(torch.tensor(0.01 * extent)) + 1.2 * r(P, 3)).cuda())
    m._rotation = torch.nn.Parameter(r(P, 4).cuda())
    m._normal = torch.nn.Parameter(r(P, 3).cuda())
    m.max_radii2D = (torch.rand(P, generator=g) * 40).cuda()
    opt = argparse.Namespace(percent_dense=0.01, position_lr_init=1.6e-4, position_lr_final=1.6e-6,
                             position_lr_delay_mult=0.01, position_lr_max_steps=30000, feature_lr=0.0025,
                             opacity_lr=0.05, scaling_lr=0.001, rotation_lr=0.001)
    m.training_setup(opt)
    for a in ATTRS:                                   # one Adam step so that exp_avg / exp_avg_sq / step exist
        p = getattr(m, a)
        p.grad = (0.01 * torch.randn(p.shape, generator=g)).cuda()
    m.density_thres_param.grad = torch.zeros_like(m.density_thres_param)
    m.optimizer.step()
    m.optimizer.zero_grad(set_to_none=True)
    m.xyz_gradient_accum = (torch.rand(P, 1, generator=g) * 4e-4).cuda()
    m.denom = torch.randint(0, 3, (P, 1), generator=g).float().cuda()      # zeros -> NaN / inf gradients
    return m


@pytest.mark.gpu
@pytest.mark.parametrize("P,size_thr", [(20000, 20), (5000, None), (333, 20)])
def test_fused_densify_and_prune_equals_reference(P, size_thr):
    cls = _reference_model_class()
    if cls is None:
        pytest.skip("oracle/_ref/dgmesh missing")
    import densify
    assert cls.densify_and_prune is densify.densify_and_prune            # the launcher swapped it in
    extent = 3.7
    a, b = _build(cls, P, 3, extent), _build(cls, P, 3, extent)
    for x in ATTRS:
        assert torch.equal(getattr(a, x), getattr(b, x))
    torch.manual_seed(5)
    cls._reference_densify_and_prune(a, 0.0002, 0.005, extent, size_thr)
    torch.manual_seed(5)
    b.densify_and_prune(0.0002, 0.005, extent, size_thr)
    n = a._xyz.shape[0]
    assert n != P and b._xyz.shape[0] == n, (P, n, b._xyz.shape[0])
    for x in ATTRS:
        pa, pb = getattr(a, x), getattr(b, x)
        assert pa.shape == pb.shape and isinstance(pb, torch.nn.Parameter) and pb.requires_grad, x
        if x in ("_xyz", "_scaling"):
            assert torch.allclose(pa, pb, rtol=1e-6, atol=1e-6), (x, float((pa - pb).abs().max()))
        else:
            assert torch.equal(pa, pb), x
    ga = {g["name"]: g for g in a.optimizer.param_groups}
    gb = {g["name"]: g for g in b.optimizer.param_groups}
    for name, attr in zip(("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation", "normal"), ATTRS):
        assert gb[name]["params"][0] is getattr(b, attr)
        sa, sb = a.optimizer.state[ga[name]["params"][0]], b.optimizer.state[gb[name]["params"][0]]
        assert torch.equal(sa["exp_avg"], sb["exp_avg"]) and torch.equal(sa["exp_avg_sq"], sb["exp_avg_sq"]), name
        assert float(sa["step"]) == float(sb["step"])
    assert len(b.optimizer.state) == len(a.optimizer.state)
    for s in ("xyz_gradient_accum", "denom", "max_radii2D"):
        assert getattr(a, s).shape == getattr(b, s).shape and float(getattr(b, s).abs().sum()) == 0.0
    # the optimiser keeps working on the re-seated parameters
    g = torch.Generator().manual_seed(9)
    for m in (a, b):
        g.manual_seed(9)
        for x in ATTRS:
            p = getattr(m, x)
            p.grad = (0.01 * torch.randn(p.shape, generator=g)).cuda()
        m.density_thres_param.grad = torch.zeros_like(m.density_thres_param)
        m.optimizer.step()
    assert torch.allclose(a._xyz, b._xyz, rtol=1e-6, atol=1e-6) and torch.equal(a._opacity, b._opacity)


@pytest.mark.gpu
def test_ply_interchange_with_the_reference_model(tmp_path):
    """SURVEY 8(f)-4 wired to the model: a point cloud written by the reference's own save_ply
    (gaussian_model_dpsr_dynamic_anchor.py:253-289) is read by checkpoint.load_gaussians_ply, and one written by
    checkpoint.save_gaussians_ply is read by the reference's load_ply (:296-362), values bit-equal both ways.
    (`plyfile` itself is not installed: the reference code runs on tools/harness_stubs' minimal stand-in, which
    only serialises the numpy structured arrays the REFERENCE builds -- names, order and layout are its own.)"""
    cls = _reference_model_class()
    if cls is None:
        pytest.skip("oracle/_ref/dgmesh missing")
    import checkpoint
    m = _build(cls, 257, 11, 3.7)
    m.gaussian_center = torch.tensor([0.1, -0.2, 0.3], device="cuda")
    m.gaussian_scale = torch.tensor([1.7], device="cuda")
    m.density_thres_param.data.fill_(0.05)
    # reference -> ours
    p1 = str(tmp_path / "a" / "point_cloud" / "iteration_7" / "point_cloud.ply")
    m.save_ply(p1)
    d = checkpoint.load_gaussians_ply(p1, max_sh_degree=3)
    for k, attr in (("xyz", "_xyz"), ("normal", "_normal"), ("features_dc", "_features_dc"),
                    ("features_rest", "_features_rest"), ("opacity", "_opacity"), ("scaling", "_scaling"),
                    ("rotation", "_rotation")):
        assert np.array_equal(d[k], getattr(m, attr).detach().cpu().numpy()), k
    assert abs(float(d["density_thres"][0]) - 0.05) < 1e-7 and abs(float(d["gaussian_scale"][0]) - 1.7) < 1e-6
    assert np.allclose(d["gaussian_center"].reshape(-1), [0.1, -0.2, 0.3], atol=1e-7)
    # ours -> reference
    p2 = str(tmp_path / "b" / "point_cloud" / "iteration_9" / "point_cloud.ply")
    checkpoint.save_gaussians_ply(p2, m._xyz, m._normal, m._features_dc, m._features_rest, m._opacity, m._scaling,
                                  m._rotation, m.density_thres_param, m.gaussian_center, m.gaussian_scale)
    m2 = cls(3, 32, 0.0, 3.0)
    m2.load_ply(str(tmp_path / "b"), iteration=9)
    for attr in ATTRS:
        assert torch.equal(getattr(m2, attr).detach(), getattr(m, attr).detach()), attr
    assert torch.allclose(m2.gaussian_center.reshape(-1).float().cpu(), m.gaussian_center.cpu())
