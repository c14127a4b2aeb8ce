"""mesh_renderer glue on the GPU: Gaussians on a sphere -> DPSR -> marching cubes -> vertex colours; the
mesh is a closed sphere of the right radius and gradients reach xyz, normals, the density threshold
and the MLP parameters (SURVEY config C3 without the nvdiffrast rasterisation)."""
import importlib
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle.oracle import mesh_topology

pytestmark = pytest.mark.gpu


def test_sphere_of_gaussians_becomes_a_sphere_mesh():
    from diso import DiffMC
    from nvdiffrast_utils.dpsr import DPSR
    tu = importlib.import_module("utils.time_utils")
    renderer = importlib.import_module("utils.renderer")
    torch.manual_seed(0)
    n, G, R = 60_000, 96, 0.6
    d = torch.nn.functional.normalize(torch.randn(n, 3), dim=1).cuda()
    xyz = (d * R).requires_grad_(True)
    normal = d.clone().requires_grad_(True)
    thres = torch.nn.Parameter(torch.tensor(0.0, device="cuda"))
    gauss = SimpleNamespace(get_xyz=xyz, get_normal=normal, gaussian_center=torch.zeros(3, device="cuda"),
                            gaussian_scale=torch.tensor([1.0], device="cuda"), density_thres_param=thres,
                            dpsr=DPSR(res=(G, G, G), sig=2.0), diffmc=DiffMC(dtype=torch.float32).cuda())
    back = SimpleNamespace(net=tu.DeformNetworkNormal(is_blender=True).cuda())
    back.step = lambda x, t: back.net(x, t)
    app = SimpleNamespace(net=tu.AppearanceNetwork(is_blender=True).cuda())
    app.step = lambda x, t: app.net(x, t)
    zeros = torch.zeros(n, 3, device="cuda")
    fid = torch.tensor([0.3], device="cuda")
    verts, faces, color = renderer.mesh_renderer(None, gauss, zeros, zeros, fid, back, app)
    assert verts.dtype == torch.float32 and faces.dtype == torch.int32 and color.shape == (verts.shape[0], 3)
    euler, manifold, oriented = mesh_topology(verts.detach().cpu().numpy(), faces.cpu().numpy())
    assert euler == 2 and manifold and oriented
    r = verts.detach().norm(dim=1)
    assert abs(float(r.mean()) - R) < 0.03 and float(r.std()) < 0.02
    assert float(color.min()) >= 0 and float(color.max()) <= 1
    (verts.sum() + color.sum()).backward()
    for name, p in (("xyz", xyz), ("normal", normal), ("thres", thres),
                    ("appearance", app.net.linear[0].weight), ("deform_back", back.net.gaussian_warp.weight)):
        assert p.grad is not None and torch.isfinite(p.grad).all() and float(p.grad.abs().sum()) > 0, name
