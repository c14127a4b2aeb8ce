"""Mesh rasterisation (rasterize / interpolate / antialias replacing nvdiffrast in
dgmesh/utils/renderer.py:33-121).  nvdiffrast is not available anywhere in this environment, so parity is pinned
(a) to an independent float64 PyTorch restatement of the same contract (oracle/meshrast_oracle.py; gradients by
autograd) and (b) to closed-form geometry: exact fractional coverage along axis-aligned silhouettes, the area of a
rasterised sphere, conservation of the mask sum under translation, tiny-step differences of the restatement."""
import math

import numpy as np
import pytest
import torch

import util
from oracle import meshrast_oracle as orc


def icosphere(sub=2, radius=1.0):
    t = (1 + 5 ** 0.5) / 2
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
         (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10),
         (8, 6, 7), (9, 8, 1)]
    v = [np.array(p, float) / np.linalg.norm(p) for p in v]
    for _ in range(sub):
        cache, nf = {}, []

        def mid(a, b):
            k = (min(a, b), max(a, b))
            if k not in cache:
                m = v[a] + v[b]
                v.append(m / np.linalg.norm(m))
                cache[k] = len(v) - 1
            return cache[k]
        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    return torch.tensor(np.array(v) * radius, dtype=torch.float32), torch.tensor(f, dtype=torch.int32)


def clip_from_world(verts, W, H, fov=0.6911, dist=4.0):
    """A pinhole camera on the -z axis looking at the origin: clip = P * V * (x, y, z, 1)."""
    f = 1.0 / math.tan(fov / 2)
    n, fa = 0.1, 20.0
    P = torch.tensor([[f * H / W, 0, 0, 0], [0, f, 0, 0], [0, 0, -(fa + n) / (fa - n), -2 * fa * n / (fa - n)],
                      [0, 0, -1, 0]], dtype=verts.dtype, device=verts.device)
    Vm = torch.eye(4, dtype=verts.dtype, device=verts.device)
    Vm[2, 3] = -dist
    pw = torch.cat([verts, torch.ones_like(verts[:, :1])], 1)
    return pw @ (P @ Vm).t()


def test_oracle_coverage_is_exact_along_axis_aligned_silhouettes_cpu():
    """The contract the kernels are held to: a pixel cut by a vertical / horizontal silhouette gets exactly its
    covered fraction (no vertex position ever needs the edge to pass through a pixel centre)."""
    W, H = 12, 10
    x0, x1, y0, y1 = 2.3, 8.8, 1.6, 7.25
    px = torch.tensor([[x0, y0], [x1, y0], [x1, y1], [x0, y1]], dtype=torch.float64)
    pos = torch.stack([px[:, 0] / W * 2 - 1, px[:, 1] / H * 2 - 1, torch.zeros(4, dtype=torch.float64),
                       torch.ones(4, dtype=torch.float64)], -1)
    tri = torch.tensor([[0, 1, 2], [0, 2, 3]])
    ids = orc.rasterize_ids(pos, tri, H, W)
    rast = orc.rast_from_ids(pos, tri, ids)
    color = orc.interpolate(torch.ones(4, 1, dtype=torch.float64), rast, tri)
    out = orc.antialias(color, rast, pos, tri, orc.edge_opposites(tri, 4))[..., 0]
    assert int((ids >= 0).sum()) == 7 * 5                     # centres 2.5..8.5 x 2.5..6.5 ... rows 1.5 is outside
    assert abs(float(out[4, 2]) - 0.7) < 1e-12                 # left edge at 2.3 cuts pixel 2: covered [2.3, 3)
    assert abs(float(out[4, 8]) - 0.8) < 1e-12                 # right edge at 8.8: pixel 8 covered [8, 8.8)
    assert abs(float(out[1, 5]) - 0.4) < 1e-12                 # bottom edge at 1.6: pixel row 1 covered [1.6, 2)
    assert abs(float(out[7, 5]) - 0.25) < 1e-12                # top edge at 7.25: row 7 covered [7, 7.25)
    assert float(out[4, 5]) == 1.0 and float(out[0, 5]) == 0.0
    # the diagonal shared by the two triangles is not a silhouette: nothing changes along it
    assert torch.equal(out[3:6, 4:7], torch.ones(3, 3, dtype=torch.float64))


def test_oracle_gradient_is_the_derivative_of_its_forward_cpu():
    """The float64 restatement's autograd gradient (what the CUDA backward is held to) equals central differences
    of its own forward with a step far below the distance to the next pixel-centre crossing."""
    W, H = 28, 24
    verts, tri = icosphere(1, 1.0)
    tri = tri.long()
    opp = orc.edge_opposites(tri, verts.shape[0])
    g = torch.Generator().manual_seed(0)
    attr = torch.rand(verts.shape[0], 3, generator=g, dtype=torch.float64)
    wimg = torch.randn(H, W, 3, generator=g, dtype=torch.float64)

    def f(shift):
        v = verts.double() + shift
        pos = clip_from_world(v, W, H)
        ids = orc.rasterize_ids(pos.detach(), tri, H, W)
        rast = orc.rast_from_ids(pos, tri, ids)
        col = orc.interpolate(attr, rast, tri)
        return (orc.antialias(col, rast, pos, tri, opp) * wimg).sum()

    sh = torch.tensor([0.013, -0.007, 0.02], dtype=torch.float64, requires_grad=True)
    f(sh).backward()
    eps = 1e-7
    for k in range(3):
        e = torch.zeros(3, dtype=torch.float64)
        e[k] = eps
        fd = float((f(sh.detach() + e) - f(sh.detach() - e)) / (2 * eps))
        assert abs(fd - float(sh.grad[k])) < 1e-4 * max(1.0, abs(fd)), (k, fd, float(sh.grad[k]))


def _random_scene(seed, W, H):
    g = torch.Generator().manual_seed(seed)
    vs, fs = icosphere(1, 0.9)
    extra = torch.randn(36, 3, generator=g) * 0.9                      # 12 loose triangles cutting through
    verts = torch.cat([vs, extra])
    tri = torch.cat([fs, (torch.arange(36).view(12, 3) + vs.shape[0]).int()])
    verts = verts + 0.01 * torch.randn(verts.shape, generator=g)
    attr = torch.rand(verts.shape[0], 3, generator=g)
    return verts, tri, attr


@pytest.mark.gpu
@pytest.mark.parametrize("seed,W,H", [(0, 64, 48), (1, 40, 56)])
def test_cuda_rasterizer_matches_float64_restatement(seed, W, H):
    import meshrast as dr
    verts, tri, attr = _random_scene(seed, W, H)
    g = torch.Generator().manual_seed(100 + seed)
    w_img = torch.randn(H, W, 3, generator=g)
    # ---- oracle (float64, autograd)
    vo = verts.double().requires_grad_(True)
    ao = attr.double().requires_grad_(True)
    pos_o = clip_from_world(vo, W, H)
    ids = orc.rasterize_ids(pos_o.detach(), tri, H, W)
    rast_o = orc.rast_from_ids(pos_o, tri.long(), ids)
    col_o = orc.interpolate(ao, rast_o, tri.long())
    opp_o = orc.edge_opposites(tri, verts.shape[0])
    out_o = orc.antialias(col_o, rast_o, pos_o, tri.long(), opp_o)
    (out_o * w_img.double()).sum().backward()
    # ---- CUDA
    vc = verts.cuda().requires_grad_(True)
    ac = attr.cuda().requires_grad_(True)
    pos_c = clip_from_world(vc, W, H)
    rast_c, _ = dr.rasterize(None, pos_c[None], tri.cuda(), resolution=[H, W])
    col_c, _ = dr.interpolate(ac[None], rast_c, tri.cuda())
    out_c = dr.antialias(col_c, rast_c, pos_c[None], tri.cuda())
    (out_c[0] * w_img.cuda()).sum().backward()
    assert torch.equal(dr.edge_opposites(tri.cuda(), verts.shape[0]).cpu().long(), opp_o)
    ids_c = rast_c[0, ..., 3].long().cpu() - 1
    assert float((ids_c == ids).float().mean()) == 1.0, "triangle ids"
    assert util.rel_err(rast_c[0, ..., :3], rast_o[..., :3]) < 1e-4
    assert util.rel_err(col_c[0], col_o) < 1e-4
    assert util.rel_err(out_c[0], out_o) < 1e-4
    assert float((out_c[0].cpu() - col_c[0].cpu()).abs().sum()) > 1.0       # the antialiasing did something
    assert util.rel_err(ac.grad, ao.grad) < 1e-4, "d/d attr"
    assert util.rel_err(vc.grad, vo.grad) < 2e-3, "d/d vertex positions (fp32 atomics vs float64 autograd)"


@pytest.mark.gpu
def test_sphere_mask_area_and_translation_invariance():
    import meshrast as dr
    W = H = 256
    verts, tri = icosphere(3, 1.0)
    verts, tri = verts.cuda(), tri.cuda()

    def mask_of(shift):
        v = verts + shift
        pos = clip_from_world(v, W, H)
        rast, _ = dr.rasterize(None, pos[None], tri, resolution=[H, W])
        ones, _ = dr.interpolate(torch.ones_like(v)[None], rast, tri)
        return dr.antialias(ones, rast, pos[None], tri)[0, ..., 0]

    # area of the silhouette of a unit sphere at distance 4: a disc of radius f * tan(asin(1/4)) in NDC
    f = 1.0 / math.tan(0.6911 / 2)
    r_pix = f * math.tan(math.asin(1 / 4.0)) * H / 2
    m = mask_of(torch.zeros(3, device="cuda"))
    # the icosphere is INSCRIBED in the sphere: its silhouette is a polygon slightly inside the disc
    # (edge angle ~0.16 rad at 3 subdivisions -> ~0.5 % less area), never outside it
    assert -1.2e-2 < float(m.sum()) / (math.pi * r_pix ** 2) - 1 < 1e-3
    assert float(m.max()) <= 1.0 + 1e-6 and float(m.min()) >= 0.0
    # sub-pixel translations parallel to the image plane keep the covered area (antialiased, not aliased)
    sums = [float(mask_of(torch.tensor([dx, 0.0, 0.0], device="cuda")).sum()) for dx in (0.0, 0.004, 0.009, 0.013)]
    assert (max(sums) - min(sums)) / sums[0] < 2e-3, sums
    hard = [float((mask_of(torch.tensor([dx, 0.0, 0.0], device="cuda")) > 0.5).float().sum()) for dx in (0.0, 0.009)]
    assert hard[0] > 0
    # (finite differences of the whole image are not a usable check: the antialiased image is only piecewise
    # smooth -- it jumps whenever a pixel centre crosses an edge, exactly as nvdiffrast's does -- so the gradient
    # is checked against float64 autograd of the restatement, and the restatement against tiny-step differences
    # on the CPU: test_oracle_gradient_is_the_derivative_of_its_forward_cpu)


@pytest.mark.gpu
def test_renderer_mask_and_mesh_image_through_the_reference_call_sequence():
    """utils.renderer.render_mask / render_mesh (reference signatures) on a marching-cubes-sized mesh at 800x800:
    shapes, value ranges, background handling and gradients to vertices and vertex colours."""
    import importlib
    rend = importlib.import_module("utils.renderer")
    verts, tri = icosphere(5, 0.8)                                    # 10 242 vertices, 20 480 faces
    verts = verts.cuda().requires_grad_(True)
    col = torch.rand(verts.shape[0], 3, device="cuda", requires_grad=True)
    H = W = 800
    focal = W / (2 * math.tan(0.6911 / 2))
    K = torch.tensor([[focal, 0, W / 2], [0, focal, H / 2], [0, 0, 1.0]], device="cuda")
    c2w = torch.eye(4, device="cuda")
    c2w[2, 3] = 4.0                                                   # blender camera at +z looking down -z
    pose = torch.inverse(c2w)
    mask = rend.render_mask(None, verts, tri.cuda(), pose, K, resolution=[H, W])
    img = rend.render_mesh(None, verts, tri.cuda(), col, pose, K, resolution=[H, W], whitebackground=True)
    assert mask.shape == (H, W, 3) and img.shape == (3, H, W)
    assert 0.0 <= float(mask.min()) and float(mask.max()) <= 1.0 + 1e-6
    r_pix = focal * math.tan(math.asin(0.8 / 4.0))
    assert -5e-3 < float(mask[..., 0].sum()) / (math.pi * r_pix ** 2) - 1 < 1e-3      # inscribed polyhedron (5 subdivisions)
    assert torch.all(img[:, 0, 0] == 1.0) and float(img[:, H // 2, W // 2].min()) >= 0.0
    gt = torch.zeros(H, W, 1, device="cuda")
    ((mask[..., :1] - gt).abs().mean() * 100 + (img - 0.5).abs().mean()).backward()
    assert float(verts.grad.abs().sum()) > 0 and float(col.grad.abs().sum()) > 0
    assert torch.isfinite(verts.grad).all() and torch.isfinite(col.grad).all()
