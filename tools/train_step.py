#!/usr/bin/env python
"""Full training-step time (the steady-state iteration of dgmesh/train.py:150-311
past dpsr_iter: 4 N-sized deformation MLPs, Gaussian rasterizer, image loss (L1 + SSIM), DPSR, marching cubes,
2 V-sized vertex MLPs, mesh rasterisation (mask + image), mask / mesh-image / Laplacian losses; forward and
backward) for this implementation and, beside it, for the reference's own modules on the same GPU: fp32
PyTorch MLPs, DPSR and loss functions (oracle/_ref/refpy), the stock CUDA rasterizer (oracle/_ref), the
reference's Laplacian arithmetic.  Two third-party packages of the reference are not in its tree and not
installable here -- `diso` (marching cubes) and `nvdiffrast` (mesh rasteriser) -- so BOTH arms use this repo's
kernels for those two stages (which favours the reference arm: its own would be no faster).

    python tools/train_step.py [--gaussians 200000] [--grid 288] [--steps 10]
prints one JSON object with the step time and a per-component breakdown of both arms.
"""
import argparse
import importlib
import json
import math
import os
import sys
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dg-mesh_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402


def build(arm, n, G, dev, azimuth_deg=30.0):
    import synth
    import util
    torch.manual_seed(0)
    if arm == "ours":
        import diff_gaussian_rasterization as dgr
        tu = importlib.import_module("utils.time_utils")
        from nvdiffrast_utils.dpsr import DPSR
    else:
        ref = util.load_reference_pymodules()
        dgr = util.load_reference_rasterizer()
        if ref is None or dgr is None:
            return None
        tu, DPSR = ref.time_utils, ref.dpsr.DPSR
    from diso import DiffMC
    renderer = importlib.import_module("utils.renderer")       # this repo's mesh branch glue in both arms
    if arm == "ours":
        loss_utils = importlib.import_module("utils.loss_utils")
        laplacian = importlib.import_module("nvdiffrast_utils.regularizer").laplace_regularizer_const
    else:
        import importlib.util as ilu
        spec = ilu.spec_from_file_location("refpy.loss_utils", os.path.join(util.REF_DIR, "refpy", "loss_utils.py"))
        loss_utils = ilu.module_from_spec(spec)
        spec.loader.exec_module(loss_utils)
        laplacian = None
    sc = synth.gaussian_scene(n=n, seed=0, device=dev)
    g = torch.Generator().manual_seed(3)
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=1)
    xyz = (d * 0.6 + 0.01 * torch.randn(n, 3, generator=g)).to(dev)      # a thick spherical shell: clean mesh
    P = SimpleNamespace(
        xyz=xyz.requires_grad_(True), normal=d.to(dev).clone().requires_grad_(True),
        opacities=sc["opacities"].clone().requires_grad_(True), scales=sc["scales"].clone().requires_grad_(True),
        rotations=sc["rotations"].clone().requires_grad_(True), shs=sc["shs"].clone().requires_grad_(True),
        thres=torch.zeros((), device=dev, requires_grad=True))
    nets = SimpleNamespace(
        deform=tu.DeformNetworkNormal(is_blender=True).to(dev), deform_normal=tu.DeformNetworkNormalSep(is_blender=True).to(dev),
        deform_back=tu.DeformNetworkNormal(is_blender=True).to(dev),
        deform_back_normal=tu.DeformNetworkNormalSep(is_blender=True).to(dev),
        appearance=tu.AppearanceNetwork(is_blender=True).to(dev))
    dpsr = DPSR(res=(G, G, G), sig=3.0)
    if arm != "ours":
        dpsr = dpsr.to(dev)
    gauss = SimpleNamespace(gaussian_center=torch.zeros(3, device=dev), gaussian_scale=torch.tensor([1.2 * 1.3], device=dev),
                            dpsr=dpsr, diffmc=DiffMC(dtype=torch.float32).to(dev))
    cam = synth.look_at_camera(azimuth_deg=azimuth_deg, elevation_deg=20.0, radius=4.0, width=800, height=800,
                               device=dev)
    bg = torch.ones(3, device=dev)
    gt = torch.rand(3, 800, 800, device=dev, generator=torch.Generator(device=dev).manual_seed(5))
    gt_mask = (torch.rand(800, 800, 1, device=dev, generator=torch.Generator(device=dev).manual_seed(6)) > 0.5).float()
    # what scene/cameras.py keeps for the mesh rasteriser: the blender camera-to-world matrix
    w2c = cam.world_view_transform.t()
    cam.orig_transform = (torch.inverse(w2c) @ torch.diag(torch.tensor([1.0, -1.0, -1.0, 1.0], device=dev))).cpu().numpy()
    cam.K = None
    params = [P.xyz, P.normal, P.opacities, P.scales, P.rotations, P.shs, P.thres] + \
        [q for m in vars(nets).values() for q in m.parameters()]
    return SimpleNamespace(arm=arm, dgr=dgr, renderer=renderer, P=P, nets=nets, gauss=gauss, cam=cam, bg=bg, gt=gt,
                           gt_mask=gt_mask, loss_utils=loss_utils, laplacian=laplacian, params=params, n=n, synth=synth)


def mlp_part(S, t):
    P, N = S.P, S.nets
    d_xyz, d_rot, d_scale, d_normal = N.deform(P.xyz, t)
    d_normal = d_normal + N.deform_normal(P.xyz, t)
    warped = P.xyz + d_xyz
    b_xyz, _, _, b_normal = N.deform_back(warped.detach(), t)
    b_normal = b_normal + N.deform_back_normal(warped.detach(), t)
    cycle = (d_xyz + b_xyz).abs().mean() + (d_normal + b_normal).abs().mean()
    return d_xyz, d_rot, d_scale, d_normal, cycle


def raster_part(S, d_xyz, d_rot, d_scale):
    P, cam = S.P, S.cam
    rs = S.dgr.GaussianRasterizationSettings(
        image_height=800, image_width=800, tanfovx=math.tan(cam.FoVx * 0.5), tanfovy=math.tan(cam.FoVy * 0.5),
        bg=S.bg, scale_modifier=1.0, viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform,
        sh_degree=3, campos=cam.camera_center, prefiltered=False, debug=False)
    m2d = torch.zeros_like(P.xyz, requires_grad=True)
    img, radii = S.dgr.GaussianRasterizer(rs)(means3D=P.xyz + d_xyz, means2D=m2d, opacities=P.opacities, shs=P.shs,
                                              scales=P.scales + d_scale, rotations=P.rotations + d_rot)
    return _img_loss(S, img, S.gt)


LAMBDA_DSSIM = 0.2


def _img_loss(S, img, gt):
    """(1 - l) L1 + l (1 - SSIM), dgmesh/train.py:308-311"""
    if S.arm == "ours":
        return S.loss_utils.image_loss(img, gt, LAMBDA_DSSIM)
    lu = S.loss_utils
    return (1.0 - LAMBDA_DSSIM) * lu.l1_loss(img, gt) + LAMBDA_DSSIM * (1.0 - lu.ssim(img, gt))


def _laplacian_ref(v_pos, t_pos_idx):
    """nvdiffrast_utils/regularizer.py:40-59, the reference's arithmetic (its module cannot be imported without
    nvdiffrast)"""
    term = torch.zeros_like(v_pos)
    norm = torch.zeros_like(v_pos[..., 0:1])
    v0, v1, v2 = v_pos[t_pos_idx[:, 0], :], v_pos[t_pos_idx[:, 1], :], v_pos[t_pos_idx[:, 2], :]
    term.scatter_add_(0, t_pos_idx[:, 0:1].repeat(1, 3), (v1 - v0) + (v2 - v0))
    term.scatter_add_(0, t_pos_idx[:, 1:2].repeat(1, 3), (v0 - v1) + (v2 - v1))
    term.scatter_add_(0, t_pos_idx[:, 2:3].repeat(1, 3), (v0 - v2) + (v1 - v2))
    two = torch.ones_like(v0) * 2.0
    norm.scatter_add_(0, t_pos_idx[:, 0:1], two)
    norm.scatter_add_(0, t_pos_idx[:, 1:2], two)
    norm.scatter_add_(0, t_pos_idx[:, 2:3], two)
    term = term / torch.clamp(norm, min=1.0)
    return torch.mean(term ** 2)


def mesh_part(S, d_xyz, d_normal, t1):
    """train.py:238-283: mesh_renderer -> mask loss, mesh image loss, Laplacian"""
    g = S.gauss
    g.get_xyz, g.get_normal, g.density_thres_param = S.P.xyz, S.P.normal, S.P.thres
    back = SimpleNamespace(step=lambda x, t: S.nets.deform_back(x, t))
    app = SimpleNamespace(step=lambda x, t: S.nets.appearance(x, t))
    mask, mesh_image, verts, faces, _ = S.renderer.mesh_renderer(None, g, d_xyz, d_normal, t1[0], back, app, False,
                                                                 True, S.cam)
    mask_loss = (mask - S.gt_mask).abs().mean() * 100
    mesh_img_loss = _img_loss(S, mesh_image, S.gt)
    lap = (S.laplacian(verts, faces.long()) if S.arm == "ours" else _laplacian_ref(verts, faces.long())) * 1000 * 0.5
    return mask_loss + mesh_img_loss + lap, verts.shape[0]


def full_step(S, t_value=0.37, clear_grads=True):
    if clear_grads:
        for q in S.params:
            q.grad = None
    t1 = torch.full((1, 1), float(t_value), device=S.bg.device)
    t = t1.expand(S.n, 1)
    d_xyz, d_rot, d_scale, d_normal, cycle = mlp_part(S, t)
    loss = raster_part(S, d_xyz, d_rot, d_scale) + cycle
    lm, V = mesh_part(S, d_xyz, d_normal, t1)
    (loss + lm).backward()
    return V


def timeit(fn, steps, warmup=3):
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


def components(S, steps):
    dev = S.bg.device
    t1 = torch.full((1, 1), 0.37, device=dev)
    t = t1.expand(S.n, 1)

    def f_mlp():
        d_xyz, d_rot, d_scale, d_normal, cycle = mlp_part(S, t)
        (cycle + d_rot.sum() + d_scale.sum()).backward()

    def f_raster():
        z = torch.zeros_like(S.P.xyz)
        raster_part(S, z, torch.zeros_like(S.P.rotations), z).backward()

    def f_mesh():
        z = torch.zeros_like(S.P.xyz)
        mesh_part(S, z, z, t1)[0].backward()

    return {"mlp_4xN_ms": timeit(f_mlp, steps), "raster_and_image_loss_ms": timeit(f_raster, steps),
            "mesh_branch_ms": timeit(f_mesh, steps)}


def measure(n=200_000, G=288, steps=10, arms=("ours", "reference")):
    dev = torch.device("cuda")
    out = {"config": f"C3: {n} Gaussians, 800x800, grid {G}: 4 N-MLPs + Gaussian raster + L1/SSIM + DPSR + MC + "
                     "2 V-MLPs + mesh raster (mask, image) + mask/L1/SSIM/Laplacian losses, fwd+bwd",
           "note": "marching cubes and the mesh rasteriser are this repo's kernels in BOTH arms (diso / nvdiffrast are "
                   "third-party, absent from the reference tree); the reference arm's MLPs, DPSR and losses are its fp32 "
                   "PyTorch modules, its Gaussian rasterizer the stock CUDA build"}
    for arm in arms:
        S = build(arm, n, G, dev)
        if S is None:
            out[arm] = {"unavailable": "oracle/_ref not built"}
            continue
        V = full_step(S)
        ms = timeit(lambda: full_step(S), steps)
        out[arm] = {"train_step_ms": ms, "mesh_vertices": V, **components(S, steps)}
        del S
        torch.cuda.empty_cache()
    if "train_step_ms" in out.get("ours", {}) and "train_step_ms" in out.get("reference", {}):
        out["speedup"] = out["reference"]["train_step_ms"] / out["ours"]["train_step_ms"]
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gaussians", type=int, default=200_000)
    ap.add_argument("--grid", type=int, default=288)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--profile", action="store_true",
                    help="two full steps of this implementation only, bracketed by cudaProfilerStart/Stop around the "
                         "second one (for `ncu --profile-from-start off --metrics gpu__time_duration.sum`)")
    a = ap.parse_args()
    if a.profile:
        S = build("ours", a.gaussians, a.grid, torch.device("cuda"))
        full_step(S)
        full_step(S)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        full_step(S)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        sys.exit(0)
    print(json.dumps(measure(a.gaussians, a.grid, a.steps)))
