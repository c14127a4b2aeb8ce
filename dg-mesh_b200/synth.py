"""Seeded synthetic scenes and cameras of the shapes BASELINE.json names.

No dataset or checkpoint is available offline, so every benchmark / parity input is
generated here: Gaussians distributed like a trained D-NeRF object (unit-ish ball,
log-normal scales, un-normalised quaternions as `render()` hands them to the
rasterizer) and D-NeRF-style cameras on a ring looking at the origin.  Camera matrices
follow the reference conventions (dgmesh/scene/cameras.py:57-71,
dgmesh/utils/graphics_utils.py:42-76): `world_view_transform` is the world-to-camera
matrix TRANSPOSED (row-vector convention) and `full_proj_transform` = view^T-proj^T product.
"""
import math
from types import SimpleNamespace

import numpy as np
import torch


def look_at_camera(azimuth_deg=30.0, elevation_deg=20.0, radius=4.0, fovx=0.6911, fovy=0.6911, width=800,
                   height=800, znear=0.01, zfar=100.0, fid=0.0, device="cpu"):
    """A camera on a sphere of `radius` looking at the origin (OpenCV/COLMAP axes: +z forward,
    +y down), expressed the way the reference `Camera` exposes it to `render()`."""
    az, el = math.radians(azimuth_deg), math.radians(elevation_deg)
    c = np.array([radius * math.cos(el) * math.sin(az), -radius * math.sin(el), -radius * math.cos(el) * math.cos(az)])
    fwd = -c / np.linalg.norm(c)
    up_hint = np.array([0.0, -1.0, 0.0])
    right = np.cross(up_hint, fwd)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    R_c2w = np.stack([right, down, fwd], axis=1)  # columns: camera axes in world
    w2c = np.eye(4)
    w2c[:3, :3] = R_c2w.T
    w2c[:3, 3] = -R_c2w.T @ c
    world_view = torch.tensor(np.float32(w2c)).transpose(0, 1).contiguous()
    # perspective matrix, same entries as getProjectionMatrix (graphics_utils.py:51-73)
    tan_y, tan_x = math.tan(fovy / 2), math.tan(fovx / 2)
    top, right_ = tan_y * znear, tan_x * znear
    Pm = torch.zeros(4, 4)
    Pm[0, 0] = 2.0 * znear / (2 * right_)
    Pm[1, 1] = 2.0 * znear / (2 * top)
    Pm[3, 2] = 1.0
    Pm[2, 2] = zfar / (zfar - znear)
    Pm[2, 3] = -(zfar * znear) / (zfar - znear)
    proj = Pm.transpose(0, 1)
    full = (world_view.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0)
    center = world_view.inverse()[3, :3]
    return SimpleNamespace(
        FoVx=fovx, FoVy=fovy, image_width=width, image_height=height, znear=znear, zfar=zfar,
        world_view_transform=world_view.to(device), projection_matrix=proj.to(device),
        full_proj_transform=full.contiguous().to(device), camera_center=center.contiguous().to(device),
        fid=torch.tensor([fid], dtype=torch.float32, device=device), K=None)


def gaussian_scene(n=100_000, seed=0, sh_degree=3, scale_median=0.01, device="cpu"):
    """the tensors `render()` passes to the rasterizer."""
    g = torch.Generator().manual_seed(seed)
    xyz = torch.randn(n, 3, generator=g) * 0.5
    norm = xyz.norm(dim=1, keepdim=True).clamp_min(1e-9)
    xyz = torch.where(norm > 1.3, xyz * (1.3 / norm), xyz)
    scaling_raw = math.log(scale_median) + 0.5 * torch.randn(n, 3, generator=g)   # _scaling (log space)
    rot_raw = torch.randn(n, 4, generator=g)
    d_rot = 0.01 * torch.randn(n, 4, generator=g)
    opacity_raw = 1.5 * torch.randn(n, 1, generator=g)
    m = (sh_degree + 1) ** 2
    shs = 0.3 * torch.randn(n, 16, 3, generator=g)
    shs[:, 0, :] += 0.5  # a visible base colour
    normals = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=1)
    out = dict(
        means3D=xyz, scales=torch.exp(scaling_raw),
        rotations=torch.nn.functional.normalize(rot_raw, dim=1) + d_rot,
        opacities=torch.sigmoid(opacity_raw), shs=shs, normals=normals,
        scaling_raw=scaling_raw, rotation_raw=rot_raw, opacity_raw=opacity_raw)
    assert m <= 16
    return {k: v.contiguous().to(device) for k, v in out.items()}


def raster_settings_for(cam, bg, sh_degree=3, scale_modifier=1.0, debug=False, settings_cls=None):
    """GaussianRasterizationSettings for a camera, as render() builds it
    (dgmesh/gaussian_renderer/__init__.py:47-64)."""
    if settings_cls is None:
        from diff_gaussian_rasterization import GaussianRasterizationSettings as settings_cls
    return settings_cls(
        image_height=int(cam.image_height), image_width=int(cam.image_width),
        tanfovx=math.tan(cam.FoVx * 0.5), tanfovy=math.tan(cam.FoVy * 0.5), bg=bg,
        scale_modifier=scale_modifier, viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform,
        sh_degree=sh_degree, campos=cam.camera_center, prefiltered=False, debug=debug)
