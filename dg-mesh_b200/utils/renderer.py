"""Drop-in `utils.renderer.mesh_renderer` (reference: dgmesh/utils/renderer.py:124-233):
Gaussians -> DPSR indicator grid -> marching cubes -> per-vertex colour, i.e. the mesh-extraction part
of the hot path.  Same signature, same return values.

What changes underneath: DPSR and marching cubes are the sm_100a kernels of this package; when
`gaussians.dpsr` is this package's DPSR, the sign fix + threshold that follow the solve
(reference :163-168) are fused into it, which removes the host read of psr[0,0,0,0].
The mesh RASTERISATION at the end (`render_mask` / `render_mesh`, reference :33-121) is nvdiffrast, a
third-party OpenGL/CUDA rasteriser outside this round's scope: with
`viewpoint_cam is None` (dynamic-mesh export, train.py:403,447) nothing else is needed; with a
camera, nvdiffrast is imported lazily and used exactly as the reference does."""
import torch

SMALL_NUMBER = 1e-6


def _nvdiffrast():
    try:
        import nvdiffrast.torch as dr
        from nvdiffrast_utils import util
    except Exception as e:  # pragma: no cover - depends on the deployment
        raise RuntimeError("mesh_renderer(viewpoint_cam=...) needs nvdiffrast (mesh rasterisation is outside the "
                           "B200 hot path; see INTEGRATION.md)") from e
    return dr, util


def render_mask(glctx, mesh_v_pos, mesh_t_pos_idx, pose, K, resolution=[800, 800]):
    """reference :33-66"""
    dr, util = _nvdiffrast()
    proj = util.K_to_projection(K, resolution[0], resolution[1])
    v_pos_clip = util.transform_pos(proj @ pose, mesh_v_pos)
    rast_out, _ = dr.rasterize(glctx, v_pos_clip, mesh_t_pos_idx, resolution=resolution)
    vtx_color = torch.ones(mesh_v_pos.shape, dtype=torch.float, device=v_pos_clip.device)
    color, _ = dr.interpolate(vtx_color[None, ...], rast_out, mesh_t_pos_idx)
    color = dr.antialias(color, rast_out, v_pos_clip, mesh_t_pos_idx)
    return torch.flip(color[0, :, :], dims=[0])


def render_mesh(glctx, mesh_v_pos, mesh_t_pos_idx, vtx_color, pose, K, resolution=[800, 800], whitebackground=False):
    """reference :69-121"""
    dr, util = _nvdiffrast()
    proj = util.K_to_projection(K, resolution[0], resolution[1])
    v_pos_clip = util.transform_pos(proj @ pose, mesh_v_pos)
    rast_out, _ = dr.rasterize(glctx, v_pos_clip, mesh_t_pos_idx, resolution=resolution)
    output, _ = dr.interpolate(vtx_color[None, ...], rast_out, mesh_t_pos_idx)
    output = dr.antialias(output, rast_out, v_pos_clip, mesh_t_pos_idx)
    output = torch.flip(output, dims=[1])[0]
    ones = torch.ones(mesh_v_pos.shape, dtype=torch.float, device=v_pos_clip.device)
    color, _ = dr.interpolate(ones[None, ...], rast_out, mesh_t_pos_idx)
    color = dr.antialias(color, rast_out, v_pos_clip, mesh_t_pos_idx)
    mask = torch.flip(color[0, :, :], dims=[0])
    output[~mask.bool()] = 1 if whitebackground else 0
    return torch.clamp(output, 0.0, 1.0).permute(2, 0, 1)


def extract_mesh(gaussians, d_xyz, d_normal, freeze_pos=False):
    """Gaussians (+ deltas) -> (verts [V,3] float32 in world space, faces [F,3] int32): reference :150-175."""
    if freeze_pos:
        dpsr_points = gaussians.get_xyz.detach() + d_xyz.detach()
    else:
        dpsr_points = gaussians.get_xyz + d_xyz
    dpsr_points = (dpsr_points - gaussians.gaussian_center) / gaussians.gaussian_scale  # [-1, 1]
    dpsr_points = dpsr_points / 2.0 + 0.5                                                 # [0, 1]
    dpsr_points = torch.clamp(dpsr_points, SMALL_NUMBER, 1 - SMALL_NUMBER)
    normals = gaussians.get_normal + d_normal
    if hasattr(gaussians.dpsr, "forward_signed"):
        psr = gaussians.dpsr.forward_signed(dpsr_points.unsqueeze(0), normals.unsqueeze(0),
                                            gaussians.density_thres_param)
    else:  # a reference DPSR module: same arithmetic, with its host synchronisation
        psr = gaussians.dpsr(dpsr_points.unsqueeze(0), normals.unsqueeze(0))
        sign = psr[0, 0, 0, 0].detach()  # Sign for Diso is opposite to dpsr
        psr = psr * (-1 if sign < 0 else 1)
        psr = (psr - gaussians.density_thres_param).squeeze(0)
    verts, faces = gaussians.diffmc(psr, deform=None, isovalue=0.0)
    verts = verts * 2.0 - 1.0  # [-1, 1]
    verts = verts * gaussians.gaussian_scale + gaussians.gaussian_center
    return verts.to(torch.float32), faces.to(torch.int32)


def mesh_renderer(glctx, gaussians, d_xyz, d_normal, fid, deform_back, appearance, freeze_pos=False,
                  whitebackground=False, viewpoint_cam=None):
    """Gaussian mesh renderer (reference :124-233)."""
    verts, faces = extract_mesh(gaussians, d_xyz, d_normal, freeze_pos)
    # Deform mesh vertices back to the canonical mesh and query vertex colour
    N = verts.shape[0]
    time_input = fid.unsqueeze(0).expand(N, -1)
    mesh_deform_back_dxyz, _, _, _ = deform_back.step(verts.detach(), time_input)
    mesh_canonical_xyz = verts + mesh_deform_back_dxyz
    vtx_color = appearance.step(mesh_canonical_xyz, time_input)
    if viewpoint_cam is None:
        return verts, faces, vtx_color
    _, util = _nvdiffrast()
    from utils.graphics_utils import fov2focal
    if viewpoint_cam.K is not None:
        K = torch.tensor(viewpoint_cam.K).float().to("cuda")
    else:
        focalx = fov2focal(viewpoint_cam.FoVx, viewpoint_cam.image_width)
        focaly = fov2focal(viewpoint_cam.FoVy, viewpoint_cam.image_height)
        K = torch.tensor([[focalx, 0, viewpoint_cam.image_width / 2], [0, focaly, viewpoint_cam.image_height / 2],
                          [0, 0, 1]]).float().to("cuda")
    c2w_blender = torch.tensor(viewpoint_cam.orig_transform).cuda().float()  # blender/OpenGL camera
    c2w_opencv = c2w_blender @ util.blender2opencv
    pose = util.opencv2blender @ torch.inverse(c2w_opencv)
    res = [viewpoint_cam.image_height, viewpoint_cam.image_width]
    mask = render_mask(glctx, verts, faces, pose, K, resolution=res)[..., [0]]
    mesh_image = render_mesh(glctx, verts, faces, vtx_color, pose, K, resolution=res, whitebackground=whitebackground)
    return mask, mesh_image, verts, faces, vtx_color


def __getattr__(name):
    """Names this drop-in does not define (`mesh_shape_renderer`, `pointcloud_renderer`: PyTorch3D /
    matplotlib visualisation helpers used by render_test.py / render_trajectory.py, outside the hot
    path) resolve to the reference's own `utils/renderer.py`, found through the merged `utils` package
    path that launch.install() sets up (this directory first, the reference's second)."""
    import importlib.util
    import os
    import sys
    pkg = sys.modules.get("utils")
    here = os.path.dirname(os.path.abspath(__file__))
    for d in list(getattr(pkg, "__path__", [])):
        cand = os.path.join(d, "renderer.py")
        if os.path.abspath(d) != here and os.path.exists(cand):
            mod = sys.modules.get("_reference_utils_renderer")
            if mod is None:
                spec = importlib.util.spec_from_file_location("_reference_utils_renderer", cand)
                mod = importlib.util.module_from_spec(spec)
                sys.modules["_reference_utils_renderer"] = mod
                spec.loader.exec_module(mod)
            if hasattr(mod, name):
                return getattr(mod, name)
    raise AttributeError(f"module 'utils.renderer' has no attribute {name!r}")
