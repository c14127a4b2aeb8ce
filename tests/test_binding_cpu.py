"""Host-side logic of the drop-in `diff_gaussian_rasterization` exercised WITHOUT a GPU: the C
library is replaced by a recording stub (same symbol table, argument counts checked against the
ctypes signatures), tensors live on the CPU.  Catches glue errors (argument order, shapes of the
returned gradients, capacity policy) before a GPU run."""
import math

import pytest
import torch

import _dgm_lib
import diff_gaussian_rasterization as dgr
import synth


class StubLib:
    def __init__(self):
        self.calls = []
        self.R = 1000

    def __getattr__(self, name):
        if name not in _dgm_lib.SIGNATURES:
            raise AttributeError(name)
        nargs = len(_dgm_lib.SIGNATURES[name][1])

        def fn(*a):
            assert len(a) == nargs, f"{name}: {len(a)} args, header declares {nargs}"
            self.calls.append((name, a))
            if name == "dgr_workspace_sizes":
                a[4]._obj.value, a[5]._obj.value, a[6]._obj.value = 4096, 60 * a[3] + 128, 8192
            return 0
        return fn


@pytest.fixture()
def stub(monkeypatch):
    st = StubLib()
    monkeypatch.setattr(_dgm_lib, "lib", lambda: st)
    monkeypatch.setattr(_dgm_lib, "stream_ptr", lambda: 0)
    monkeypatch.setattr(dgr, "_f32c", lambda t, name: None if (t is None or t.numel() == 0) else t.contiguous())
    # early notification: pretend R instances per frame, depth range [2, 6]; a frame overflows when the
    # capacity it was enqueued with is below R (as the tile scan reports it)
    import struct

    class FakeNotify:
        host_ptr, event = 0x1000, 0x2000

        def __init__(self):
            self.words = [0] * (64 * 8)

        def wait(self):
            last = [c for c in st.calls if c[0] in ("dgr_forward", "dgr_forward_batch")][-1]
            cap = last[1][26] if last[0] == "dgr_forward" else last[1][28]
            for f in range(64):
                self.words[8 * f:8 * f + 5] = [st.R, int(st.R > cap), 0, struct.unpack("i", struct.pack("f", 2.0))[0],
                                               struct.unpack("i", struct.pack("f", 6.0))[0]]

    fake = FakeNotify()
    monkeypatch.setattr(dgr._Notify, "get", classmethod(lambda cls, dev: fake))
    monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: False)
    dgr._Sizing.hint.clear()
    dgr._ws_sizes.clear()
    yield st


def _byref_shim(monkeypatch):
    # ctypes passes c_size_t by reference; the stub reads `_obj`
    pass


def test_forward_backward_glue(stub, monkeypatch):
    import ctypes
    # emulate ctypes by-reference semantics for the size query
    orig = dgr._sizes

    def sizes(P, W, H, R_cap):
        return (4096, 60 * R_cap + 128, 8192, 4096, (60 * R_cap + 128 + 511) // 512 * 512, 8192)
    monkeypatch.setattr(dgr, "_sizes", sizes)
    sc = synth.gaussian_scene(n=64, seed=0)
    cam = synth.look_at_camera(width=48, height=32)
    rs = synth.raster_settings_for(cam, torch.ones(3), settings_cls=dgr.GaussianRasterizationSettings)
    leaves = {k: sc[k].clone().requires_grad_(True) for k in ("means3D", "opacities", "scales", "rotations", "shs")}
    m2d = torch.zeros_like(leaves["means3D"], requires_grad=True)
    color, radii = dgr.GaussianRasterizer(rs)(means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"],
                                              shs=leaves["shs"], scales=leaves["scales"],
                                              rotations=leaves["rotations"])
    assert color.shape == (3, 32, 48) and radii.shape == (64,) and radii.dtype == torch.int32
    color.backward(torch.ones_like(color))
    names = [c[0] for c in stub.calls]
    assert names.count("dgr_forward") == 1 and names.count("dgr_backward") == 1
    for k, shape in dict(means3D=(64, 3), opacities=(64, 1), scales=(64, 3), rotations=(64, 4), shs=(64, 16, 3)).items():
        assert leaves[k].grad is not None and tuple(leaves[k].grad.shape) == shape, k
    assert tuple(m2d.grad.shape) == (64, 3)
    # every gradient sub-buffer handed to the library is 16-byte aligned
    bwd = [c for c in stub.calls if c[0] == "dgr_backward"][0][1]
    for ptr in bwd[24:33]:
        assert ptr is None or ptr % 16 == 0
    # second call for the same shape reuses the capacity and passes the previous frame's depth range as the hint
    n_before = len(stub.calls)
    dgr.GaussianRasterizer(rs)(means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"],
                               shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"])
    assert [c[0] for c in stub.calls[n_before:]].count("dgr_forward") == 1
    last = [c for c in stub.calls if c[0] == "dgr_forward"][-1][1]
    assert last[26] >= dgr._grow(stub.R) and last[26] % 32 == 0
    assert (last[32], last[33]) == (2.0, 6.0) and last[30] == 0x1000 and last[31] == 0x2000 and last[34] == 0
    # a frame that does not fit (R grows 100x) is re-run transparently with a larger capacity: two enqueues,
    # the second into the SAME output tensors, no exception
    stub.R = 100 * stub.R + 1_000_000
    n_before = len(stub.calls)
    color2, _ = dgr.GaussianRasterizer(rs)(means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"],
                                           shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"])
    fw = [c[1] for c in stub.calls[n_before:] if c[0] == "dgr_forward"]
    assert len(fw) == 2 and fw[1][26] >= stub.R > fw[0][26] and fw[0][20] == fw[1][20] == color2.data_ptr()
    assert len(dgr._Sizing.hint) == 1


def test_precomputed_inputs_and_errors(stub, monkeypatch):
    monkeypatch.setattr(dgr, "_sizes", lambda P, W, H, R: (4096, 60 * R + 128, 8192, 4096,
                                                           (60 * R + 128 + 511) // 512 * 512, 8192))
    sc = synth.gaussian_scene(n=32, seed=1)
    cam = synth.look_at_camera(width=32, height=32)
    rs = synth.raster_settings_for(cam, torch.zeros(3), settings_cls=dgr.GaussianRasterizationSettings)
    r = dgr.GaussianRasterizer(rs)
    cols = torch.rand(32, 3, requires_grad=True)
    cov = torch.rand(32, 6, requires_grad=True)
    m3 = sc["means3D"].clone().requires_grad_(True)
    color, _ = r(means3D=m3, means2D=None, opacities=sc["opacities"], colors_precomp=cols, cov3D_precomp=cov)
    color.sum().backward()
    assert tuple(cols.grad.shape) == (32, 3) and tuple(cov.grad.shape) == (32, 6) and tuple(m3.grad.shape) == (32, 3)
    fwd = [c for c in stub.calls if c[0] == "dgr_forward"][0][1]
    assert fwd[7] is None and fwd[8] is not None          # shs absent, colors_precomp present
    assert fwd[10] is None and fwd[12] is None and fwd[13] is not None
    with pytest.raises(Exception, match="excatly one of either SHs"):
        r(means3D=m3, means2D=None, opacities=sc["opacities"], scales=sc["scales"], rotations=sc["rotations"])
    with pytest.raises(Exception, match="scale/rotation pair"):
        r(means3D=m3, means2D=None, opacities=sc["opacities"], shs=sc["shs"], scales=sc["scales"])


def test_settings_fields_match_reference_order():
    assert dgr.GaussianRasterizationSettings._fields == (
        "image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix", "projmatrix",
        "sh_degree", "campos", "prefiltered", "debug")
    assert math.isclose(dgr._cap_from_bytes(60 * 4096 + 128), 4096)


def test_batch_api_glue(stub, monkeypatch):
    monkeypatch.setattr(dgr, "_sizes", lambda P, W, H, R: (4096, 60 * R + 128, 8192, 4096,
                                                           (60 * R + 128 + 511) // 512 * 512, 8192))
    sc = synth.gaussian_scene(n=48, seed=2)
    cams = [synth.look_at_camera(azimuth_deg=40.0 * k, width=32, height=32) for k in range(3)]
    sets = [synth.raster_settings_for(c, torch.ones(3), settings_cls=dgr.GaussianRasterizationSettings) for c in cams]
    leaves = {k: sc[k].clone().requires_grad_(True) for k in ("means3D", "opacities", "scales", "rotations", "shs")}
    m2d = torch.zeros(3, 48, 3, requires_grad=True)
    color, radii = dgr.BatchGaussianRasterizer(sets)(means3D=leaves["means3D"], means2D=m2d,
                                                     opacities=leaves["opacities"], shs=leaves["shs"],
                                                     scales=leaves["scales"], rotations=leaves["rotations"])
    assert color.shape == (3, 3, 32, 32) and radii.shape == (3, 48)
    color.backward(torch.ones_like(color))
    assert tuple(m2d.grad.shape) == (3, 48, 3) and tuple(leaves["shs"].grad.shape) == (48, 16, 3)
    fwd = [c for c in stub.calls if c[0] == "dgr_forward_batch"][0][1]
    assert fwd[0] == 3 and fwd[1] == 48 and fwd[15] == 0 and len(fwd[19]) == 3   # F, P, per_frame, tan_fovx array
    assert fwd[25] % 128 == 0 and fwd[27] % 128 == 0 and fwd[30] % 128 == 0  # workspace strides
    bwd = [c for c in stub.calls if c[0] == "dgr_backward_batch"][0][1]
    for ptr in bwd[29:37]:
        assert ptr is None or ptr % 16 == 0
    # per-frame deformed means / scales / rotations (DG-Mesh: one time per frame), shared opacity and SH
    pfl = {k: sc[k][None].repeat(3, 1, 1).clone().requires_grad_(True) for k in ("means3D", "scales", "rotations")}
    n_before = len(stub.calls)
    color, _ = dgr.BatchGaussianRasterizer(sets)(means3D=pfl["means3D"], means2D=None, opacities=leaves["opacities"],
                                                 shs=leaves["shs"], scales=pfl["scales"], rotations=pfl["rotations"])
    color.backward(torch.ones_like(color))
    assert [c for c in stub.calls[n_before:] if c[0] == "dgr_forward_batch"][0][1][15] == 1 | 2 | 4
    assert tuple(pfl["means3D"].grad.shape) == (3, 48, 3) and tuple(pfl["rotations"].grad.shape) == (3, 48, 4)
    assert tuple(pfl["scales"].grad.shape) == (3, 48, 3)
    with pytest.raises(ValueError, match="leading dimension"):
        dgr.BatchGaussianRasterizer(sets)(means3D=pfl["means3D"][:2], means2D=None, opacities=leaves["opacities"],
                                          shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"])
    mixed = list(sets)
    mixed[1] = mixed[1]._replace(image_width=48)
    with pytest.raises(ValueError, match="share image size"):
        dgr.BatchGaussianRasterizer(mixed)(means3D=leaves["means3D"], means2D=None, opacities=leaves["opacities"],
                                           shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"])


def test_batch_camera_cache_is_keyed_by_address_and_version(stub, monkeypatch):
    """The batch call stacks its cameras once per (tensor address, version counter): the same tensors again hit
    the cache, an in-place update or a fresh tensor does not.  (The GPU twin checks the rendered images.)"""
    monkeypatch.setattr(dgr, "_sizes", lambda P, W, H, R: (4096, 60 * R + 128, 8192, 4096,
                                                           (60 * R + 128 + 511) // 512 * 512, 8192))
    calls = []
    orig = dgr._stack_settings_uncached
    monkeypatch.setattr(dgr, "_stack_settings_uncached", lambda s, d: (calls.append(1), orig(s, d))[1])
    dgr._stacked.clear()
    sc = synth.gaussian_scene(n=16, seed=4)
    cams = [synth.look_at_camera(azimuth_deg=30.0 * k, width=32, height=32) for k in range(2)]
    sets = [synth.raster_settings_for(c, torch.ones(3), settings_cls=dgr.GaussianRasterizationSettings) for c in cams]

    def run(s):
        with torch.no_grad():
            dgr.BatchGaussianRasterizer(s)(means3D=sc["means3D"], means2D=None, opacities=sc["opacities"],
                                           shs=sc["shs"], scales=sc["scales"], rotations=sc["rotations"])
        return [c for c in stub.calls if c[0] == "dgr_forward_batch"][-1][1][16]      # pointer of the stacked views

    p1 = run(sets)
    p2 = run(sets)
    assert len(calls) == 1 and p1 == p2                                  # hit
    cams[0].world_view_transform.mul_(1.0)                               # in-place: version counter moves
    run(sets)
    assert len(calls) == 2
    fresh = [s._replace(viewmatrix=s.viewmatrix.clone()) for s in sets]  # new tensors, same values
    run(fresh)
    assert len(calls) == 3
    # the cache keeps the source tensors alive (their addresses cannot be recycled while an entry is valid)
    assert all(len(v[1]) == 2 for v in dgr._stacked.values())
    assert len(dgr._stacked) <= 8


def test_nvls_flat_grad_needs_a_process_group():
    import dp
    p = [torch.nn.Parameter(torch.zeros(8))]
    with pytest.raises(RuntimeError, match="process group"):
        dp.NvlsFlatGrad(p)
