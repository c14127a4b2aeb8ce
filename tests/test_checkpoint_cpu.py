"""Checkpoint formats: the reference's 4-element PLY written / read without plyfile,
and the exact-resume training state."""
import numpy as np
import pytest
import torch

from dgmesh_b200 import checkpoint as ck


def _model(P=57, deg=3, seed=0):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    K = (deg + 1) ** 2 - 1
    return dict(xyz=r(P, 3), normal=r(P, 3), features_dc=r(P, 1, 3), features_rest=r(P, K, 3), opacity=r(P, 1),
                scaling=r(P, 3), rotation=r(P, 4), density_thres=torch.tensor(0.0123),
                gaussian_center=r(1, 3), gaussian_scale=torch.tensor([1.56]))


@pytest.mark.parametrize("deg", [3, 1, 0])
def test_ply_roundtrip_is_exact_and_layout_matches_the_reference(tmp_path, deg):
    m = _model(deg=deg)
    path = str(tmp_path / "point_cloud" / "iteration_7" / "point_cloud.ply")
    ck.save_gaussians_ply(path, **m)
    raw = open(path, "rb").read()
    head = raw[:raw.index(b"end_header\n")].decode().split("\n")
    assert head[:3] == ["ply", "format binary_little_endian 1.0", "element vertex 57"]
    props = [l.split()[-1] for l in head if l.startswith("property")]
    n_rest = 3 * ((deg + 1) ** 2 - 1)
    assert props[:9] == ["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"]
    assert props[9:9 + n_rest] == [f"f_rest_{i}" for i in range(n_rest)]
    assert props[9 + n_rest:] == ["opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3",
                                  "density_thres", "gaussian_center_x", "gaussian_center_y", "gaussian_center_z",
                                  "gaussian_scale"]
    elems = [l.split()[1:] for l in head if l.startswith("element")]
    assert elems == [["vertex", "57"], ["density_thres", "1"], ["gaussian_center", "1"], ["gaussian_scale", "1"]]
    # body size: P rows of float32 + 1 + 3 + 1 floats
    body = len(raw) - raw.index(b"end_header\n") - len(b"end_header\n")
    assert body == 4 * (57 * (17 + n_rest) + 5)
    back = ck.load_gaussians_ply(path, max_sh_degree=deg)
    for k in ("xyz", "normal", "features_dc", "features_rest", "opacity", "scaling", "rotation"):
        assert back[k].shape == tuple(m[k].shape), k
        assert np.array_equal(back[k], m[k].numpy()), k
    assert np.float32(back["density_thres"][0]) == np.float32(0.0123)
    assert np.array_equal(back["gaussian_center"], m["gaussian_center"].numpy())
    assert back["gaussian_scale"][0] == np.float32(1.56)
    # channel-major SH storage, like `_features_rest.transpose(1, 2).flatten(start_dim=1)` (:258)
    if deg:
        row0 = np.frombuffer(raw[raw.index(b"end_header\n") + 11:][:4 * (17 + n_rest)], dtype="<f4")
        assert np.array_equal(row0[9:9 + n_rest], m["features_rest"][0].t().reshape(-1).numpy())


def test_ply_errors(tmp_path):
    p = tmp_path / "x.ply"
    p.write_bytes(b"ply\nformat ascii 1.0\nelement vertex 0\nend_header\n")
    with pytest.raises(ValueError):
        ck.load_gaussians_ply(str(p))
    m = _model(deg=1)
    ck.save_gaussians_ply(str(p), **m)
    with pytest.raises(ValueError):
        ck.load_gaussians_ply(str(p), max_sh_degree=3)     # f_rest count does not match


def test_training_state_resume_is_exact(tmp_path):
    torch.manual_seed(0)
    w = torch.nn.Parameter(torch.randn(5, 3))
    opt = torch.optim.Adam([{"params": [w], "lr": 1e-2, "name": "xyz"}], eps=1e-15)
    for _ in range(3):
        opt.zero_grad()
        (w ** 2).sum().backward()
        opt.step()
    path = str(tmp_path / "state" / "iteration_3.pt")
    ck.save_training_state(path, 3, {"gaussians": opt}, extra={"denom": torch.arange(5.)})
    w_saved = w.detach().clone()
    # continue two steps
    ref = []
    for _ in range(2):
        opt.zero_grad()
        (w ** 2).sum().backward()
        opt.step()
        ref.append(w.detach().clone())
    # resume from the file into a fresh optimiser
    w2 = torch.nn.Parameter(w_saved.clone())
    opt2 = torch.optim.Adam([{"params": [w2], "lr": 1e-2, "name": "xyz"}], eps=1e-15)
    st = ck.load_training_state(path, {"gaussians": opt2})
    assert st["iteration"] == 3 and torch.equal(st["extra"]["denom"], torch.arange(5.))
    for k in range(2):
        opt2.zero_grad()
        (w2 ** 2).sum().backward()
        opt2.step()
        assert torch.equal(w2.detach(), ref[k])


def test_training_state_restores_every_generator_and_loads_without_pickle(tmp_path):
    """the loop draws from python `random`, numpy and torch (CPU + CUDA): all are saved and restored;
    the file loads with weights_only=True."""
    import random
    import numpy as np
    import checkpoint
    p = torch.nn.Parameter(torch.randn(4, 3))
    opt = torch.optim.Adam([p], lr=0.1)
    p.grad = torch.ones_like(p)
    opt.step()
    random.seed(3), np.random.seed(4), torch.manual_seed(5)
    path = str(tmp_path / "state.pt")
    checkpoint.save_training_state(path, 7, {"g": opt}, {"denom": torch.arange(5.0)})
    want = (random.random(), float(np.random.rand()), float(torch.rand(1)))
    random.seed(0), np.random.seed(0), torch.manual_seed(0)
    opt2 = torch.optim.Adam([p], lr=0.1)
    st = checkpoint.load_training_state(path, {"g": opt2})
    assert st["iteration"] == 7 and torch.equal(st["extra"]["denom"], torch.arange(5.0))
    assert (random.random(), float(np.random.rand()), float(torch.rand(1))) == want
    assert torch.equal(opt2.state_dict()["state"][0]["exp_avg"], opt.state_dict()["state"][0]["exp_avg"])
