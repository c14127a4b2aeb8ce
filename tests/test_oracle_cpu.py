"""CPU tests of the oracle itself (no GPU): fp32 vs fp64 agreement, binning invariants,
and a finite-difference check that the oracle's backward is the gradient of its forward."""
import math

import numpy as np
import pytest

import util
from oracle.oracle import RasterOracle


@pytest.fixture(scope="module")
def scene():
    return util.small_scene(n=1500, W=128, H=80, seed=3)


def test_fp32_matches_fp64(scene):
    sc, cam = scene
    o32 = util.oracle_forward(RasterOracle(32), sc, cam, [1, 1, 1])
    o64 = util.oracle_forward(RasterOracle(64), sc, cam, [1, 1, 1])
    assert o32["num_rendered"] > 1000
    assert (o32["radii"] == o64["radii"]).mean() > 0.999
    assert util.rel_err(o32["color"], o64["color"]) < 1e-5
    dpix = np.random.default_rng(1).standard_normal((3, 80, 128)).astype(np.float32)
    g32, g64 = RasterOracle(32).backward(o32, dpix), RasterOracle(64).backward(o64, dpix)
    for k in g32:
        assert util.rel_err(g32[k], g64[k]) < 1e-4, k


def test_binning_invariants(scene):
    sc, cam = scene
    o = util.oracle_forward(RasterOracle(32), sc, cam, [0, 0, 0])
    keys, pl, ranges = o["point_list_keys"], o["point_list"], o["ranges"]
    assert len(keys) == o["num_rendered"] == int(o["tiles_touched"].sum())
    assert np.all(keys[:-1] <= keys[1:])                      # sorted by (tile, depth)
    tiles = (keys >> np.uint64(32)).astype(np.int64)
    for t in np.unique(tiles)[:50]:
        a, b = ranges[t]
        assert np.all(tiles[a:b] == t) and (b - a) == int((tiles == t).sum())
    # ties in (tile, depth) keep gaussian-id order (stable sort, emission order)
    same = keys[:-1] == keys[1:]
    assert np.all(pl[:-1][same] < pl[1:][same])
    # depth bits in the key are the fp32 depth of the listed gaussian
    dbits = o["depths"].astype(np.float32).view(np.uint32)[pl]
    assert np.all((keys & np.uint64(0xFFFFFFFF)).astype(np.uint32) == dbits)


def test_empty_and_culled():
    sc, cam = util.small_scene(n=50, W=32, H=32, seed=1)
    sc["means3D"][:] = sc["means3D"] * 0 + np.float32(100.0)  # everything behind / outside
    o = util.oracle_forward(RasterOracle(32), sc, cam, [0.2, 0.4, 0.6])
    assert o["num_rendered"] == 0 and np.all(o["radii"] == 0)
    assert np.allclose(o["color"][1], 0.4) and np.all(o["final_T"] == 1) and np.all(o["n_contrib"] == 0)


@pytest.mark.parametrize("use_colors,use_cov", [(False, False), (True, True)])
def test_backward_is_gradient_of_forward(use_colors, use_cov):
    """Central finite differences in fp64 on a tiny scene.  The forward is piecewise
    smooth (radius ceil, 1/255 and 1e-4 thresholds), so we demand agreement on the
    bulk of randomly probed coordinates rather than on every one."""
    sc, cam = util.small_scene(n=60, W=48, H=32, seed=5, scale=0.12)
    orc = RasterOracle(64)
    W, H = cam.image_width, cam.image_height
    dpix = np.random.default_rng(2).standard_normal((3, H, W))

    def loss(s):
        return float((util.oracle_forward(orc, s, cam, [0.3, 0.5, 0.7], use_colors=use_colors,
                                          use_cov=use_cov)["color"] * dpix).sum())

    sc64 = {k: v.double() for k, v in sc.items()}
    fwd = util.oracle_forward(orc, sc64, cam, [0.3, 0.5, 0.7], use_colors=use_colors, use_cov=use_cov)
    g = orc.backward(fwd, dpix)
    probes = [("means3D", "dL_dmean3D"), ("opacities", "dL_dopacity")]
    if not use_cov:
        probes += [("scales", "dL_dscale"), ("rotations", "dL_drot")]
    if not use_colors:
        probes += [("shs", "dL_dsh")]
    rng = np.random.default_rng(7)
    vis = np.nonzero(fwd["radii"] > 0)[0]
    bad = tot = 0
    for name, gname in probes:
        for _ in range(12):
            i = int(rng.choice(vis))
            flat = sc64[name][i].reshape(-1)
            j = int(rng.integers(flat.numel()))
            old = float(flat[j])
            eps = 1e-6 * max(1.0, abs(old))
            flat[j] = old + eps
            lp = loss(sc64)
            flat[j] = old - eps
            lm = loss(sc64)
            flat[j] = old
            fd = (lp - lm) / (2 * eps)
            an = float(np.asarray(g[gname][i]).reshape(-1)[j])
            tot += 1
            if abs(fd - an) > 1e-4 * max(1.0, abs(fd), abs(an)):
                bad += 1
    assert bad <= max(2, tot // 10), (bad, tot)
