#!/usr/bin/env python
"""Raster forward+backward at any named configuration, this implementation and the stock
extension side by side on one GPU through the one-frame `GaussianRasterizer` API:
    python tools/raster_config_bench.py --config c5      # 500k Gaussians, 1920x1080 (BASELINE configs[4])
    python tools/raster_config_bench.py --config c2      # 100k Gaussians, 800x800
Prints one JSON object (ms per frame fwd+bwd, frames/s, R, speed-up)."""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dg-mesh_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import synth  # noqa: E402
import util  # noqa: E402

CONFIGS = {"c2": (100_000, 800, 800), "c5": (500_000, 1920, 1080),
           # an OBJECT-shaped scene (what DG-Mesh trains on): 200k Gaussians on a thick shell of radius 0.6 that
           # covers ~14 % of the image -> thousands of instances per tile, narrow per-tile depth ranges
           "object": (200_000, 800, 800)}


def run(dgr, sc, cams, dpix, bg, W, H, steps):
    leaves = {k: sc[k].clone().requires_grad_(True) for k in ("means3D", "opacities", "scales", "rotations", "shs")}

    def frame(k):
        cam = cams[k]
        rs = dgr.GaussianRasterizationSettings(
            image_height=H, image_width=W, tanfovx=math.tan(cam.FoVx * 0.5), tanfovy=math.tan(cam.FoVy * 0.5), bg=bg,
            scale_modifier=1.0, viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform, sh_degree=3,
            campos=cam.camera_center, prefiltered=False, debug=False)
        m2d = torch.zeros_like(leaves["means3D"], requires_grad=True)
        color, _ = dgr.GaussianRasterizer(rs)(means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"],
                                              shs=leaves["shs"], scales=leaves["scales"],
                                              rotations=leaves["rotations"])
        color.backward(dpix)

    for k in range(len(cams)):
        frame(k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        for k in range(len(cams)):
            frame(k)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (steps * len(cams))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c5", choices=sorted(CONFIGS))
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    n, W, H = CONFIGS[a.config]
    dev = torch.device("cuda")
    sc = synth.gaussian_scene(n=n, seed=0, device=dev)
    if a.config == "object":
        g = torch.Generator().manual_seed(3)
        d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=1)
        sc["means3D"] = (d * 0.6 + 0.01 * torch.randn(n, 3, generator=g)).to(dev)
    cams = [synth.look_at_camera(azimuth_deg=45.0 * k, elevation_deg=20.0, radius=4.0, width=W, height=H,
                                 fovx=2 * math.atan(math.tan(0.6911 / 2) * W / H), fovy=0.6911, device=dev)
            for k in range(8)]
    dpix = torch.randn(3, H, W, generator=torch.Generator().manual_seed(1)).to(dev)
    bg = torch.ones(3, device=dev)
    import diff_gaussian_rasterization as ours
    out = {"config": a.config, "gaussians": n, "width": W, "height": H, "views": 8}
    ms = run(ours, sc, cams, dpix, bg, W, H, a.steps)
    out["ours_ms_per_frame"], out["ours_frames_per_s"] = ms, 1e3 / ms
    import _dgm_lib
    _dgm_lib.lib().dgm_profile_enable(1)
    run(ours, sc, cams[:1], dpix, bg, W, H, 1)
    out["ours_kernel_us_one_frame"] = {k: round(v * 1e3, 1) for k, v in _dgm_lib.profile_read().items()}
    _dgm_lib.lib().dgm_profile_enable(0)
    st = ours._Notify.get(dev.index).words
    out["num_rendered"], out["max_tile_length"] = int(st[0]), int(st[2])
    ref = util.load_reference_rasterizer()
    if ref is not None:
        msr = run(ref, sc, cams, dpix, bg, W, H, a.steps)
        out["reference_ms_per_frame"], out["reference_frames_per_s"], out["speedup"] = msr, 1e3 / msr, msr / ms
    print(json.dumps(out))
