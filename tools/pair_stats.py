#!/usr/bin/env python
"""Diagnostic (needs the -DDGM_COUNT_PAIRS build: see NOTES_R2.md): how many (warp, Gaussian) pairs the blend
kernel evaluates after the exact cull, and in how many of them at least one lane actually blends the Gaussian.
    DGMESH_B200_LIB=build_dbg/libdgmesh_b200_dbg.so python tools/pair_stats.py"""
import ctypes
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dg-mesh_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import _dgm_lib  # noqa: E402
import diff_gaussian_rasterization as dgr  # noqa: E402
import synth  # noqa: E402

lib = _dgm_lib.lib()
fn = lib.dgm_debug_pair_counters
fn.restype, fn.argtypes = ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]
dev = torch.device("cuda")
out = {}
for name, (n, W, H) in {"c2": (100_000, 800, 800), "c5": (500_000, 1920, 1080)}.items():
    sc = synth.gaussian_scene(n=n, seed=0, device=dev)
    bg = torch.ones(3, device=dev)
    buf = (ctypes.c_ulonglong * 4)()
    fn(buf, 1)
    R = 0
    for k in range(8):
        cam = synth.look_at_camera(azimuth_deg=45.0 * k, elevation_deg=20.0, radius=4.0, width=W, height=H,
                                   fovx=2 * math.atan(math.tan(0.6911 / 2) * W / H), fovy=0.6911, device=dev)
        rs = synth.raster_settings_for(cam, bg, settings_cls=dgr.GaussianRasterizationSettings)
        with torch.no_grad():
            color, radii = dgr.GaussianRasterizer(rs)(means3D=sc["means3D"], means2D=None, opacities=sc["opacities"],
                                                      shs=sc["shs"], scales=sc["scales"], rotations=sc["rotations"])
    fn(buf, 0)
    ev, contrib, before = int(buf[0]), int(buf[1]), int(buf[2])
    out[name] = {"warp_gaussian_pairs_staged": before, "evaluated_after_cull": ev, "with_a_contributing_lane": contrib,
                 "cull_keeps": ev / max(before, 1), "contributing_of_evaluated": contrib / max(ev, 1)}
print(json.dumps(out))
