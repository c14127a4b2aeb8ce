"""`python train.py` puts the script directory ahead of PYTHONPATH, so a plain PYTHONPATH drop-in leaves
`gaussian_renderer`, `utils.time_utils`, `utils.renderer`, `nvdiffrast_utils.dpsr` on the reference's copies.
dg-mesh_b200/launch.py fixes the resolution order; this runs a stand-in script directory through it (in a
subprocess: the launcher edits sys.modules) and checks which file every module resolved to."""
import json
import os
import subprocess
import sys
import textwrap


def test_launcher_makes_the_dropins_win_over_the_script_directory(tmp_path, repo_root):
    ref = tmp_path / "dgmesh"
    for d in ("utils", "nvdiffrast_utils", "gaussian_renderer", "scene"):
        (ref / d).mkdir(parents=True)
    # a reference-shaped tree: namespace packages utils / nvdiffrast_utils, regular package gaussian_renderer
    (ref / "utils" / "time_utils.py").write_text("WHO = 'reference'\n")
    (ref / "utils" / "renderer.py").write_text("WHO = 'reference'\ndef mesh_shape_renderer():\n    return 'ref-shape'\n")
    (ref / "utils" / "loss_utils.py").write_text("WHO = 'reference'\n")
    (ref / "utils" / "general_utils.py").write_text("WHO = 'reference'\n")
    (ref / "nvdiffrast_utils" / "dpsr.py").write_text("WHO = 'reference'\n")
    (ref / "nvdiffrast_utils" / "util.py").write_text("WHO = 'reference'\n")
    (ref / "gaussian_renderer" / "__init__.py").write_text("WHO = 'reference'\n")
    (ref / "scene" / "__init__.py").write_text("")
    (ref / "train.py").write_text(textwrap.dedent("""
        import json, sys
        import gaussian_renderer, diff_gaussian_rasterization, simple_knn._C, diso, nvdiffrast.torch
        import utils.time_utils, utils.renderer, utils.loss_utils, utils.general_utils
        import nvdiffrast_utils.dpsr, nvdiffrast_utils.util, nvdiffrast_utils.regularizer
        mods = [gaussian_renderer, diff_gaussian_rasterization, simple_knn._C, diso, nvdiffrast.torch, utils.time_utils,
                utils.renderer, utils.loss_utils, utils.general_utils, nvdiffrast_utils.dpsr, nvdiffrast_utils.util,
                nvdiffrast_utils.regularizer]
        out = {m.__name__: m.__file__ for m in mods}
        out["fallthrough"] = utils.renderer.mesh_shape_renderer()      # not defined by the drop-in
        out["argv"] = sys.argv[1:]
        print("RESOLVED " + json.dumps(out))
    """))
    pkg = os.path.join(repo_root, "dg-mesh_b200")
    r = subprocess.run([sys.executable, os.path.join(pkg, "launch.py"), str(ref / "train.py"), "--config", "x.yaml"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESOLVED ")][-1]
    out = json.loads(line[len("RESOLVED "):])
    ours = ("gaussian_renderer", "diff_gaussian_rasterization", "simple_knn._C", "diso", "nvdiffrast.torch",
            "utils.time_utils", "utils.renderer", "utils.loss_utils", "nvdiffrast_utils.dpsr",
            "nvdiffrast_utils.regularizer")
    for name in ours:
        assert out[name].startswith(pkg), (name, out[name])
    for name in ("utils.general_utils", "nvdiffrast_utils.util"):           # not replaced: the reference's own
        assert out[name].startswith(str(ref)), (name, out[name])
    assert out["fallthrough"] == "ref-shape" and out["argv"] == ["--config", "x.yaml"]
