"""Drop-in surface: the public functions / classes of this package take the SAME arguments (names, order,
defaults) as the reference's, read straight from the reference sources with `ast` (they cannot be
imported here: CUDA extensions, OpenGL).  Skipped where /root/reference is not mounted."""
import ast
import importlib
import inspect
import os

import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")


def _ref_def(rel, name, cls=None):
    tree = ast.parse(open(os.path.join(REF, rel)).read())
    scope = tree.body
    if cls:
        scope = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls).body
    return next(n for n in scope if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name == name)


def _ref_args(fn):
    a = fn.args
    names = [x.arg for x in a.args]
    defaults = [ast.literal_eval(d) if isinstance(d, (ast.Constant, ast.List, ast.UnaryOp)) else "<expr>"
                for d in a.defaults]
    return names, defaults


def _our_args(fn):
    sig = inspect.signature(fn)
    names = list(sig.parameters)
    defaults = [p.default for p in sig.parameters.values() if p.default is not inspect.Parameter.empty]
    return names, defaults


def _same(ref_fn, our_fn, allow_extra=()):
    rn, rd = _ref_args(ref_fn)
    on, od = _our_args(our_fn)
    on_core = [n for n in on if n not in allow_extra]
    assert on_core[:len(rn)] == rn, (rn, on)
    k = len(rd)
    ours_tail = od[:k] if allow_extra else od[-k:] if k else []
    for a, b in zip(rd, ours_tail):
        if a != "<expr>":
            assert a == b or (isinstance(a, list) and list(b) == a), (rn, rd, od)


def test_render_signature_and_result_keys():
    ours = importlib.import_module("gaussian_renderer")
    ref = _ref_def("dgmesh/gaussian_renderer/__init__.py", "render")
    _same(ref, ours.render)
    src = ast.get_source_segment(open(os.path.join(REF, "dgmesh/gaussian_renderer/__init__.py")).read(), ref)
    keys = ["render", "viewspace_points", "visibility_filter", "radii"]
    assert all(f'"{k}"' in src for k in keys)
    our_src = inspect.getsource(ours.render)
    assert all(f'"{k}"' in our_src for k in keys)


def test_rasterizer_surface():
    dgr = importlib.import_module("diff_gaussian_rasterization")
    rel = "dgmesh/submodules/diff-gaussian-rasterization/diff_gaussian_rasterization/__init__.py"
    if not os.path.exists(os.path.join(REF, rel)):
        import glob
        c = glob.glob(os.path.join(REF, "**/diff_gaussian_rasterization/__init__.py"), recursive=True)
        assert c
        rel = os.path.relpath(c[0], REF)
    settings = _ref_def(rel, "GaussianRasterizationSettings")
    fields = [n.target.id for n in settings.body if isinstance(n, ast.AnnAssign)]
    assert list(dgr.GaussianRasterizationSettings._fields) == fields
    _same(_ref_def(rel, "forward", cls="GaussianRasterizer"), dgr.GaussianRasterizer.forward)
    _same(_ref_def(rel, "markVisible", cls="GaussianRasterizer"), dgr.GaussianRasterizer.markVisible)
    _same(_ref_def(rel, "rasterize_gaussians"), dgr.rasterize_gaussians)


def test_mlp_dpsr_and_mesh_renderer_surfaces():
    tu = importlib.import_module("utils.time_utils")
    for cls in ("DeformNetwork", "DeformNetworkNormal", "DeformNetworkNormalSep", "AppearanceNetwork"):
        _same(_ref_def("dgmesh/utils/time_utils.py", "__init__", cls=cls), getattr(tu, cls).__init__)
        _same(_ref_def("dgmesh/utils/time_utils.py", "forward", cls=cls), getattr(tu, cls).forward)
    dpsr = importlib.import_module("nvdiffrast_utils.dpsr")
    _same(_ref_def("dgmesh/nvdiffrast_utils/dpsr.py", "__init__", cls="DPSR"), dpsr.DPSR.__init__)
    _same(_ref_def("dgmesh/nvdiffrast_utils/dpsr.py", "forward", cls="DPSR"), dpsr.DPSR.forward)
    rnd = importlib.import_module("utils.renderer")
    _same(_ref_def("dgmesh/utils/renderer.py", "mesh_renderer"), rnd.mesh_renderer)
