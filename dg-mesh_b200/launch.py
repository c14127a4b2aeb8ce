"""Run a DG-Mesh script (train.py, render_test.py, ...) UNMODIFIED on top of this package.

    python /path/to/dg-mesh_b200/launch.py /path/to/DG-Mesh/dgmesh/train.py --config ... [script args]

Why a launcher: `python train.py` puts the script's directory at sys.path[0], ahead of PYTHONPATH, so
the reference's own `gaussian_renderer/`, `utils/time_utils.py`, `utils/renderer.py`,
`utils/loss_utils.py` and `nvdiffrast_utils/dpsr.py` would win over the drop-ins of the same name (only
`diff_gaussian_rasterization`, `simple_knn` and `diso`, which the script directory does not contain,
would be replaced).  `install()` therefore fixes the resolution order BEFORE the script runs:

  * `utils` and `nvdiffrast_utils` (namespace packages in the reference) become packages whose search
    path lists THIS package's directory first and the reference's second: `utils.time_utils`,
    `utils.renderer`, `utils.loss_utils`, `nvdiffrast_utils.dpsr` resolve here, every other submodule
    (`utils.general_utils`, `nvdiffrast_utils.regularizer`, ...) falls through to the reference;
  * `gaussian_renderer`, `diff_gaussian_rasterization`, `simple_knn`, `diso` and `nvdiffrast` (the CUDA
    triangle rasteriser that stands in for the third-party package) are imported from here and pinned in
    sys.modules;
  * once the reference's `scene` package is imported, `GaussianModelDPSRDynamicAnchor.densify_and_prune`
    is replaced by the fused device version (densify.py).

Nothing of the reference is modified on disk; `install()` is idempotent."""
import importlib
import importlib.machinery
import importlib.util
import os
import runpy
import sys
import types

PKG = os.path.dirname(os.path.abspath(__file__))
_MERGED = ("utils", "nvdiffrast_utils")
_PINNED = ("diff_gaussian_rasterization", "simple_knn", "simple_knn._C", "diso", "gaussian_renderer", "nvdiffrast",
           "nvdiffrast.torch")


def _merged_package(name, dirs):
    spec = importlib.machinery.ModuleSpec(name, None, is_package=True)
    spec.submodule_search_locations = list(dirs)
    m = types.ModuleType(name)
    m.__path__ = list(dirs)
    m.__spec__ = spec
    m.__package__ = name
    return m


class _PatchScene:
    """meta-path hook: after `scene.gaussian_model_dpsr_dynamic_anchor` has been imported, swap in the
    fused densify_and_prune and the device-side anchor_mesh."""

    def find_spec(self, fullname, path=None, target=None):
        if fullname != "scene.gaussian_model_dpsr_dynamic_anchor":
            return None
        sys.meta_path.remove(self)
        try:
            spec = importlib.util.find_spec(fullname)
        finally:
            sys.meta_path.insert(0, self)
        if spec is None or spec.loader is None:
            return None
        inner = spec.loader

        class Loader(importlib.abc.Loader):
            def create_module(self, s):
                return inner.create_module(s)

            def exec_module(self, module):
                inner.exec_module(module)
                import anchor
                import densify
                densify.install(module.GaussianModelDPSRDynamicAnchor)
                anchor.install(module.GaussianModelDPSRDynamicAnchor)

        spec.loader = Loader()
        return spec


def install(reference_dir):
    """Make the drop-ins win over `reference_dir` (the directory that holds train.py)."""
    import importlib.abc  # noqa: F401  (used by _PatchScene)
    reference_dir = os.path.abspath(reference_dir)
    if PKG not in sys.path:
        sys.path.insert(0, PKG)
    if reference_dir not in sys.path:
        sys.path.append(reference_dir)
    for name in _MERGED:
        dirs = [os.path.join(PKG, name), os.path.join(reference_dir, name)]
        cur = sys.modules.get(name)
        cur_path = list(getattr(cur, "__path__", []) or []) if cur is not None else []
        if cur_path and os.path.abspath(cur_path[0]) == dirs[0]:
            # already resolving this package first (e.g. a test harness): just make the reference reachable
            if dirs[1] not in cur_path:
                cur.__path__.append(dirs[1])
            continue
        # drop anything already resolved through another order
        for k in [k for k in sys.modules if k == name or k.startswith(name + ".")]:
            del sys.modules[k]
        sys.modules[name] = _merged_package(name, dirs)
    for name in _PINNED:
        mod = sys.modules.get(name)
        if mod is not None and not os.path.abspath(getattr(mod, "__file__", "") or "").startswith(PKG):
            del sys.modules[name]
        importlib.import_module(name)
    if not any(isinstance(h, _PatchScene) for h in sys.meta_path):
        sys.meta_path.insert(0, _PatchScene())


def resolved():
    """{module name: file} of the modules the drop-in is responsible for (diagnostics / tests)."""
    out = {}
    for name in ("utils.time_utils", "utils.renderer", "utils.loss_utils", "nvdiffrast_utils.dpsr") + _PINNED:
        try:
            out[name] = importlib.import_module(name).__file__
        except Exception as e:  # reported, not hidden
            out[name] = f"<{type(e).__name__}: {e}>"
    return out


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        print(__doc__)
        return 2
    script = os.path.abspath(argv[0])
    install(os.path.dirname(script))
    sys.argv = [script] + argv[1:]
    # runpy puts the script directory at sys.path[0]; the pinned / merged modules above already decide
    runpy.run_path(script, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())
