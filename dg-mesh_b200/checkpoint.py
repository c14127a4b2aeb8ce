"""Checkpoint formats adjacent to the hot path.

1. The reference's Gaussian point cloud: a binary little-endian PLY with FOUR elements --
   `vertex` (x y z nx ny nz f_dc_* f_rest_* opacity scale_* rot_*, all float32), `density_thres`,
   `gaussian_center` (x/y/z) and `gaussian_scale` -- exactly as
   dgmesh/scene/gaussian_model_dpsr_dynamic_anchor.py:238-289 writes it through `plyfile` and :296-340
   reads it back (SH features are stored channel-major: `_features_dc.transpose(1, 2).flatten(1)`).
   Written / parsed here with numpy only (plyfile is not a dependency), so files interchange with the
   reference in both directions.
2. `save_training_state / load_training_state`: everything needed for an exact resume that the
   reference does not store -- iteration, every parameter group's optimiser state, the densification
   statistics -- in one torch file next to the PLY / per-network .pth files the reference keeps.
"""
import os

import numpy as np
import torch

_F4 = np.dtype("<f4")


def vertex_attributes(n_dc, n_rest, n_scale=3, n_rot=4):
    """Property order of the vertex element (construct_list_of_attributes, :238-251)."""
    names = ["x", "y", "z", "nx", "ny", "nz"]
    names += [f"f_dc_{i}" for i in range(n_dc)]
    names += [f"f_rest_{i}" for i in range(n_rest)]
    names.append("opacity")
    names += [f"scale_{i}" for i in range(n_scale)]
    names += [f"rot_{i}" for i in range(n_rot)]
    return names


def _np(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def save_gaussians_ply(path, xyz, normal, features_dc, features_rest, opacity, scaling, rotation, density_thres,
                       gaussian_center, gaussian_scale):
    """features_dc [P,1,3], features_rest [P,K,3] as the model holds them (SH index before channel)."""
    xyz, normal = _np(xyz).astype(_F4), _np(normal).astype(_F4)
    P = xyz.shape[0]
    f_dc = np.ascontiguousarray(_np(features_dc).transpose(0, 2, 1).reshape(P, -1)).astype(_F4)
    f_rest = np.ascontiguousarray(_np(features_rest).transpose(0, 2, 1).reshape(P, -1)).astype(_F4)
    opacity = _np(opacity).reshape(P, 1).astype(_F4)
    scaling, rotation = _np(scaling).astype(_F4), _np(rotation).astype(_F4)
    table = np.concatenate([xyz, normal, f_dc, f_rest, opacity, scaling, rotation], axis=1).astype(_F4)
    names = vertex_attributes(f_dc.shape[1], f_rest.shape[1], scaling.shape[1], rotation.shape[1])
    assert table.shape[1] == len(names)
    center = _np(gaussian_center).reshape(-1).astype(_F4)
    header = ["ply", "format binary_little_endian 1.0", f"element vertex {P}"]
    header += [f"property float {n}" for n in names]
    header += ["element density_thres 1", "property float density_thres",
               "element gaussian_center 1", "property float gaussian_center_x", "property float gaussian_center_y",
               "property float gaussian_center_z",
               "element gaussian_scale 1", "property float gaussian_scale", "end_header"]
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        f.write(np.ascontiguousarray(table).tobytes())
        f.write(np.asarray(_np(density_thres), dtype=_F4).reshape(1).tobytes())
        f.write(center[:3].tobytes())
        f.write(np.asarray(_np(gaussian_scale), dtype=_F4).reshape(-1)[:1].tobytes())


def _parse_header(f):
    if f.readline().strip() != b"ply":
        raise ValueError("not a PLY file")
    fmt, elements = None, []
    while True:
        line = f.readline()
        if not line:
            raise ValueError("PLY header not terminated")
        tok = line.decode("ascii").split()
        if not tok or tok[0] == "comment":
            continue
        if tok[0] == "format":
            fmt = tok[1]
        elif tok[0] == "element":
            elements.append((tok[1], int(tok[2]), []))
        elif tok[0] == "property":
            if tok[1] == "list":
                raise ValueError("list properties are not part of this format")
            elements[-1][2].append((tok[2], tok[1]))
        elif tok[0] == "end_header":
            break
    if fmt != "binary_little_endian":
        raise ValueError(f"unsupported PLY format {fmt!r} (the reference writes binary_little_endian)")
    return elements


_TYPES = {"float": "<f4", "float32": "<f4", "double": "<f8", "float64": "<f8", "int": "<i4", "int32": "<i4",
          "uint": "<u4", "uchar": "u1", "uint8": "u1", "short": "<i2", "ushort": "<u2"}


def load_gaussians_ply(path, max_sh_degree=3):
    """-> dict of float32 numpy arrays shaped as the model holds them (features_* as [P,K,3])."""
    with open(path, "rb") as f:
        elements = _parse_header(f)
        data = {}
        for name, count, props in elements:
            dt = np.dtype([(p, _TYPES[t]) for p, t in props])
            data[name] = np.frombuffer(f.read(dt.itemsize * count), dtype=dt, count=count)
    v = data["vertex"]
    col = lambda *names: np.stack([v[n].astype(np.float32) for n in names], axis=1)  # noqa: E731
    P = v.shape[0]
    rest = sorted((n for n in v.dtype.names if n.startswith("f_rest_")), key=lambda n: int(n.split("_")[-1]))
    if len(rest) != 3 * (max_sh_degree + 1) ** 2 - 3:
        raise ValueError(f"{len(rest)} f_rest_* properties do not match SH degree {max_sh_degree}")
    scales = sorted((n for n in v.dtype.names if n.startswith("scale_")), key=lambda n: int(n.split("_")[-1]))
    rots = sorted((n for n in v.dtype.names if n.startswith("rot")), key=lambda n: int(n.split("_")[-1]))
    out = {
        "xyz": col("x", "y", "z"), "normal": col("nx", "ny", "nz"),
        "features_dc": col("f_dc_0", "f_dc_1", "f_dc_2").reshape(P, 3, 1).transpose(0, 2, 1).copy(),
        "features_rest": (col(*rest).reshape(P, 3, -1).transpose(0, 2, 1).copy() if rest
                          else np.zeros((P, 0, 3), np.float32)),
        "opacity": col("opacity"), "scaling": col(*scales), "rotation": col(*rots),
    }
    if "density_thres" in data:
        out["density_thres"] = data["density_thres"]["density_thres"].astype(np.float32)
        c = data["gaussian_center"]
        out["gaussian_center"] = np.stack([c["gaussian_center_x"], c["gaussian_center_y"], c["gaussian_center_z"]],
                                          axis=1).astype(np.float32)
        out["gaussian_scale"] = data["gaussian_scale"]["gaussian_scale"].astype(np.float32)
    return out


def _rng_state():
    """Every generator the training loop draws from: python `random` (camera pick, train.py:150), numpy, the
    torch CPU generator and all CUDA generators (torch.randn(..., device='cuda'), torch.normal in densification)."""
    import random
    np_state = np.random.get_state()
    st = {"torch": torch.get_rng_state(),
          "python": torch.tensor(list(random.getstate()[1]), dtype=torch.int64),
          "python_version": int(random.getstate()[0]),
          "numpy_keys": torch.from_numpy(np_state[1].astype(np.int64)),
          "numpy_meta": torch.tensor([int(np_state[2]), int(np_state[3])], dtype=torch.int64),
          "numpy_gauss": float(np_state[4])}
    if torch.cuda.is_available():
        st["cuda"] = torch.cuda.get_rng_state_all()
    return st


def _restore_rng(st):
    import random
    torch.set_rng_state(st["torch"].cpu())
    if "python" in st:
        random.setstate((int(st["python_version"]), tuple(int(x) for x in st["python"].tolist()), None))
    if "numpy_keys" in st:
        np.random.set_state(("MT19937", st["numpy_keys"].numpy().astype(np.uint32), int(st["numpy_meta"][0]),
                             int(st["numpy_meta"][1]), float(st["numpy_gauss"])))
    if "cuda" in st and torch.cuda.is_available() and len(st["cuda"]) == torch.cuda.device_count():
        torch.cuda.set_rng_state_all([t.cpu() for t in st["cuda"]])


def save_training_state(path, iteration, optimizers, extra=None):
    """optimizers: {name: torch.optim.Optimizer}; extra: {name: tensor} (e.g. xyz_gradient_accum, denom,
    max_radii2D) -- with the PLY, the networks' .pth files and the generator states kept here (python, numpy,
    torch CPU and CUDA) a resume is exact.  The payload is tensors / numbers / strings only, so it is read
    back with `weights_only=True` (no arbitrary unpickling)."""
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save({"format": "dgmesh_b200.training_state.v2", "iteration": int(iteration),
                "optimizers": {k: o.state_dict() for k, o in optimizers.items()},
                "extra": {k: v.detach().cpu() for k, v in (extra or {}).items()},
                "rng": _rng_state()}, path)


def load_training_state(path, optimizers=None, map_location="cpu", restore_rng=True):
    st = torch.load(path, map_location=map_location, weights_only=True)
    if st.get("format") not in ("dgmesh_b200.training_state.v1", "dgmesh_b200.training_state.v2"):
        raise ValueError("not a dgmesh_b200 training state")
    for k, o in (optimizers or {}).items():
        o.load_state_dict(st["optimizers"][k])
    if restore_rng and "rng" in st:
        _restore_rng(st["rng"])
    return st
