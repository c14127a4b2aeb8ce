"""numpy front-end of the CPU oracle (oracle/liboracle.so, built from raster_oracle.c).

TEST INFRASTRUCTURE ONLY -- see oracle/README.md.  Nothing under dg-mesh_b200/ may
import this module.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `make oracle/liboracle.so`")
        _LIB = ctypes.CDLL(path)
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _c(a, dt):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


class RasterOracle:
    """precision: 32 (reference op order in fp32) or 64 (same algorithm in fp64)."""

    def __init__(self, precision=32):
        assert precision in (32, 64)
        self.prec = precision
        self.dt = np.float32 if precision == 32 else np.float64
        self.creal = ctypes.c_float if precision == 32 else ctypes.c_double

    def _fn(self, name):
        return getattr(_lib(), f"{name}_{self.prec}")

    def forward(self, means3D, opacities, view, proj, campos, W, H, tan_fovx, tan_fovy, bg, shs=None, degree=0,
                colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None, scale_modifier=1.0):
        dt, cr = self.dt, self.creal
        means3D = _c(means3D, dt)
        P = means3D.shape[0]
        shs_, col_ = _c(shs, dt), _c(colors_precomp, dt)
        M = 0 if shs_ is None else shs_.shape[1]
        sc_, ro_, cv_ = _c(scales, dt), _c(rotations, dt), _c(cov3D_precomp, dt)
        op_ = _c(np.asarray(opacities).reshape(-1), dt)
        view_, proj_, cam_, bg_ = _c(view, dt).reshape(-1), _c(proj, dt).reshape(-1), _c(campos, dt), _c(bg, dt)
        o = dict(radii=np.zeros(P, np.int32), means2D=np.zeros((P, 2), dt), depths=np.zeros(P, dt),
                 cov3D=np.zeros((P, 6), dt), rgb=np.zeros((P, 3), dt), conic_opacity=np.zeros((P, 4), dt),
                 tiles_touched=np.zeros(P, np.uint32), clamped=np.zeros((P, 3), np.uint8))
        f = self._fn("orc_preprocess")
        f.restype = None
        f(ctypes.c_int(P), ctypes.c_int(degree), ctypes.c_int(M), _p(means3D), _p(sc_), cr(scale_modifier), _p(ro_),
          _p(op_), _p(shs_), _p(cv_), _p(col_), _p(view_), _p(proj_), _p(cam_), ctypes.c_int(W), ctypes.c_int(H),
          cr(tan_fovx), cr(tan_fovy), _p(o["radii"]), _p(o["means2D"]), _p(o["depths"]), _p(o["cov3D"]), _p(o["rgb"]),
          _p(o["conic_opacity"]), _p(o["tiles_touched"]), _p(o["clamped"]))
        R = int(o["tiles_touched"].astype(np.int64).sum())
        T = ((W + 15) // 16) * ((H + 15) // 16)
        o["point_list_keys"] = np.zeros(max(R, 1), np.uint64)
        o["point_list"] = np.zeros(max(R, 1), np.uint32)
        o["ranges"] = np.zeros((T, 2), np.uint32)
        fb = self._fn("orc_bin")
        fb.restype = ctypes.c_int64
        n = fb(ctypes.c_int(P), ctypes.c_int(W), ctypes.c_int(H), _p(o["radii"]), _p(o["means2D"]), _p(o["depths"]),
               _p(o["point_list_keys"]), _p(o["point_list"]), _p(o["ranges"]), ctypes.c_int64(max(R, 1)))
        assert n == R, (n, R)
        o["point_list_keys"], o["point_list"] = o["point_list_keys"][:R], o["point_list"][:R]
        o["num_rendered"] = R
        feats = col_ if col_ is not None else o["rgb"]
        o["final_T"] = np.zeros(H * W, dt)
        o["n_contrib"] = np.zeros(H * W, np.uint32)
        o["color"] = np.zeros((3, H, W), dt)
        fr = self._fn("orc_render_fwd")
        fr.restype = None
        fr(ctypes.c_int(W), ctypes.c_int(H), _p(o["ranges"]), _p(o["point_list"]), _p(o["means2D"]), _p(feats),
           _p(o["conic_opacity"]), _p(bg_), _p(o["final_T"]), _p(o["n_contrib"]), _p(o["color"]))
        o["_inputs"] = dict(means3D=means3D, shs=shs_, colors=col_, scales=sc_, rotations=ro_, cov3D_precomp=cv_,
                            view=view_, proj=proj_, campos=cam_, bg=bg_, W=W, H=H, tan_fovx=tan_fovx,
                            tan_fovy=tan_fovy, degree=degree, M=M, scale_modifier=scale_modifier, feats=feats)
        return o

    def backward(self, fwd, dL_dpix):
        dt, cr = self.dt, self.creal
        i = fwd["_inputs"]
        P, W, H, M = i["means3D"].shape[0], i["W"], i["H"], i["M"]
        dpix = _c(dL_dpix, dt)
        g = dict(dL_dmean2D=np.zeros((P, 3), dt), dL_dconic=np.zeros((P, 4), dt), dL_dopacity=np.zeros(P, dt),
                 dL_dcolor=np.zeros((P, 3), dt), dL_dmean3D=np.zeros((P, 3), dt), dL_dcov3D=np.zeros((P, 6), dt),
                 dL_dsh=np.zeros((P, max(M, 1), 3), dt), dL_dscale=np.zeros((P, 3), dt), dL_drot=np.zeros((P, 4), dt))
        f = self._fn("orc_render_bwd")
        f.restype = None
        f(ctypes.c_int(P), ctypes.c_int(W), ctypes.c_int(H), _p(fwd["ranges"]), _p(fwd["point_list"]), _p(i["bg"]),
          _p(fwd["means2D"]), _p(fwd["conic_opacity"]), _p(i["feats"]), _p(fwd["final_T"]), _p(fwd["n_contrib"]),
          _p(dpix), _p(g["dL_dmean2D"]), _p(g["dL_dconic"]), _p(g["dL_dopacity"]), _p(g["dL_dcolor"]))
        cov3D = i["cov3D_precomp"] if i["cov3D_precomp"] is not None else fwd["cov3D"]
        f2 = self._fn("orc_preprocess_bwd")
        f2.restype = None
        f2(ctypes.c_int(P), ctypes.c_int(i["degree"]), ctypes.c_int(M), _p(i["means3D"]), _p(fwd["radii"]),
           _p(i["shs"]), _p(fwd["clamped"]), _p(i["scales"]), _p(i["rotations"]), cr(i["scale_modifier"]), _p(cov3D),
           _p(i["view"]), _p(i["proj"]), ctypes.c_int(W), ctypes.c_int(H), cr(i["tan_fovx"]), cr(i["tan_fovy"]),
           _p(i["campos"]), _p(g["dL_dmean2D"]), _p(g["dL_dconic"]), _p(g["dL_dcolor"]), _p(g["dL_dmean3D"]),
           _p(g["dL_dcov3D"]), _p(g["dL_dsh"]) if i["shs"] is not None else None, _p(g["dL_dscale"]),
           _p(g["dL_drot"]))
        if M == 0:
            g["dL_dsh"] = np.zeros((P, 0, 3), dt)
        return g


def knn_mean_dist2(points):
    """CPU restatement of simple-knn's contract (simple_knn.cu:147-183): for every point the mean of
    the squared distances to its 3 nearest OTHER points (duplicates count with distance 0); fewer than
    3 neighbours leave FLT_MAX slots exactly as the reference's initial values do."""
    from scipy.spatial import cKDTree
    pts = np.ascontiguousarray(points, dtype=np.float64)
    P = pts.shape[0]
    k = min(4, P)
    d, _ = cKDTree(pts).query(pts, k=k)
    d = d.reshape(P, k)[:, 1:] ** 2          # drop the point itself (distance 0 at rank 0)
    best = np.full((P, 3), np.finfo(np.float32).max, dtype=np.float64)
    best[:, :d.shape[1]] = d
    with np.errstate(over="ignore"):
        return ((best[:, 0].astype(np.float32) + best[:, 1].astype(np.float32)) + best[:, 2].astype(np.float32)) \
            / np.float32(3.0)


# ----------------------------------------------------------------------------- DPSR
def _dpsr_stencil(pts, G):
    """ind0 / ind1 / per-axis weights exactly as point_rasterize / grid_interp form them
    (dgmesh/nvdiffrast_utils/dpsr_utils.py:158-176), in fp32."""
    f32 = np.float32
    cs = f32(1.0) / f32(G)
    t = (pts.astype(f32) / cs).astype(f32)
    fl, ce = np.floor(t), np.ceil(t)
    i0 = fl.astype(np.int64)
    i1 = np.fmod(ce, f32(G)).astype(np.int64)
    xyz0, xyz1 = (fl * cs).astype(f32), ((fl + f32(1)) * cs).astype(f32)
    w0 = (np.abs(pts.astype(f32) - xyz1) / cs).astype(f32)   # weight of node ind0: distance to the opposite corner
    w1 = (np.abs(pts.astype(f32) - xyz0) / cs).astype(f32)
    return i0, i1, w0, w1


def dpsr_forward_np(V, N, G, sig):
    """numpy restatement of DPSR.forward (dgmesh/nvdiffrast_utils/dpsr.py:28-70), batch 1:
    V [n,3] in (0,1), N [n,3] -> phi [G,G,G] (fp32 semantics, fp64 FFT)."""
    f32 = np.float32
    V, N = np.asarray(V, f32), np.asarray(N, f32)
    i0, i1, w0, w1 = _dpsr_stencil(V, G)
    ras = np.zeros((3, G, G, G), np.float64)
    for a in (0, 1):
        for b in (0, 1):
            for c in (0, 1):
                ix, iy, iz = (i1 if a else i0)[:, 0], (i1 if b else i0)[:, 1], (i1 if c else i0)[:, 2]
                w = ((w1 if a else w0)[:, 0] * (w1 if b else w0)[:, 1] * (w1 if c else w0)[:, 2]).astype(f32)
                for ch in range(3):
                    np.add.at(ras[ch], (ix, iy, iz), (w * N[:, ch]).astype(np.float64))
    ras_s = np.fft.rfftn(ras.astype(f32), axes=(1, 2, 3))                       # dpsr.py:37
    fx = np.fft.fftfreq(G, d=1 / G)
    fz = np.fft.rfftfreq(G, d=1 / G)
    om = np.stack(np.meshgrid(fx, fx, fz, indexing="ij"), -1)                   # fftfreqs, dpsr_utils.py:25-46
    dis = np.sqrt((om ** 2).sum(-1))
    filt = np.exp(-0.5 * ((sig * 2 * dis / G) ** 2)).astype(f32)               # spec_gaussian_filter, :58-64
    omega = (om.astype(f32) * f32(2 * np.pi)).astype(f32)
    Nf = ras_s * filt[None]
    div = sum((-1j * omega[..., c]) * Nf[c] for c in range(3))                  # dpsr.py:47
    lap = -(omega ** 2).sum(-1)
    Phi = div / (lap + f32(1e-6))
    Phi[0, 0, 0] = 0
    phi = np.fft.irfftn(Phi, s=(G, G, G), axes=(0, 1, 2))                       # dpsr.py:55
    fv = np.zeros(V.shape[0], np.float64)
    for a in (0, 1):
        for b in (0, 1):
            for c in (0, 1):
                ix, iy, iz = (i1 if a else i0)[:, 0], (i1 if b else i0)[:, 1], (i1 if c else i0)[:, 2]
                w = (w1 if a else w0)[:, 0] * (w1 if b else w0)[:, 1] * (w1 if c else w0)[:, 2]
                fv += phi[ix, iy, iz] * w
    phi = phi - fv.mean()
    fv0 = phi[0, 0, 0]
    return (-phi / abs(fv0) * 0.5).astype(f32)


# ----------------------------------------------------------------------------- marching cubes
def marching_cubes_np(phi, iso=0.0):
    """numpy restatement of the marching-cubes contract visible at DG-Mesh's call sites
    (dgmesh/utils/renderer.py:171): phi [G,G,G] -> verts [V,3] in [0,1]^3 (index/(G-1)), faces [F,3].
    `diso` (the reference's third-party implementation) is unavailable: PARITY UNPINNED; the case
    table is the derived one of tools/gen_mc_tables.py, shared with the CUDA kernels but consumed by
    this independent, loop-based emitter.  Small grids only."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(_HERE), "tools"))
    from gen_mc_tables import tables_for_python
    ntri, tri, edge_lo = tables_for_python()
    phi = np.asarray(phi, np.float32)
    G = phi.shape[0]
    inside = phi < np.float32(iso)
    vid = {}
    verts, faces = [], []

    def vertex(i, j, k, e):
        c, axis = int(edge_lo[e]), e // 4
        a = (i + (c & 1), j + ((c >> 1) & 1), k + (c >> 2))
        key = (a, axis)
        if key not in vid:
            b = list(a)
            b[axis] += 1
            p0, p1 = phi[a], phi[tuple(b)]
            t = (np.float32(iso) - p0) / (p1 - p0)
            pos = np.array(a, np.float32)
            pos[axis] += t
            vid[key] = len(verts)
            verts.append(pos / np.float32(G - 1))
        return vid[key]

    for i in range(G - 1):
        for j in range(G - 1):
            for k in range(G - 1):
                cs = 0
                for v in range(8):
                    if inside[i + (v & 1), j + ((v >> 1) & 1), k + (v >> 2)]:
                        cs |= 1 << v
                for t in range(int(ntri[cs])):
                    faces.append([vertex(i, j, k, int(tri[cs, 3 * t + q])) for q in range(3)])
    return (np.array(verts, np.float32).reshape(-1, 3), np.array(faces, np.int64).reshape(-1, 3))


def mesh_topology(verts, faces):
    """(V - E + F, all edges shared by exactly two consistently oriented faces?)"""
    f = np.asarray(faces)
    de = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
    und = np.sort(de, axis=1)
    uniq, cnt = np.unique(und, axis=0, return_counts=True)
    manifold = bool(np.all(cnt == 2))
    # orientation: every directed edge appears exactly once (its reverse belongs to the neighbour)
    dcode = de[:, 0].astype(np.int64) * (f.max() + 1) + de[:, 1]
    oriented = len(np.unique(dcode)) == len(dcode)
    used = len(np.unique(f))
    return used - len(uniq) + len(f), manifold, oriented


# ----------------------------------------------------------------------------- image loss (next-tier row f3)
def _gauss11():
    """The 11-tap window of loss_utils.gaussian(11, 1.5) (dgmesh/utils/loss_utils.py:33-36), float32."""
    import math
    g = np.array([math.exp(-(x - 5) ** 2 / float(2 * 1.5 ** 2)) for x in range(11)], dtype=np.float32)
    return (g / g.sum()).astype(np.float32)


def _blur11(a):
    """Zero-padded 11x11 separable Gaussian blur of [C,H,W] (== F.conv2d(a, outer(g,g), padding=5, groups=C))."""
    g = _gauss11().astype(np.float64)
    C, H, W = a.shape
    p = np.zeros((C, H + 10, W + 10))
    p[:, 5:5 + H, 5:5 + W] = a
    h = sum(g[k] * p[:, :, k:k + W] for k in range(11))
    return sum(g[k] * h[:, k:k + H, :] for k in range(11))


def image_loss_np(img, gt, lam=0.2):
    """(1 - lam) * L1 + lam * (1 - SSIM) as dgmesh/train.py:308-311 composes l1_loss and ssim
    (dgmesh/utils/loss_utils.py:18-19, 39-76), and its gradient w.r.t. the rendered image `img`
    (hand-derived; test_loss pins it against autograd through the reference functions).
    Returns (loss, l1, ssim, dL_dimg)."""
    x, y = np.asarray(img, np.float64), np.asarray(gt, np.float64)
    n = x.size
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    m1, m2 = _blur11(x), _blur11(y)
    e11, e22, e12 = _blur11(x * x), _blur11(y * y), _blur11(x * y)
    s1, s2, s12 = e11 - m1 * m1, e22 - m2 * m2, e12 - m1 * m2
    N1, N2 = 2 * m1 * m2 + C1, 2 * s12 + C2
    D1, D2 = m1 * m1 + m2 * m2 + C1, s1 + s2 + C2
    S = N1 * N2 / (D1 * D2)
    l1, ssim = np.abs(x - y).mean(), S.mean()
    loss = (1 - lam) * l1 + lam * (1 - ssim)
    # d S / d(blurred maps), everything else fixed
    dS_dm1 = (2 * m2 * N2 - 2 * m2 * N1) / (D1 * D2) - S * (2 * m1 / D1 - 2 * m1 / D2)
    dS_de11 = -S / D2
    dS_de12 = 2 * N1 / (D1 * D2)
    # the window is symmetric: the adjoint of the blur is the blur
    dssim_dx = (_blur11(dS_dm1) + 2 * x * _blur11(dS_de11) + y * _blur11(dS_de12)) / n
    grad = (1 - lam) * np.sign(x - y) / n - lam * dssim_dx
    return loss, l1, ssim, grad.astype(np.float32)
