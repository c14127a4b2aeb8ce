"""Stand-in for the third-party `diso` package as DG-Mesh uses it
(`from diso import DiffDMC, DiffMC`, dgmesh/scene/gaussian_model_dpsr_dynamic_anchor.py:40,84;
`gaussians.diffmc(psr, deform=None, isovalue=0.0)`, dgmesh/utils/renderer.py:171).

`diso` is not vendored with the reference and cannot be installed offline, so this follows the
CONTRACT visible at those call sites -- grid [G,G,G] fp32 -> (verts [V,3] fp32 in [0,1]^3, faces
[F,3] integer), differentiable w.r.t. the grid -- with our own sm_100a marching cubes.  Vertex and
face ORDER are this implementation's (parity with diso is unpinned)."""
import os
import sys

import torch
import torch.nn as nn

_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _root not in sys.path:
    sys.path.insert(0, _root)
import _dgm_lib  # noqa: E402


class _MCFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, grid, isovalue, normalize):
        if grid.dim() != 3 or not (grid.shape[0] == grid.shape[1] == grid.shape[2]):
            raise ValueError("DiffMC: cubic grid [G, G, G] expected")
        if not grid.is_cuda:
            raise ValueError("DiffMC: CUDA tensor required (no CPU fallback)")
        lib = _dgm_lib.lib()
        g = grid.contiguous().float()
        G = g.shape[0]
        nbytes = _dgm_lib.c_size_t()
        _dgm_lib.check(lib.dgmc_workspace_size(G, nbytes), "dgmc_workspace_size")
        ws = torch.empty((nbytes.value,), dtype=torch.uint8, device=g.device)
        totals = torch.empty((2,), dtype=torch.int32, device=g.device)
        st = _dgm_lib.stream_ptr()
        _dgm_lib.check(lib.dgmc_count(G, g.data_ptr(), float(isovalue), ws.data_ptr(), nbytes.value,
                                      totals.data_ptr(), st), "dgmc_count")
        V, F = (int(x) for x in totals.cpu())  # the one host sync: output sizes are data dependent
        verts = torch.empty((V, 3), dtype=torch.float32, device=g.device)
        faces = torch.empty((F, 3), dtype=torch.int32, device=g.device)
        if V:
            _dgm_lib.check(lib.dgmc_emit(G, g.data_ptr(), float(isovalue), ws.data_ptr(), nbytes.value,
                                         verts.data_ptr(), faces.data_ptr(), st), "dgmc_emit")
        if not normalize:
            verts = verts * float(G - 1)
        ctx.save_for_backward(g, ws)
        ctx.iso, ctx.normalize, ctx.nbytes = float(isovalue), normalize, nbytes.value
        ctx.mark_non_differentiable(faces)
        return verts, faces

    @staticmethod
    def backward(ctx, dverts, _):
        g, ws = ctx.saved_tensors
        G = g.shape[0]
        dphi = torch.empty_like(g)
        dv = dverts.contiguous().float()
        if not ctx.normalize:
            dv = dv * float(G - 1)
        if dv.shape[0] == 0:
            return torch.zeros_like(g), None, None
        _dgm_lib.check(_dgm_lib.lib().dgmc_backward(G, g.data_ptr(), ctx.iso, ws.data_ptr(), ctx.nbytes,
                                                    dv.data_ptr(), dphi.data_ptr(), _dgm_lib.stream_ptr()),
                       "dgmc_backward")
        return dphi, None, None


class DiffMC(nn.Module):
    def __init__(self, dtype=torch.float32):
        super().__init__()
        if dtype != torch.float32:
            raise NotImplementedError("DiffMC (B200): float32 only (what DG-Mesh instantiates)")

    def forward(self, grid, deform=None, isovalue=0.0, normalize=True):
        if deform is not None:
            raise NotImplementedError("DiffMC (B200): deform=None only (what DG-Mesh passes)")
        verts, faces = _MCFunction.apply(grid, float(isovalue), bool(normalize))
        return verts, faces.long()


class DiffDMC(nn.Module):
    """Imported but never instantiated by DG-Mesh (gaussian_model_dpsr_dynamic_anchor.py:40)."""

    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError("DiffDMC is not part of the DG-Mesh hot path")
