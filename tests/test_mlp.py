"""Deformation / appearance MLPs: the tcgen05 GEMM chain vs (a) a bf16-rounding restatement of the
reference network (same rounding points as the kernels: bf16 operands, fp32 accumulate) and (b) the
reference's own fp32 PyTorch modules (oracle/_ref/refpy/time_utils.py).  Tolerances are stated
relative to each tensor's scale: 2e-3 (L2) / 1e-2 (max) against (a) forward, 2e-2 against (a) backward (the backward
also rounds dZ to bf16), 5e-2 against the fp32 reference (the bf16 gap, reported)."""
import importlib

import pytest
import torch

import util

ref = util.load_reference_pymodules()
needs_ref = pytest.mark.skipif(ref is None, reason="oracle/_ref/refpy missing")


def bf(x):
    """round to bf16, straight-through gradient (what mixed-precision backward assumes)"""
    return x + (x.bfloat16().float() - x).detach()


def emulate(net, x, t):
    """bf16-operand / fp32-accumulate restatement of time_utils.py:178-204 for any of the four nets."""
    def pe(v, L):
        out = [v]
        for k in range(L):
            out += [torch.sin(v * 2.0 ** k), torch.cos(v * 2.0 ** k)]
        return torch.cat(out, -1)

    def lin(m, h):
        return bf(h) @ bf(m.weight).t() + m.bias

    x_emb = pe(x, 10)
    if net.is_blender:
        t_emb = lin(net.timenet[2], torch.relu(lin(net.timenet[0], pe(t, 6))))
    else:
        t_emb = pe(t, 10)
    e = torch.cat([x_emb, t_emb], -1)
    h = e
    for i, l in enumerate(net.linear):
        h = torch.relu(lin(l, h))
        if i == 4:
            h = torch.cat([e, h], -1)
    return h


def heads(net, h):
    name = type(net).__name__
    if name == "AppearanceNetwork":
        return (torch.sigmoid(bf(h) @ bf(net.color_warp[0].weight).t() + net.color_warp[0].bias),)
    outs = []
    for n in ("gaussian_warp", "gaussian_rotation", "gaussian_scaling", "gaussian_normal"):
        if hasattr(net, n):
            m = getattr(net, n)
            outs.append(bf(h) @ bf(m.weight).t() + m.bias)
    return tuple(outs)


def inputs(n, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 3, generator=g) * 0.6
    t = torch.full((n, 1), 0.37)
    t[: n // 3] = torch.rand(n // 3, 1, generator=g)      # per-point times are allowed by the interface
    return x, t


@needs_ref
def test_emulation_is_close_to_fp32_reference_cpu():
    torch.manual_seed(0)
    net = ref.time_utils.DeformNetworkNormal(is_blender=True)
    x, t = inputs(500, 1)
    with torch.no_grad():
        a = torch.cat(heads(net, emulate(net, x, t)), -1)
        b = torch.cat(net(x, t), -1)
    assert util.rel_err(a, b) < 3e-2


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("cls,blender,n", [("DeformNetworkNormal", True, 3000), ("DeformNetworkNormal", False, 1000),
                                           ("DeformNetwork", True, 517), ("DeformNetworkNormalSep", True, 2048),
                                           ("AppearanceNetwork", True, 1500), ("DeformNetworkNormal", True, 100_000)])
def test_mlp_forward_backward(cls, blender, n):
    tu = importlib.import_module("utils.time_utils")
    torch.manual_seed(1)
    kw = dict(is_blender=blender)
    mine = getattr(tu, cls)(**kw).cuda()
    theirs = getattr(ref.time_utils, cls)(**kw).cuda()
    theirs.load_state_dict(mine.state_dict())
    if cls == "DeformNetworkNormalSep":      # zero-initialised head: give it something to compute
        for m in (mine, theirs):
            torch.manual_seed(2)
            torch.nn.init.normal_(m.gaussian_normal.weight, std=0.05)
    x, t = inputs(n, 3)
    x, t = x.cuda(), t.cuda()
    xa = x.clone().requires_grad_(True)
    out = mine(xa, t)
    out = out if isinstance(out, tuple) else (out,)
    ya = torch.cat(out, -1)
    xb = x.clone().requires_grad_(True)
    yb = torch.cat(heads(theirs, emulate(theirs, xb, t)), -1)
    with torch.no_grad():
        yc = theirs(x, t)
        yc = torch.cat(yc if isinstance(yc, tuple) else (yc,), -1)
    assert ya.shape == yb.shape == yc.shape
    # max-norm: one bf16 rounding flip (2^-8) in a late activation is visible in a single output;
    # the L2 bound is the tight one
    assert util.rel_err(ya, yb) < 1e-2, "vs bf16 restatement (max)"
    assert util.rel_l2(ya, yb) < 2e-3, "vs bf16 restatement (L2)"
    assert util.rel_err(ya, yc) < 5e-2, "vs fp32 reference"
    g = torch.randn(ya.shape, generator=torch.Generator().manual_seed(4)).cuda()
    ya.backward(g)
    yb.backward(g)
    # gradients: relative L2 per tensor.  dx passes through d pe(x)/dx with frequencies up to 2^9, which
    # multiplies the bf16 rounding of the upstream gradient by up to 512: looser bound.
    errs = {"dx": util.rel_l2(xa.grad, xb.grad)}
    pa, pb = dict(mine.named_parameters()), dict(theirs.named_parameters())
    for k in pa:
        assert pa[k].grad is not None and pa[k].grad.shape == pb[k].grad.shape, k
        errs[k] = util.rel_l2(pa[k].grad, pb[k].grad)
    print({k: round(v, 4) for k, v in errs.items()})
    assert errs.pop("dx") < 6e-2
    bad = {k: v for k, v in errs.items() if v > 2e-2}
    assert not bad, bad


@pytest.mark.gpu
def test_mlp_inference_mode_and_errors():
    tu = importlib.import_module("utils.time_utils")
    torch.manual_seed(0)
    net = tu.DeformNetworkNormal(is_blender=True).cuda()
    x, t = inputs(777, 5)
    with torch.no_grad():
        a = torch.cat(net(x.cuda(), t.cuda()), -1)     # ping-pong activation buffers, no stash
    b = torch.cat(net(x.cuda(), t.cuda()), -1)          # training path
    assert torch.equal(a, b.detach())
    with pytest.raises(NotImplementedError):
        tu.DeformNetworkNormal(is_6dof=True)
    with pytest.raises(ValueError):
        net(x, t)                                        # CPU tensors: no fallback
