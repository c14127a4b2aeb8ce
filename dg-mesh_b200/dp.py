"""Data parallelism over training frames -- new in this implementation: the reference
is single-process (dgmesh/train.py renders one frame per iteration on cuda:0).

A frame's forward/backward depends only on the replicated parameters, its camera and its time, so
frame k of a batch goes to rank k % world and the only data-path exchange is ONE all-reduce (NCCL over
NVLink on the B200 box, gloo in the CPU tests) of a single flat fp32 buffer holding every gradient:
canonical-Gaussian parameters followed by the MLP parameters.  Batch of one frame per step -> no
collective at all.  Densification statistics need the same reduction so all ranks prune identically
(gaussian_model_dpsr_dynamic_anchor.py:679-682)."""
import torch
import torch.distributed as dist


def shard_frames(n_frames, rank, world):
    """Frame indices of this rank (round-robin, like the one-frame-per-GPU layout of the 8xB200 box)."""
    return [k for k in range(n_frames) if k % world == rank]


class FlatGrad:
    """Gradients of `params` as views of ONE flat buffer, so the exchange is a single collective."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatGrad: no parameter requires grad")
        dev, dt = self.params[0].device, torch.float32
        self.flat = torch.zeros(sum(p.numel() for p in self.params), device=dev, dtype=dt)
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n

    def zero(self):
        self.flat.zero_()

    def allreduce(self, average=True, group=None):
        """Sum (or mean) of the gradients over the ranks; no-op for a single process."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            if average:
                self.flat.div_(dist.get_world_size(group))
        return self.flat


def sync_densification_stats(xyz_gradient_accum, denom, max_radii2D, group=None):
    """Sum the accumulated view-space gradient norms / visit counts and take the max of the screen radii
    over ranks (two small collectives, off the per-step critical path: every densification interval)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    buf = torch.cat([xyz_gradient_accum.reshape(-1), denom.reshape(-1)])
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    n = xyz_gradient_accum.numel()
    xyz_gradient_accum.copy_(buf[:n].view_as(xyz_gradient_accum))
    denom.copy_(buf[n:].view_as(denom))
    dist.all_reduce(max_radii2D, op=dist.ReduceOp.MAX, group=group)


def broadcast_parameters(params, src=0, group=None):
    """Topology-changing steps (densify / prune / anchor) run on rank `src`; everyone else receives."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for p in params:
        dist.broadcast(p.data, src=src, group=group)
