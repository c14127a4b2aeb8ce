"""Frame data parallelism on CPU (gloo, world_size 2): the all-reduced flat gradient equals the sum of
the per-frame gradients a single process computes; statistics reductions; frame sharding."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import dp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _frame_loss(params, k):
    a, b = params
    return ((a * (k + 1)).sin() * b[k % b.shape[0]]).sum() + (b ** 2).sum() * 0.1 * (k + 1)


def _worker(rank, world, port, n_frames, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    params = [torch.randn(50, 3, requires_grad=True), torch.randn(7, 3, requires_grad=True)]
    fg = dp.FlatGrad(params)
    fg.zero()
    for k in dp.shard_frames(n_frames, rank, world):
        _frame_loss(params, k).backward()
    flat = fg.allreduce(average=False).clone()
    acc, den, rad = torch.full((50, 1), float(rank + 1)), torch.full((50, 1), 1.0), torch.arange(50.) * (rank + 1)
    dp.sync_densification_stats(acc, den, rad)
    if rank == 0:
        torch.save((flat, acc, den, rad), out)      # a file, not a multiprocessing queue (spawn + pipes proved flaky)
    dist.destroy_process_group()


def test_allreduced_gradient_equals_sum_of_frame_gradients(tmp_path):
    n_frames, world = 8, 2
    ctx = mp.get_context("spawn")
    out = str(tmp_path / "rank0.pt")
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    flat, acc, den, rad = torch.load(out)
    torch.manual_seed(0)
    params = [torch.randn(50, 3, requires_grad=True), torch.randn(7, 3, requires_grad=True)]
    fg = dp.FlatGrad(params)
    for k in range(n_frames):
        _frame_loss(params, k).backward()
    assert torch.allclose(flat, fg.flat, rtol=1e-5, atol=1e-6)
    assert torch.all(acc == 3.0) and torch.all(den == 2.0) and torch.equal(rad, torch.arange(50.) * 2)


def test_shard_frames_and_single_process_noop():
    assert dp.shard_frames(8, 1, 4) == [1, 5] and sum(len(dp.shard_frames(8, r, 8)) for r in range(8)) == 8
    p = [torch.ones(3, requires_grad=True)]
    fg = dp.FlatGrad(p)
    (p[0] * 2).sum().backward()
    assert torch.equal(fg.allreduce(), torch.full((3,), 2.0))     # not initialised: no collective


def test_flatgrad_survives_zero_grad_set_to_none_and_densification():
    """the reference loop calls optimizer.zero_grad(set_to_none=True) and densification replaces
    the Parameter objects; the flat buffer must still hold the true gradients (never a stale / zero one)."""
    a, b = torch.nn.Parameter(torch.randn(5, 3)), torch.nn.Parameter(torch.randn(4))
    fg = dp.FlatGrad([a, b])
    opt = torch.optim.SGD([a, b], lr=0.1)
    (a.sum() * 2 + (b ** 2).sum()).backward()
    assert torch.equal(fg.allreduce(), torch.cat([torch.full((15,), 2.0), 2 * b.detach()]))
    opt.zero_grad(set_to_none=True)                     # drops the views
    assert a.grad is None
    (a.sum() * 3).backward()                            # autograd allocates a fresh a.grad; b gets none
    flat = fg.allreduce()
    assert torch.equal(flat, torch.cat([torch.full((15,), 3.0), torch.zeros(4)]))
    assert a.grad.data_ptr() == fg.flat.data_ptr()      # re-bound: the next backward accumulates in place
    a2 = torch.nn.Parameter(torch.randn(9, 3))          # densification: a new, larger parameter
    fg.params[0] = a2
    (a2.sum()).backward()
    import pytest
    with pytest.raises(RuntimeError, match="rebuild"):
        fg.allreduce()
    fg.rebuild([a2, b])
    fg.zero()
    (a2.sum()).backward()
    assert float(fg.allreduce().sum()) == 27.0


def _exchange_worker(rank, world, port, out, want_nvls):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    if want_nvls:
        os.environ["DGMESH_B200_EXCHANGE"] = "nvls"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    ex = bench.Exchange(1001, torch.device("cpu"), world)      # no symmetric memory on the CPU: must agree on NCCL
    ex.flat.fill_(rank + 1.0)
    ex.allreduce()
    if rank == 0:
        torch.save((ex.flat.clone(), ex.kind, ex.handle is None), out)
    dist.destroy_process_group()


def test_bench_exchange_falls_back_consistently_on_every_rank(tmp_path):
    """bench.Exchange: when the NVSwitch path cannot be set up (here: CPU tensors), every rank ends on the library
    all-reduce -- the choice is agreed across ranks, so no rank can wait in a kernel the others never launch."""
    world = 2
    ctx = mp.get_context("spawn")
    for want_nvls in (False, True):
        out = str(tmp_path / f"ex{int(want_nvls)}.pt")
        port = _free_port()
        procs = [ctx.Process(target=_exchange_worker, args=(r, world, port, out, want_nvls)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(timeout=240)
            assert p.exitcode == 0
        flat, kind, no_handle = torch.load(out)
        assert no_handle and kind.startswith("nccl") and torch.all(flat == 3.0)
        assert ("unavailable" in kind) == want_nvls
