#!/usr/bin/env python
"""Run under torchrun on >= 2 GPUs: the library's own NVSwitch all-reduce (dp.NvlsFlatGrad, csrc/nvls.cu) against
torch.distributed.all_reduce (NCCL) -- equality of the result and time per call at the path's sizes.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/nvls_test.py
Rank 0 prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dg-mesh_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import dp  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)


def timed(fn, steps=20, warmup=5):
    for _ in range(warmup):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms)


out = {"world": world}
# a first buffer that performs exactly ONE all-reduce (epoch 1): the next buffer also starts at epoch 1 and must
# not see a stale "go" value from this one
_p0 = [torch.nn.Parameter(torch.zeros(4096, device=dev))]
_fg0 = dp.NvlsFlatGrad(_p0)
_fg0.flat.fill_(float(rank + 1))
_r0 = _fg0.allreduce(average=False).clone()
torch.cuda.synchronize()
assert float((_r0 - world * (world + 1) / 2).abs().max()) == 0.0, "single-call buffer"
for name, n in (("c2_gaussian_grads", 5_900_000), ("c4_gaussian_plus_mlp_grads", 15_000_000), ("small", 1000)):
    torch.manual_seed(100 + rank)
    params = [torch.nn.Parameter(torch.zeros(n // 2, device=dev)), torch.nn.Parameter(torch.zeros(n - n // 2, device=dev))]
    fg = dp.NvlsFlatGrad(params)
    src = torch.randn(fg.flat.numel(), device=dev)
    want = src.clone()
    dist.all_reduce(want)
    fg.flat.copy_(src)
    torch.cuda.synchronize()
    dist.barrier()
    got = fg.allreduce(average=False).clone()
    torch.cuda.synchronize()
    err = float((got - want).abs().max() / want.abs().max())
    # averaged variant
    fg.flat.copy_(src)
    got2 = fg.allreduce(average=True).clone()
    err2 = float((got2 - want / world).abs().max() / want.abs().max())
    nccl_buf = src.clone()
    row = {"floats": n, "bytes": 4 * n, "max_rel_err_sum": err, "max_rel_err_mean": err2,
           "nvls_ms": timed(lambda: fg.allreduce(average=False)), "nccl_ms": timed(lambda: dist.all_reduce(nccl_buf))}
    row["speedup"] = row["nccl_ms"] / row["nvls_ms"]
    out[name] = row
    del fg
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
