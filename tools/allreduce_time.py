#!/usr/bin/env python
"""Time of one all-reduce of `count` doubles: the peer-memory pull kernel (csrc/collective.cu) next to NCCL.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29520 tools/allreduce_time.py [count]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "optimal-py-learning_b200"))
import torch, torch.distributed as dist
from learning_b200 import _collective

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
for count in [int(v) for v in sys.argv[1:]] or [203521, 83329, 336129]:
    ar = _collective.make(count)
    x = torch.randn(count, dtype=torch.float64, device="cuda")
    out = torch.empty_like(x)
    res = {}
    for name in ("peer", "nccl"):
        def step():
            if name == "peer":
                ar.next_slot().copy_(x)
                ar.reduce(out)
            else:
                out.copy_(x)
                dist.all_reduce(out)
        for _ in range(5): step()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): step()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 50], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[name] = float(t.item())
    # the copy alone
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): out.copy_(x)
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print("world %d count %d: peer kernel (%s) %.1f us, NCCL %.1f us (both incl. a %.1f us device copy of the vector)"
              % (world, count, ar.mode, res["peer"] * 1e3, res["nccl"] * 1e3, e0.elapsed_time(e1) / 50 * 1e3), flush=True)
dist.destroy_process_group()
