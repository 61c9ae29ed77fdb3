#!/usr/bin/env python
"""Where a weak-scaling step's time goes at N GPUs: evaluation alone, evaluation + all-reduce, all-reduce alone, per rank.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29530 tools/weak_diag.py"""
import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "optimal-py-learning_b200"))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import learning_b200 as L
from learning_b200 import _native as nat, _device as dev
from bench import synthetic, ROWS, SHAPE

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local), init_method=None if "RANK" in os.environ else "tcp://127.0.0.1:29531",
                        rank=int(os.environ.get("RANK", "0")), world_size=int(os.environ.get("WORLD_SIZE", "1")))
rank, world = dist.get_rank(), dist.get_world_size()
x, y, w = synthetic(ROWS, 100 + rank)
_, _, w = synthetic(8, 0)
n = w.size
model = L.MLP(SHAPE, transfers=L.SoftmaxTransfer(), error_func=L.CrossEntropyError(), optimizer=L.optimize.LBFGS())
xd, yd = dev.to_device(x), dev.to_device(y)
ws = model._make_resident(xd, yd)
wd = dev.to_device(w)
out = torch.empty(n + 2, dtype=torch.float64, device="cuda")
peer = model._peer_allreduce(ws, n + 2) if world > 1 else None


def ev(dst):
    nat.check(nat.lib.opl_mlp_eval_device(ws.pointer, dev.ptr(wd), 0.0, None, 1, dev.ptr(dst)))


def timed(fn, k=40):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(k):
        fn()
    e1.record()
    host = (time.perf_counter() - t0) / k * 1e3
    torch.cuda.synchronize()
    mine = e0.elapsed_time(e1) / k
    t = torch.tensor([mine], dtype=torch.float64, device="cuda")
    all_t = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(all_t, t)
    return [round(float(v.item()), 4) for v in all_t], round(host, 4)


res = {}
res["eval only (into a plain buffer)"] = timed(lambda: ev(out))
cycles = int(0.028e-3 * 1.9e9)
res["eval + 28 us sleep kernel"] = timed(lambda: (ev(out), torch.cuda._sleep(cycles)))
res["28 us sleep kernel only"] = timed(lambda: torch.cuda._sleep(cycles))
scratch = torch.empty(n + 2, dtype=torch.float64, device="cuda")
res["eval + device copy of the vector"] = timed(lambda: (ev(out), scratch.copy_(out)))
if peer is not None:
    res["eval only (into the symmetric slot)"] = timed(lambda: ev(peer.next_slot()))
    res["eval + peer all-reduce"] = timed(lambda: (ev(peer.next_slot()), peer.reduce(out)))
    res["peer all-reduce only"] = timed(lambda: (peer.next_slot(), peer.reduce(out)))
    res["eval + NCCL all-reduce"] = timed(lambda: (ev(out), dist.all_reduce(out)))
# step-to-step jitter of the evaluation: one event per step
for _ in range(5):
    ev(out)
torch.cuda.synchronize()
marks = [torch.cuda.Event(enable_timing=True) for _ in range(201)]
marks[0].record()
for i in range(200):
    ev(out)
    marks[i + 1].record()
torch.cuda.synchronize()
import numpy
dts = numpy.array([marks[i].elapsed_time(marks[i + 1]) for i in range(200)])
if rank == 0:
    print("per-step evaluation time over 200 steps: mean %.4f  std %.4f  min %.4f  p5 %.4f  p50 %.4f  p95 %.4f  max %.4f ms"
          % (dts.mean(), dts.std(), dts.min(), numpy.percentile(dts, 5), numpy.percentile(dts, 50), numpy.percentile(dts, 95), dts.max()), flush=True)
    for k, (per_rank, host) in res.items():
        print("%-40s per-rank ms %s   host issue ms/step %.4f" % (k, per_rank, host), flush=True)
dist.destroy_process_group()
