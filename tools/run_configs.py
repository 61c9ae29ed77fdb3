#!/usr/bin/env python
"""Run every BASELINE.json config for a few optimizer steps through the public plugin API (Model.train_step with the
reference's default optimizers) on one GPU and print one JSON line per config: wall time per step, objective+gradient
evaluations per step, evaluations per second, objective trajectory.  Config 4 (H row-sharded over 8 GPUs) runs when the
script is launched by torchrun with one rank per GPU.

    python tools/run_configs.py [1 2 3 5 ...]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_configs.py     # config 4
"""
import json
import os
import sys
import time

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "optimal-py-learning_b200"))
import torch  # noqa: E402
import learning_b200 as L  # noqa: E402
from learning_b200 import _device as dev  # noqa: E402


def synthetic(rows, d_in, d_out, seed, classification=True):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.rand(rows, d_in, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    if classification:
        y = torch.zeros(rows, d_out, dtype=torch.float64, device="cuda")
        idx = torch.randint(0, d_out, (rows,), device="cuda", generator=g)
        y[torch.arange(rows, device="cuda"), idx] = 1.0
    else:
        y = torch.rand(rows, d_out, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    return x, y


def run(name, model, x, y, steps):
    numpy.random.seed(0)
    model.logging = False
    model.reset()
    objs, times, evals = [], [], []
    torch.cuda.synchronize()
    with model.resident(x, y):                       # what Model.train does around its train_steps
        for _ in range(steps):
            before = model.obj_jac_evaluations
            t0 = time.perf_counter()
            objs.append(float(model.train_step(x, y)))
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
            evals.append(model.obj_jac_evaluations - before)   # evaluations of THIS step (line searches differ per step)
    first = 2 if len(times) > 3 else 0               # steady state: steps after the first two (allocations, first upload)
    out = {"config": name, "steps": steps, "objective": [round(o, 6) for o in objs],
           "s_per_step": [round(t, 5) for t in times], "obj_grad_evaluations_per_step": evals,
           "obj_grad_evaluations": sum(evals),
           "evals_per_s_steady": round(sum(evals[first:]) / sum(times[first:]), 2) if sum(times[first:]) > 0 else None,
           "gpu_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}
    assert all(numpy.isfinite(objs)), objs
    assert objs[-1] <= objs[0] * (1 + 1e-9), ("objective did not decrease", objs)
    print(json.dumps(out), flush=True)


def run_config4():
    """Config 4 under torchrun (one rank per GPU): (1024,256,256,32), BFGS with n = 336128 weights, H (904 GB FP64)
    row-sharded over the ranks, 65536 patterns sharded by rows with an NCCL all-reduce of [gradient ; objective]."""
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    world, rank = dist.get_world_size(), dist.get_rank()
    rows = 65536 // world
    x, y = synthetic(rows, 1024, 32, 100 + rank)
    numpy.random.seed(0)
    m = L.MLP((1024, 256, 256, 32), transfers=L.SoftmaxTransfer(), error_func=L.CrossEntropyError(),
              optimizer=L.optimize.BFGS(shard_rows=True))
    m.logging = False
    objs, times = [], []
    for _ in range(6):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        objs.append(float(m.train_step(x, y)))
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    free, total = torch.cuda.mem_get_info()
    if rank == 0:
        n = 256 + 1024 * 256 + 256 * 256 + 256 * 32
        print(json.dumps({"config": "4: (1024,256,256,32) 65536 patterns over %d GPUs, BFGS n=%d, H row-sharded (%.1f GB per GPU)"
                                    % (world, n, n * n * 8 / world / 2 ** 30),
                          "objective": [round(o, 6) for o in objs], "s_per_step": [round(t, 4) for t in times],
                          "obj_grad_evaluations": m.obj_jac_evaluations,
                          "gpu_mem_used_gb_rank0": round((total - free) / 2 ** 30, 1)}), flush=True)
        assert all(numpy.isfinite(objs)) and objs[-1] < objs[0], objs
    dist.destroy_process_group()


def main():
    if "WORLD_SIZE" in os.environ and int(os.environ["WORLD_SIZE"]) > 1:
        return run_config4()
    which = [int(a) for a in sys.argv[1:]] or [1, 2, 3, 5]
    opt = L.optimize
    if 1 in which:   # README iris-sized MLP (4,2,3), BFGS + Wolfe (runs on the CPU in the reference)
        rng = numpy.random.default_rng(0)
        x = rng.random((90, 4)) * 2 - 1
        y = numpy.zeros((90, 3))
        y[numpy.arange(90), rng.integers(0, 3, 90)] = 1.0
        m = L.MLP((4, 2, 3), transfers=L.SoftmaxTransfer(), error_func=L.CrossEntropyError(), optimizer=opt.BFGS())
        run("1: (4,2,3) 90 patterns, BFGS + Wolfe", m, x, y, 30)
    if 2 in which:   # 60k x 784 MLP (784,256,10), L-BFGS full batch
        x, y = synthetic(60000, 784, 10, 1)
        m = L.MLP((784, 256, 10), transfers=L.SoftmaxTransfer(), error_func=L.CrossEntropyError(), optimizer=opt.LBFGS())
        run("2: (784,256,10) 60000 patterns, L-BFGS, FP64 parity mode", m, x, y, 12)
        del m, x, y
        torch.cuda.empty_cache()
    if 3 in which:   # (512,128,128,10), BFGS with n = 83328 (H = 55.5 GB FP64) on one GPU
        x, y = synthetic(65536, 512, 10, 2)
        m = L.MLP((512, 128, 128, 10), transfers=L.SoftmaxTransfer(), error_func=L.CrossEntropyError(), optimizer=opt.BFGS())
        run("3: (512,128,128,10) 65536 patterns, BFGS n=83328 (H 55.5 GB), FP64 parity mode", m, x, y, 8)
        del m, x, y
        torch.cuda.empty_cache()
    if 5 in which:   # deep MLP (256,1024x4,10), 1M patterns, steepest descent + momentum, TF32 fast mode
        rows = int(os.environ.get("OPL_CFG5_ROWS", 1 << 20))
        x, y = synthetic(rows, 256, 10, 3, classification=False)
        m = L.MLP((256, 1024, 1024, 1024, 1024, 10), optimizer=opt.SteepestDescentMomentum(), precision="tf32")
        run("5: (256,1024x4,10) %d patterns, steepest descent + momentum, TF32 fast mode" % rows, m, x, y, 6)
        del m, x, y
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
