#!/usr/bin/env python
"""Multi-GPU check, launched with torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/dist_gpu_check.py

Batch-sharded objective/gradient (NCCL all-reduce of [grad; obj]) and row-sharded BFGS
(all-gather / all-reduce of the product vectors) must reproduce the single-device oracle results.
"""
import os
import sys

import numpy
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for sub in ("optimal-py-learning_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, sub))

import mlp_oracle as mo            # noqa: E402
import optimize_oracle as oo       # noqa: E402
import learning_b200 as L          # noqa: E402
from util import rel_err           # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    rng = numpy.random.default_rng(0)
    shape, transfers, eid = (40, 33, 7), [(mo.SOFTPLUS, 0), (mo.SOFTMAX, 0)], mo.CROSS_ENTROPY
    x, y = mo.random_classification(1001, 40, 7, rng)
    w = mo.random_params(shape, rng)
    want_obj, want_grad = mo.objective_gradient(w, shape, transfers, eid, x, y)

    model = L.MLP(shape, transfers=L.SoftmaxTransfer(), error_func=L.CrossEntropyError(),
                  optimizer=L.optimize.BFGS(shard_rows=True))
    model.logging = False
    obj, grad = model._get_obj_jac(w.copy(), x, y)          # shards rows over the ranks
    assert model._workspace.sharded
    assert rel_err(obj, want_obj) <= 1e-10 and rel_err(grad, want_grad) <= 1e-10, (rel_err(obj, want_obj), rel_err(grad, want_grad))
    assert rel_err(model._get_obj(w.copy(), x, y), want_obj) <= 1e-10

    # 12 BFGS steps with H row-sharded vs the oracle's single-process BFGS
    bias, rest = w[:33].copy(), w[33:]
    model._bias_vec, model._weight_matrices = bias, [rest[:40 * 33].reshape(40, 33).copy(), rest[40 * 33:].reshape(33, 7).copy()]
    ref = oo.BFGS()
    cur = w.copy()
    f = lambda v: mo.objective_gradient(v, shape, transfers, eid, x, y)
    fo = lambda v: mo.objective(v, shape, transfers, eid, x, y)
    for step in range(12):
        err = model.train_step(x, y)
        want_err, cur = ref.step(f, fo, cur)
        got = numpy.hstack([model._bias_vec] + [m.ravel() for m in model._weight_matrices])
        assert rel_err(err, want_err) <= 1e-8 and rel_err(got, cur) <= 1e-8, (step, rel_err(err, want_err), rel_err(got, cur))
    h_local = model._optimizer.inverse_hessian()
    begin, end, _ = model._optimizer._row_range(w.size)
    assert rel_err(h_local, ref.h[begin:end]) <= 1e-7, rel_err(h_local, ref.h[begin:end])
    # the peer-memory all-reduce kernel on its own: equal to NCCL's sum, bit-identical on every rank, over several rounds
    # (double buffering / monotonic flags), odd and even lengths
    from learning_b200 import _collective
    used_peer = getattr(model._workspace, "peer_allreduce", None) is not None
    for count in (203521, 7):
        ar = _collective.make(count)
        if ar is None:
            assert not used_peer
            break
        g = torch.Generator(device="cuda").manual_seed(100 + rank)
        for rnd in range(5):
            mine = torch.randn(count, dtype=torch.float64, device="cuda", generator=g) * (10.0 ** (rnd - 2))
            ar.next_slot().copy_(mine)
            out = torch.empty(count, dtype=torch.float64, device="cuda")
            ar.reduce(out)
            want = mine.clone()
            dist.all_reduce(want)
            assert float((out - want).abs().max()) <= 1e-13 * float(want.abs().max()), (count, rnd)
            gathered = [torch.empty_like(out) for _ in range(world)]
            dist.all_gather(gathered, out)
            assert all(torch.equal(gathered[0], t) for t in gathered), "ranks disagree bitwise"
    dist.barrier()
    if rank == 0:
        print("dist_gpu_check ok: world=%d, sharded obj/grad + 12 row-sharded BFGS steps match the oracle; gradient all-reduce: %s"
              % (world, "peer-memory kernel (%s)" % model._workspace.peer_allreduce.mode if used_peer else "NCCL"))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
