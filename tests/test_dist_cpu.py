"""CPU, world_size 2 (gloo): the host-side sharding logic and the exchange pattern of the N>1 path.

The kernels are replaced by the oracle here (tests may use it); what is under test is the product's
partitioning code (MLP._shard, BFGS._row_range), the global-N scaling contract of the per-rank partial
sums, and the collectives the product issues (all-reduce of [grad; obj], all-gather + all-reduce of the
BFGS product vectors).
"""
import os
import socket
import sys

import numpy
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    for sub in ("optimal-py-learning_b200", "oracle", "tests"):
        sys.path.insert(0, os.path.join(ROOT, sub))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import mlp_oracle as mo
        import learning_b200 as L

        # ---- batch-sharded objective / gradient ------------------------------------------------
        rng = numpy.random.default_rng(0)                      # same data on every rank
        for shape, transfers, eid, rows in [((6, 5, 3), [(mo.SOFTPLUS, 0), (mo.SOFTMAX, 0)], mo.CROSS_ENTROPY, 37),
                                            ((4, 7, 2), [(mo.TANH, 0), (mo.LINEAR, 0)], mo.MSE, 10)]:
            gen = mo.random_classification if eid == mo.CROSS_ENTROPY else mo.random_regression
            x, y = gen(rows, shape[0], shape[-1], rng)
            w = mo.random_params(shape, rng)
            model = L.MLP(shape, transfers=L.SoftmaxTransfer() if eid else L.LinearTransfer(),
                          optimizer=L.optimize.SteepestDescent())
            first, count, sharded = model._shard(rows)
            assert sharded and count > 0
            # what a rank's kernels produce: sums over ITS rows scaled by the GLOBAL N
            obj_s, grad_s = mo.objective_gradient(w, shape, transfers, eid, x[first:first + count], y[first:first + count])
            partial = torch.from_numpy(numpy.append(grad_s, obj_s) * (count / float(rows)))
            dist.all_reduce(partial)                            # the product's collective (mlp.py _probe)
            obj, grad = mo.objective_gradient(w, shape, transfers, eid, x, y)
            got = partial.numpy()
            assert abs(got[-1] - obj) <= 1e-12 * abs(obj)
            assert numpy.linalg.norm(got[:-1] - grad) <= 1e-12 * numpy.linalg.norm(grad)
            # shards tile the batch exactly
            spans = [None] * world
            dist.all_gather_object(spans, (first, count))
            assert sorted(spans)[0][0] == 0 and sum(c for _, c in spans) == rows
            assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(world - 1))

        # ---- row-sharded BFGS exchange ------------------------------------------------------------
        n = 23
        h = rng.standard_normal((n, n))
        yv, g = rng.standard_normal(n), rng.standard_normal(n)
        opt = L.optimize.BFGS(shard_rows=True)
        begin, end, per = opt._row_range(n)
        u = torch.zeros(per * world, dtype=torch.float64)
        wv = torch.zeros_like(u)
        v = torch.zeros(n, dtype=torch.float64)
        u[begin:end] = torch.from_numpy(h[begin:end].dot(yv))          # local rows of H y
        wv[begin:end] = torch.from_numpy(h[begin:end].dot(g))         # local rows of H g
        v += torch.from_numpy(yv[begin:end].dot(h[begin:end]))        # local rows' share of y^T H
        dist.all_gather_into_tensor(u, u[begin:begin + per].clone())
        dist.all_gather_into_tensor(wv, wv[begin:begin + per].clone())
        dist.all_reduce(v)
        assert numpy.allclose(u[:n].numpy(), h.dot(yv), rtol=1e-13, atol=1e-13)
        assert numpy.allclose(wv[:n].numpy(), h.dot(g), rtol=1e-13, atol=1e-13)
        assert numpy.allclose(v.numpy(), yv.dot(h), rtol=1e-13, atol=1e-13)
        with open(os.path.join(out_dir, "ok%d" % rank), "w") as fh:
            fh.write("ok")
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_contract(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(str(tmp_path))) == ["ok0", "ok1"]


def test_single_process_is_not_sharded():
    import learning_b200 as L
    model = L.MLP((3, 2, 2), optimizer=L.optimize.SteepestDescent())
    assert model._shard(10) == (0, 10, False)
    assert L.optimize.BFGS(shard_rows=True)._row_range(9) == (0, 9, 9)
