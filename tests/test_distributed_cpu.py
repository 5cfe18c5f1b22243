"""world_size-2 gloo tests (CPU) of the host-side logic of the row-sharded paths: shard ranges,
unique-id exchange, and that per-rank gradients of row shards SUM to the full-batch gradient (the
identity the NCCL all-reduce relies on) -- checked with the oracle standing in for each rank's GPU."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch
    from nimfm_b200 import distributed as nd
    from oracle import oracle as orc
    from oracle.oracle import CSR
    from helpers import make_dense, make_fm_params

    # 1. the 128-byte communicator id reaches every rank unchanged
    token = nd.exchange_unique_id(rank, lambda: bytes(range(128)))
    assert token == bytes(range(128))

    # 2. row shards partition [0, n)
    n = 101
    b, e = nd.shard_rows(n, rank, world)
    sizes = [None] * world
    dist.all_gather_object(sizes, (b, e))
    assert sizes[0][0] == 0 and sizes[-1][1] == n
    assert all(sizes[i][1] == sizes[i + 1][0] for i in range(world - 1))
    assert max(s[1] - s[0] for s in sizes) - min(s[1] - s[0] for s in sizes) <= 1
    assert sum(nd.local_batch(37, r, world) for r in range(world)) == 37

    # 3. sum over ranks of shard gradients (coef = dloss / GLOBAL minibatch) == full-batch gradient
    d, k, degree = 9, 4, 3
    X = make_dense(n, d, 5, density=0.5, positive=False)
    y = np.sign(np.random.default_rng(0).standard_normal(n))
    P, w, _ = make_fm_params(d, degree, k, "explicit", True, seed=2, scale=0.2)
    csr = CSR.from_dense(X)
    shard = orc.csr_take_rows(csr, np.arange(b, e))
    g = orc.fm_loss_grad(shard, y[b:e], P, w, 0.1, degree, "logistic", mini_batch_size=n)
    flat = np.concatenate([g["gP"].ravel(), g["gw"], [g["gb"], g["loss"]]])
    t = torch.from_numpy(flat.copy())
    dist.all_reduce(t)                     # what ncclAllReduce(sum) does to [grad P | grad w | gb, loss]
    full = orc.fm_loss_grad(csr, y, P, w, 0.1, degree, "logistic", mini_batch_size=n)
    ref = np.concatenate([full["gP"].ravel(), full["gw"], [full["gb"], full["loss"]]])
    np.testing.assert_allclose(t.numpy(), ref, rtol=1e-12, atol=1e-15)
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


def test_world2_gloo(tmp_path):
    from oracle import oracle as orc
    orc.build()
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_shard_rows_edge_cases():
    from nimfm_b200.distributed import shard_rows
    assert shard_rows(0, 0, 4) == (0, 0)
    assert [shard_rows(3, r, 4) for r in range(4)] == [(0, 1), (1, 2), (2, 3), (3, 3)]
    assert [shard_rows(10, r, 3) for r in range(3)] == [(0, 4), (4, 7), (7, 10)]
