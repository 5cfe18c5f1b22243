#!/usr/bin/env python
"""bench_multi.py -- the row-sharded configs at N GPUs (weak scaling: every rank holds its own shard),
one process per GPU:  python -m torch.distributed.run --nproc-per-node N ... bench_multi.py [--rows R]

  C4-predict  batched decisionFunction (row kernel only, shards independent, no collective)
  C3-mbpsgd   MBPSGD epoch, FM degree 2 rank 16, global minibatch = N x 256Ki rows, grad all-reduce per minibatch
  C4-adagrad  AdaGrad synchronous minibatch epoch, HOFM degree 3 rank 32, global minibatch = N x 512Ki rows
  C5-ffm      FFM predict+grad pass + all-reduce of the 2.5 GB gradient

Times are device/host times of the blocking library calls, MAX over ranks; one JSON line per config
from rank 0.  bench.py (the driver's contract) carries the headline C4 predict+grad line.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import bench_configs  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=2_000_000, help="rows per GPU")
    ap.add_argument("--ffm-rows", type=int, default=400_000, help="FFM rows per GPU")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist
    import nimfm_b200 as nf
    from nimfm_b200 import _lib, distributed as nd
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    lib, ctx = _lib.load(), _lib.ctx(lr)
    nd.init_comm(rank, world)
    want = set(args.only.split(",")) if args.only else None

    def on(tag):
        return want is None or tag in want

    def maxr(v):
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def emit(d):
        if rank == 0:
            d.update(n_gpus=world, scaling="weak")
            print(json.dumps(d), flush=True)

    n = args.rows
    if on("C4-predict") or on("C3-mbpsgd") or on("C4-adagrad"):
        data, idx, ptr, y = bench.gen_criteo_rows(n, 5000 + rank)
        ds = nf.newCSRDataset(data, idx, ptr, n, bench.D_FEATURES)
        ds.set_targets(y)

    if on("C4-predict"):
        P, w, b = bench.model_params(7)
        fm = nf.newFactorizationMachine(nf.classification, degree=3, nComponents=32, warmStart=True)
        fm.P, fm.w, fm.intercept, fm.isInitialized = P, w, b, True
        h = fm._to_device(bench.D_FEATURES)
        ms = C.c_float()
        barrier()
        _lib.check(lib.nimfm_fm_time_loss_grad(ctx, h, ds.handle(), 2, n, n, 5, 0, C.byref(ms)))
        t = maxr(ms.value)
        lib.nimfm_fm_free(ctx, h)
        emit({"config": "C4-predict", "what": "batched decisionFunction kernel, HOFM degree 3 rank 32, rows sharded",
              "rows_per_gpu": n, "ms_per_pass": t, "samples_per_s": n * world / (t / 1e3)})

    if on("C3-mbpsgd"):
        rng = np.random.default_rng(2)
        fm = nf.newFactorizationMachine(nf.classification, degree=2, nComponents=16, warmStart=True)
        fm.P, fm.w = rng.standard_normal((1, 16, bench.D_FEATURES)) * 0.01, np.zeros(bench.D_FEATURES)
        fm.intercept, fm.isInitialized = 0.0, True
        mb = world * (1 << 18)
        opt = nf.newMBPSGD(maxIter=3, eta0=0.1, alpha0=1e-6, alpha=1e-3, beta=1e-4, gamma=0.0, reg=nf.newL1(),
                           loss=nf.Logistic(), miniBatchSize=mb, verbose=0, tol=0.0, shuffle=False)
        barrier()
        opt.fit(ds, y, fm)
        t = maxr(float(np.min(opt.epoch_seconds)))
        inner = max((n - 1) // (mb // world) + 1, 1)
        emit({"config": "C3-mbpsgd", "what": "MBPSGD epoch, FM degree 2 rank 16, logistic, grad all-reduce per minibatch",
              "rows_per_gpu": n, "global_minibatch": mb, "inner_iterations": inner, "s_per_epoch": t,
              "samples_per_s": inner * mb / t, "epoch_losses": opt.history,
              "allreduce_doubles_per_minibatch": 16 * bench.D_FEATURES + bench.D_FEATURES + 2})

    if on("C4-adagrad"):
        P, w, b = bench.model_params(7)
        fm = nf.newFactorizationMachine(nf.classification, degree=3, nComponents=32, warmStart=True)
        fm.P, fm.w, fm.intercept, fm.isInitialized = P, w, b, True
        mb = world * (1 << 19)
        # eta0: the synchronous variant accumulates SUMS over the minibatch (adagrad.nim:119-124 semantics per
        # sample), so its first dual-averaging step scales like eta0*sqrt(minibatch): keep it small at 512Ki rows
        opt = nf.newAdaGrad(maxIter=2, eta0=1e-4, eps=1e-10, loss=nf.Logistic(), verbose=0, tol=0.0, shuffle=False,
                            miniBatchSize=mb)
        barrier()
        opt.fit(ds, y, fm)
        t = maxr(float(np.min(opt.epoch_seconds)))
        emit({"config": "C4-adagrad", "what": "AdaGrad synchronous-minibatch epoch, HOFM degree 3 rank 32, logistic",
              "rows_per_gpu": n, "global_minibatch": mb, "s_per_epoch": t, "samples_per_s": n * world / t,
              "epoch_loss": opt.history[-1][1]})

    if on("C5-ffm"):
        nf_rows = args.ffm_rows
        data, idx, ptr, fields, y, d = bench_configs.gen_ffm_rows(nf_rows, 6000 + rank)
        fds = nf.newCSRFieldDataset(data, idx, ptr, fields, nf_rows, d, 39)
        fds.set_targets(y)
        rng = np.random.default_rng(3)
        m = nf.newFieldAwareFactorizationMachine(nf.classification, nComponents=8, warmStart=True)
        m.P = rng.standard_normal((39, d, 8)) * 0.01
        m.w, m.intercept, m.isInitialized = np.zeros(d), 0.0, True
        h = m._to_device(fds)
        ls = C.c_double()

        def step():
            _lib.check(lib.nimfm_ffm_loss_grad(ctx, h, fds.handle(), 2, 1.0, 0, nf_rows, None, nf_rows * world, 1,
                                               int(world > 1), C.byref(ls)))
        step()
        barrier()
        _lib.check(lib.nimfm_timer_start(ctx))
        for _ in range(3):
            step()
        ms = C.c_float()
        _lib.check(lib.nimfm_timer_stop(ctx, C.byref(ms)))
        t = maxr(ms.value / 3)
        lib.nimfm_ffm_free(ctx, h)
        emit({"config": "C5-ffm", "what": "FFM 39 fields rank 8 predict+grad pass (zero grads + pair kernel + all-reduce "
              "of [gP | gw | gb, loss])", "rows_per_gpu": nf_rows, "ms_per_step": t,
              "samples_per_s": nf_rows * world / (t / 1e3), "allreduce_doubles": 39 * d * 8 + d + 2})

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
