"""Row-sharded (data-parallel) paths on 2 GPUs of one box, one process per GPU over NCCL: batched
decisionFunction, MBPSGD with the per-minibatch gradient all-reduce, synchronous-minibatch AdaGrad
(FM and FFM) -- each compared with the single-process oracle on the FULL data (the all-reduce only
changes the summation order).  Skipped on a 1-GPU box (run with `gpurun --gpus 2`)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import nimfm_b200 as nf
    from nimfm_b200 import distributed as nd
    from oracle import oracle as orc
    from oracle.oracle import CSR
    from helpers import make_dense, make_fm_params, make_field_csr, max_rel
    nd.init_comm(rank, world)

    n, d, k, degree = 96, 10, 4, 3
    X = make_dense(n, d, 5, density=0.5, positive=False)
    y = np.sign(np.random.default_rng(0).standard_normal(n))
    P, w, _ = make_fm_params(d, degree, k, "explicit", True, seed=2, scale=0.1)
    csr = CSR.from_dense(X)
    b, e = nd.shard_rows(n, rank, world)
    sh = orc.csr_take_rows(csr, np.arange(b, e))
    ds = nf.newCSRDataset(sh.data, sh.indices, sh.indptr, sh.n, d)

    def fm_new():
        fm = nf.newFactorizationMachine(nf.classification, degree=degree, nComponents=k, warmStart=True)
        fm.P, fm.w, fm.intercept, fm.isInitialized = P.copy(), w.copy(), 0.05, True
        return fm

    # 1. decisionFunction: no collective, every rank predicts its shard
    got = fm_new().decisionFunction(ds)
    ref = orc.fm_decision_function(csr, P, w, 0.05, degree)[b:e]
    assert max_rel(got, ref) <= 1e-10

    # 2. MBPSGD: global minibatch 16 = 8 rows per rank per step.  The oracle sees the global minibatches
    # in the order the ranks feed them: step t = rows [8t, 8t+8) of shard 0 then of shard 1.
    mb, local = 16, 8
    order = np.concatenate([np.concatenate([np.arange(r * (n // world) + t * local, r * (n // world) + (t + 1) * local)
                                            for r in range(world)]) for t in range(n // mb)])
    perm_csr = orc.csr_take_rows(csr, order)
    kw = dict(eta0=0.1, alpha0=1e-6, alpha=1e-3, beta=1e-4, gamma=1e-3)
    r2 = orc.mbpsgd_fit(perm_csr, y[order], P, w, 0.05, degree, "logistic", max_iter=3, reg="l1", mini_batch_size=mb,
                        it=0, **kw)
    fm = fm_new()
    opt = nf.newMBPSGD(maxIter=3, loss=nf.Logistic(), reg=nf.newL1(), miniBatchSize=mb, verbose=0, tol=0.0,
                       shuffle=False, **kw)
    opt.fit(ds, y[b:e], fm)
    np.testing.assert_allclose(opt.history, r2["epoch_loss"], rtol=1e-8)
    assert max_rel(fm.P, r2["P"]) <= 1e-8 and max_rel(fm.w, r2["w"]) <= 1e-8
    assert abs(fm.intercept - r2["intercept"]) <= 1e-9

    # 3. AdaGrad, synchronous minibatch (global 16): identical parameters on every rank, equal to the
    # oracle's minibatch variant on the interleaved order
    r3 = orc.adagrad_fit(perm_csr, y[order], P, w, 0.05, degree, "logistic", max_iter=2, mini_batch_size=mb)
    fm = fm_new()
    opt = nf.newAdaGrad(maxIter=2, loss=nf.Logistic(), miniBatchSize=mb, verbose=0, tol=0.0, shuffle=False)
    opt.fit(ds, y[b:e], fm)
    assert max_rel(fm.P, r3["P"]) <= 1e-8 and max_rel(fm.w, r3["w"]) <= 1e-8
    t = torch.from_numpy(np.ascontiguousarray(fm.P)).cuda()
    tmax = t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    assert torch.equal(t, tmax)                       # bit-identical replicas

    # 3b. SGD, synchronous minibatch (global 16; the device analogue of Hogwild fit(..., maxThreads)): touch
    # counts and gradients all-reduced, identical step on every rank == the restated rule on the interleaved order
    kws = dict(eta0=0.02, alpha0=1e-4, alpha=1e-2, beta=2e-2)
    r3b = orc.sgd_minibatch_fit(perm_csr, y[order], P, w, 0.05, degree, "logistic", B=mb, max_iter=2, it=1, **kws)
    fm = fm_new()
    opt = nf.newSGD(maxIter=2, loss=nf.Logistic(), miniBatchSize=mb, verbose=0, tol=0.0, shuffle=False, **kws)
    opt.fit(ds, y[b:e], fm)
    assert max_rel(fm.P, r3b["P"]) <= 1e-9 and max_rel(fm.w, r3b["w"]) <= 1e-9 and opt.it == r3b["it"]
    np.testing.assert_allclose([h_[0] for h_ in opt.history], r3b["viol"], rtol=1e-8)
    np.testing.assert_allclose([h_[1] for h_ in opt.history], r3b["loss"], rtol=1e-8)

    # 4. FFM predict+grad with the gradient all-reduce == full-batch oracle gradient
    Xf, fcsr, _ = make_field_csr(n, 12, 4, 9)
    Pf = np.random.default_rng(3).standard_normal((4, 12, 4)) * 0.1
    yf = np.random.default_rng(4).standard_normal(n)
    fsh = orc.csr_take_rows(fcsr, np.arange(b, e))
    fsh.fields = np.concatenate([fcsr.fields[fcsr.indptr[r]:fcsr.indptr[r + 1]] for r in range(b, e)]).astype(np.int64)
    fds = nf.newCSRFieldDataset(fsh.data, fsh.indices, fsh.indptr, fsh.fields, fsh.n, 12, 4)
    fds.set_targets(yf[b:e])
    m = nf.newFieldAwareFactorizationMachine(nf.regression, nComponents=4, warmStart=True)
    m.P, m.w, m.intercept, m.isInitialized = Pf, np.zeros(12), 0.0, True
    import ctypes as C
    from nimfm_b200 import _lib
    lib, ctx = _lib.load(), _lib.ctx()
    h = m._to_device(fds)
    ls = C.c_double()
    _lib.check(lib.nimfm_ffm_loss_grad(ctx, h, fds.handle(), 0, 1.0, 0, fsh.n, None, n, 1, 1, C.byref(ls)))
    gP, gw, gb = np.zeros_like(Pf), np.zeros(12), C.c_double()
    _lib.check(lib.nimfm_ffm_get_grads(ctx, h, _lib.ptr(gP), _lib.ptr(gw), C.byref(gb)))
    lib.nimfm_ffm_free(ctx, h)
    rf = orc.ffm_loss_grad(fcsr, yf, Pf, np.zeros(12), 0.0, "squared")
    assert max_rel(gP, rf["gP"]) <= 1e-9 and abs(ls.value - rf["loss"]) <= 1e-9 * abs(rf["loss"])

    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_row_sharded_paths(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    from oracle import oracle as orc
    orc.build()
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
