"""Parity of the row-sharded (data-parallel) paths against the single-process CPU oracle, for any number of
ranks -- TEST INFRASTRUCTURE shared by tests/test_gpu_multi.py (2 GPUs under pytest) and by bench.py, which runs
it under torchrun before timing so that every multi-GPU bench line carries a `parity_check`.

Every rank holds a contiguous row shard (distributed.shard_rows) and feeds its share of each global minibatch
(distributed.local_batch); shards and shares are deliberately UNEVEN (n % world != 0, miniBatchSize % world != 0).
The oracle replays the global minibatches in the order the ranks feed them: minibatch t = rank 0's rows, then
rank 1's, ...  The collectives only change the summation order, so the bars are the single-GPU ones.

Needs: torch.distributed initialised (any backend) and distributed.init_comm called -- see run_checks().
"""
import ctypes as C

import numpy as np


def _orders(n, mb, world, shard_rows, local_batch):
    """Row ranges / shares per rank, and the per-epoch minibatch lists of the AdaGrad / minibatch-SGD schedule
    (every rank runs max_r ceil(n_r / l_r) minibatches; a rank that ran out feeds 0 rows)."""
    spans = [shard_rows(n, r, world) for r in range(world)]
    shares = [local_batch(mb, r, world) for r in range(world)]
    T = max(-(-(e - b) // l) for (b, e), l in zip(spans, shares))
    batches = []
    for t in range(T):
        rows = [np.arange(min(e, b + t * l), min(e, b + (t + 1) * l)) for (b, e), l in zip(spans, shares)]
        batches.append(np.concatenate(rows))
    return spans, shares, batches


def run_checks(rank, world, n=97, mb=None):
    """Runs every sharded path on this rank's shard and compares with the oracle.  Returns
    {"ok": bool, "max_rel": worst relative error over all cases, "cases": {name: error}, "ranks": world}.
    Raises nothing: a failing case is reported through "ok" / "failed"."""
    import nimfm_b200 as nf
    from nimfm_b200 import _lib, distributed as nd
    from oracle import oracle as orc
    from oracle.oracle import CSR
    from helpers import make_dense, make_fm_params, make_field_csr, max_rel

    if mb is None:
        mb = 2 * world + 3 if world > 1 else 5          # never a multiple of the number of ranks
    d, k, degree = 10, 4, 3
    X = make_dense(n, d, 5, density=0.5, positive=False)
    y = np.sign(np.random.default_rng(0).standard_normal(n))
    P, w, _ = make_fm_params(d, degree, k, "explicit", True, seed=2, scale=0.1)
    csr = CSR.from_dense(X)
    spans, shares, batches = _orders(n, mb, world, nd.shard_rows, nd.local_batch)
    b, e = spans[rank]
    sh = orc.csr_take_rows(csr, np.arange(b, e))
    ds = nf.newCSRDataset(sh.data, sh.indices, sh.indptr, sh.n, d)
    cases, failed = {}, []

    def record(name, err, bar):
        cases[name] = float(err)
        if not (err <= bar):
            failed.append(f"{name}: {err:.3e} > {bar:.0e}")

    def fm_new():
        fm = nf.newFactorizationMachine(nf.classification, degree=degree, nComponents=k, warmStart=True)
        fm.P, fm.w, fm.intercept, fm.isInitialized = P.copy(), w.copy(), 0.05, True
        return fm

    # 1. decisionFunction: no collective, every rank predicts its shard
    got = fm_new().decisionFunction(ds)
    ref = orc.fm_decision_function(csr, P, w, 0.05, degree)[b:e]
    record("decision_function", max_rel(got, ref), 1e-10)

    # 2. MBPSGD (reduce-scatter -> sharded step / prox -> all-gather; SquaredL12 columns: all-reduce + dense step).
    # Every rank walks its own shard cyclically, l_r rows per minibatch, the cursor carried across epochs.
    inner, epochs = -(-n // mb), 3
    order = []
    for g in range(inner * epochs):
        for (sb, se), l in zip(spans, shares):
            order.append(sb + (g * l + np.arange(l)) % (se - sb))
    order = np.concatenate(order)
    perm_csr = orc.csr_take_rows(csr, order)
    kw = dict(eta0=0.1, alpha0=1e-6, alpha=1e-3, beta=1e-4, gamma=1e-3)
    for reg_name, reg in (("l1", nf.newL1()), ("l21", nf.newL21())):
        r2 = orc.mbpsgd_fit(perm_csr, y[order], P, w, 0.05, degree, "logistic", max_iter=epochs, reg=reg_name,
                            mini_batch_size=mb, max_iter_inner=inner, it=0, **kw)
        fm = fm_new()
        opt = nf.newMBPSGD(maxIter=epochs, loss=nf.Logistic(), reg=reg, miniBatchSize=mb, verbose=0, tol=0.0,
                           shuffle=False, **kw)
        opt.fit(ds, y[b:e], fm)
        record(f"mbpsgd_{reg_name}_epoch_loss", max_rel(opt.history, r2["epoch_loss"]), 1e-8)
        record(f"mbpsgd_{reg_name}_P", max_rel(fm.P, r2["P"]), 1e-8)
        record(f"mbpsgd_{reg_name}_w", max_rel(fm.w, r2["w"]), 1e-8)
        record(f"mbpsgd_{reg_name}_intercept", abs(fm.intercept - r2["intercept"]), 1e-9)
    # degree 2 with the default regulariser (column-wise SquaredL12): the all-reduce route
    P2, w2, _ = make_fm_params(d, 2, k, "explicit", True, seed=3, scale=0.1)
    r2 = orc.mbpsgd_fit(perm_csr, y[order], P2, w2, 0.05, 2, "logistic", max_iter=2, reg="squaredl12",
                        mini_batch_size=mb, max_iter_inner=inner, it=0, **kw)
    fm = nf.newFactorizationMachine(nf.classification, degree=2, nComponents=k, warmStart=True)
    fm.P, fm.w, fm.intercept, fm.isInitialized = P2.copy(), w2.copy(), 0.05, True
    opt = nf.newMBPSGD(maxIter=2, loss=nf.Logistic(), miniBatchSize=mb, verbose=0, tol=0.0, shuffle=False, **kw)
    opt.fit(ds, y[b:e], fm)
    record("mbpsgd_squaredl12_epoch_loss", max_rel(opt.history, r2["epoch_loss"][:2]), 1e-8)
    record("mbpsgd_squaredl12_P", max_rel(fm.P, r2["P"]), 1e-8)

    # 3. AdaGrad, synchronous minibatch: the oracle is chained one (variable-size) global minibatch at a time
    # with its state carried; every call ends in finalize, which is what the device's fit ends in as well
    st, itr, Pr, wr, br = None, 1, P, w, 0.05
    loss_ep = []
    for ep in range(2):
        ls = 0.0
        for rows in batches:
            sub = orc.csr_take_rows(csr, rows)
            r3 = orc.adagrad_fit(sub, y[rows], Pr, wr, br, degree, "logistic", max_iter=1, mini_batch_size=len(rows),
                                 it=itr, state=st)
            st, itr, Pr, wr, br = r3["state"], r3["it"], r3["P"], r3["w"], r3["intercept"]
            ls += r3["loss"][0] * len(rows)
        loss_ep.append(ls / n)
    fm = fm_new()
    opt = nf.newAdaGrad(maxIter=2, loss=nf.Logistic(), miniBatchSize=mb, verbose=0, tol=0.0, shuffle=False)
    opt.fit(ds, y[b:e], fm)
    record("adagrad_P", max_rel(fm.P, Pr), 1e-8)
    record("adagrad_w", max_rel(fm.w, wr), 1e-8)
    record("adagrad_loss", max_rel([h[1] for h in opt.history], loss_ep), 1e-8)
    record("adagrad_it", abs(opt.it - itr), 0)
    # bit-identical replicas: every rank must hold exactly rank 0's parameters
    same = nd.allgather_i64(np.frombuffer(np.ascontiguousarray(fm.P).tobytes(), dtype=np.int64)[:4096])
    record("adagrad_replicas_identical", float(np.any(same != same[0])), 0)

    # 3b. synchronous-minibatch SGD (the device analogue of Hogwild fit(..., maxThreads)), chained the same way
    kws = dict(eta0=0.02, alpha0=1e-4, alpha=1e-2, beta=2e-2)
    itr, Pr, wr, br = 1, P, w, 0.05
    viol_ep, loss_ep = [], []
    for ep in range(2):
        v, ls = 0.0, 0.0
        for rows in batches:
            sub = orc.csr_take_rows(csr, rows)
            r3 = orc.sgd_minibatch_fit(sub, y[rows], Pr, wr, br, degree, "logistic", B=len(rows), max_iter=1, it=itr, **kws)
            itr, Pr, wr, br = r3["it"], r3["P"], r3["w"], r3["intercept"]
            v += r3["viol"][0]
            ls += r3["loss"][0] * len(rows)
        viol_ep.append(v)
        loss_ep.append(ls / n)
    fm = fm_new()
    opt = nf.newSGD(maxIter=2, loss=nf.Logistic(), miniBatchSize=mb, verbose=0, tol=0.0, shuffle=False, **kws)
    opt.fit(ds, y[b:e], fm)
    record("sgd_minibatch_P", max_rel(fm.P, Pr), 1e-9)
    record("sgd_minibatch_w", max_rel(fm.w, wr), 1e-9)
    record("sgd_minibatch_it", abs(opt.it - itr), 0)
    record("sgd_minibatch_viol", max_rel([h[0] for h in opt.history], viol_ep), 1e-8)
    record("sgd_minibatch_loss", max_rel([h[1] for h in opt.history], loss_ep), 1e-8)

    # 4. FFM: predict+grad with the gradient all-reduce == full-batch oracle gradient; AdaGrad minibatch epochs
    Xf, fcsr, _ = make_field_csr(n, 12, 4, 9)
    Pf = np.random.default_rng(3).standard_normal((4, 12, 4)) * 0.1
    yf = np.random.default_rng(4).standard_normal(n)

    def take_f(rows):
        s = orc.csr_take_rows(fcsr, rows)
        s.fields = (np.concatenate([fcsr.fields[fcsr.indptr[r]:fcsr.indptr[r + 1]] for r in rows]).astype(np.int64)
                    if len(rows) else np.zeros(0, np.int64))
        s.n_fields = 4
        return s
    fsh = take_f(np.arange(b, e))
    fds = nf.newCSRFieldDataset(fsh.data, fsh.indices, fsh.indptr, fsh.fields, fsh.n, 12, 4)
    fds.set_targets(yf[b:e])
    m = nf.newFieldAwareFactorizationMachine(nf.regression, nComponents=4, warmStart=True)
    m.P, m.w, m.intercept, m.isInitialized = Pf.copy(), np.zeros(12), 0.0, True
    lib, ctx = _lib.load(), _lib.ctx()
    h = m._to_device(fds)
    ls = C.c_double()
    _lib.check(lib.nimfm_ffm_loss_grad(ctx, h, fds.handle(), 0, 1.0, 0, fsh.n, None, n, 1, 1, C.byref(ls)))
    gP, gw, gb = np.zeros_like(Pf), np.zeros(12), C.c_double()
    _lib.check(lib.nimfm_ffm_get_grads(ctx, h, _lib.ptr(gP), _lib.ptr(gw), C.byref(gb)))
    lib.nimfm_ffm_free(ctx, h)
    rf = orc.ffm_loss_grad(fcsr, yf, Pf, np.zeros(12), 0.0, "squared")
    record("ffm_grad_P", max_rel(gP, rf["gP"]), 1e-9)
    record("ffm_grad_loss", abs(ls.value - rf["loss"]) / abs(rf["loss"]), 1e-9)
    st, itr, Pr, wr, br = None, 1, Pf, np.zeros(12), 0.0
    for ep in range(2):
        for rows in batches:
            r4 = orc.ffm_adagrad_fit(take_f(rows), yf[rows], Pr, wr, br, "squared", max_iter=1,
                                     mini_batch_size=len(rows), it=itr, state=st)
            st, itr, Pr, wr, br = r4["state"], r4["it"], r4["P"], r4["w"], r4["intercept"]
    m = nf.newFieldAwareFactorizationMachine(nf.regression, nComponents=4, warmStart=True)
    m.P, m.w, m.intercept, m.isInitialized = Pf.copy(), np.zeros(12), 0.0, True
    opt = nf.newAdaGrad(maxIter=2, loss=nf.Squared(), miniBatchSize=mb, verbose=0, tol=0.0, shuffle=False)
    opt.fit(fds, yf[b:e], m)
    record("ffm_adagrad_P", max_rel(m.P, Pr), 1e-8)
    record("ffm_adagrad_w", max_rel(m.w, wr), 1e-8)

    worst = max(v for k_, v in cases.items() if not k_.endswith(("_it", "_identical")))
    return {"ok": not failed, "max_rel": worst, "ranks": world, "n": n, "miniBatchSize": mb,
            "shard_rows": [e_ - b_ for b_, e_ in spans], "shares": shares, "cases": cases, "failed": failed}
