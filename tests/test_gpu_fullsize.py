"""Parity at BASELINE.json's FULL sizes: the C4 workload of bench.py -- 10 M rows x 39 nnz, 1 M features, HOFM
degree 3 rank 32 -- through size-independent properties (the oracle cannot run 10 M rows of it in seconds); C3
(FM degree 2 rank 16, MBPSGD, the same 10 M rows) against the oracle's predict+grad replayed on all host threads
for the full epoch at 1 Mi-row minibatches and against the oracle solver itself at the reference-default
minibatch; C2 (ML-100K shape, HOFM-3 CD) against the oracle outright; C5 (FFM, 10 M rows) through halves that
add up and the oracle on a random row sample."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench            # noqa: E402
import bench_configs    # noqa: E402
import nimfm_b200 as nf  # noqa: E402
from nimfm_b200 import _lib  # noqa: E402
from oracle.oracle import CSR  # noqa: E402
from helpers import max_rel, sums_agree  # noqa: E402

pytestmark = pytest.mark.gpu
N_FULL = int(os.environ.get("NIMFM_FULLSIZE_ROWS", 10_000_000))


@pytest.fixture(scope="module")
def c4():
    data, indices, indptr, y = bench.gen_criteo_rows(N_FULL, 4242)
    ds = nf.newCSRDataset(data, indices, indptr, N_FULL, bench.D_FEATURES)
    ds.set_targets(y)
    P, w, b = bench.model_params(11)
    fm = nf.newFactorizationMachine(nf.classification, degree=3, nComponents=32, warmStart=True)
    fm.P, fm.w, fm.intercept, fm.isInitialized = P, w, 0.05, True
    h = fm._to_device(bench.D_FEATURES)
    yield dict(ds=ds, y=y, fm=fm, h=h, csr=CSR(data, indices, indptr, N_FULL, bench.D_FEATURES))
    _lib.load().nimfm_fm_free(_lib.ctx(), h)
    ds.free()


def loss_grad(c, row_begin, n_rows, rows=None, mb=None):
    lib, ctx = _lib.load(), _lib.ctx()
    ls = C.c_double()
    idx = None if rows is None else _lib.i64(rows)
    nr = n_rows if rows is None else len(rows)
    _lib.check(lib.nimfm_fm_loss_grad(ctx, c["h"], c["ds"].handle(), 2, 1.0, row_begin, nr, _lib.ptr(idx),
                                      N_FULL if mb is None else mb, 1, 0, C.byref(ls)))
    gP, gw, gb = np.zeros_like(c["fm"].P), np.zeros(bench.D_FEATURES), C.c_double()
    _lib.check(lib.nimfm_fm_get_grads(ctx, c["h"], _lib.ptr(gP), _lib.ptr(gw), C.byref(gb)))
    return ls.value, gP, gw, gb.value


def test_c4_bookkeeping_and_prediction_full_size(oracle, c4):
    ds, csr = c4["ds"], c4["csr"]
    inf = ds.info()
    assert inf["n"] == N_FULL and inf["nnz"] == N_FULL * 39 and inf["maxSegNnz"] == 39      # nnz bookkeeping exact
    # device-resident rows are bit-exact: a random row gather equals the host slices
    rows = np.random.default_rng(1).integers(0, N_FULL, 5000)
    sub = ds[rows]
    ref = oracle.csr_take_rows(csr, rows)
    assert np.array_equal(sub.indices, ref.indices) and np.array_equal(sub.data, ref.data)
    assert np.array_equal(sub.indptr, ref.indptr)
    # decisionFunction over all 10 M rows: the resident kernel and the chunked host-streaming path agree
    # bit for bit (a row's value depends on that row only), and a random sample matches the oracle
    lib, ctx = _lib.load(), _lib.ctx()
    out_dev, out_host = np.zeros(N_FULL), np.zeros(N_FULL)
    _lib.check(lib.nimfm_fm_decision_function(ctx, c4["h"], ds.handle(), _lib.ptr(out_dev)))
    _lib.check(lib.nimfm_fm_decision_function_host(ctx, c4["h"], N_FULL, bench.D_FEATURES, _lib.ptr(csr.data),
                                                   _lib.ptr(csr.indices), _lib.ptr(csr.indptr), 0, _lib.ptr(out_host)))
    assert np.array_equal(out_dev, out_host)
    want = oracle.fm_decision_function(ref, c4["fm"].P, c4["fm"].w, 0.05, 3)
    assert max_rel(out_dev[rows], want) <= 1e-10


def test_c4_gradient_full_size(oracle, c4):
    # linearity: the 10 M-row gradient is the sum of the gradients of its two halves
    ls, gP, gw, gb = loss_grad(c4, 0, N_FULL)
    h1 = N_FULL // 2
    lsa, gPa, gwa, gba = loss_grad(c4, 0, h1)
    lsb, gPb, gwb, gbb = loss_grad(c4, h1, N_FULL - h1)
    assert abs(ls - (lsa + lsb)) <= 1e-10 * abs(ls)
    assert sums_agree(gPa + gPb, gP) and sums_agree(gwa + gwb, gw)
    assert abs(gb - (gba + gbb)) <= 1e-10 * max(abs(gb), 1e-6)
    # the loss sum equals the loss of the full-size predictions
    # and a random row list matches the oracle's predict+grad on the same rows
    rows = np.random.default_rng(2).integers(0, N_FULL, 3000)
    lsr, gPr, gwr, gbr = loss_grad(c4, 0, 0, rows=rows, mb=len(rows))
    sub = oracle.csr_take_rows(c4["csr"], rows)
    ref = oracle.fm_loss_grad(sub, c4["y"][rows], c4["fm"].P, c4["fm"].w, 0.05, 3, "logistic", mini_batch_size=len(rows))
    assert abs(lsr - ref["loss"]) <= 1e-10 * abs(ref["loss"])
    assert max_rel(gPr, ref["gP"]) <= 1e-9 and max_rel(gwr, ref["gw"]) <= 1e-9
    # host streaming of the first 3 M rows == the resident kernel on the same rows
    m = min(3_000_000, N_FULL)
    lib, ctx = _lib.load(), _lib.ctx()
    lsh = C.c_double()
    csr = c4["csr"]
    _lib.check(lib.nimfm_fm_loss_grad_host(ctx, c4["h"], m, bench.D_FEATURES, _lib.ptr(csr.data), _lib.ptr(csr.indices),
                                           _lib.ptr(csr.indptr), _lib.ptr(c4["y"]), 2, 1.0, N_FULL, 0, 1, 0, C.byref(lsh)))
    gPh = np.zeros_like(gP)
    _lib.check(lib.nimfm_fm_get_grads(ctx, c4["h"], _lib.ptr(gPh), None, None))
    lsm, gPm, _, _ = loss_grad(c4, 0, m)
    assert abs(lsh.value - lsm) <= 1e-10 * abs(lsm) and sums_agree(gPh, gPm)


def _oracle_grad_all_threads(orc, csr, y, Pf, w, b, r0, r1, mb, k, threads):
    """predict+grad of rows [r0, r1) (minibatch_psgd.updateGradient) by the oracle port, the rows dealt to
    `threads` host threads (each with its own gradient buffers; ctypes releases the GIL), summed in thread order"""
    from concurrent.futures import ThreadPoolExecutor
    lib = orc.lib()
    d = csr.d
    cuts = np.linspace(r0, r1, threads + 1).astype(np.int64)

    def work(t):
        gP, gw, gb, dA = np.zeros_like(Pf), np.zeros(d), C.c_double(0.0), np.zeros_like(Pf)
        ls = lib.ref_fm_loss_grad(C.c_int64(d), orc._d(csr.data), orc._i(csr.indices), orc._i(csr.indptr), orc._d(y),
                                  C.c_int64(int(cuts[t])), C.c_int64(int(cuts[t + 1])), C.c_int(2), C.c_int(k), C.c_int(1),
                                  C.c_int(0), C.c_int(1), C.c_int(1), orc._d(Pf), orc._d(w), C.c_double(b), C.c_int(2),
                                  C.c_double(1.0), C.c_int64(mb), orc._d(gP), orc._d(gw), C.byref(gb), None, orc._d(dA))
        return ls, gP, gw, gb.value
    with ThreadPoolExecutor(threads) as ex:
        parts = list(ex.map(work, range(threads)))
    ls = sum(p[0] for p in parts)
    gP, gw, gb = parts[0][1], parts[0][2], parts[0][3]
    for p in parts[1:]:
        gP += p[1]
        gw += p[2]
        gb += p[3]
    return ls, gP, gw, gb


def test_c3_mbpsgd_epoch_full_size(oracle, c4):
    """C3: FM degree 2 rank 16, MBPSGD logistic (gamma = 0) over the full 10 M-row shard.
    (a) one epoch at 1 Mi-row minibatches (dense step): the oracle's updateGradient replayed per minibatch on all
        host threads + Params.step as the reference writes it (params.nim:90-98) -- epoch loss and parameters <= 1e-8;
    (b) the reference-default minibatch (d*n div nnz = 25 641 rows, the lazy touched-features-only epoch): the
        first 40 minibatches against the oracle SOLVER on the same rows -- epoch loss and parameters <= 1e-8."""
    ds, csr, y = c4["ds"], c4["csr"], c4["y"]
    d, k = bench.D_FEATURES, 16
    rng = np.random.default_rng(21)
    P0 = rng.standard_normal((1, k, d)) * 0.01
    w0 = rng.standard_normal(d) * 0.01
    kw = dict(eta0=0.1, alpha0=1e-6, alpha=1e-3, beta=1e-4)
    lib, ctx = _lib.load(), _lib.ctx()

    def device_epoch(mb, inner):
        fm = nf.newFactorizationMachine(nf.classification, degree=2, nComponents=k, warmStart=True)
        fm.P, fm.w, fm.intercept, fm.isInitialized = P0.copy(), w0.copy(), 0.02, True
        h = fm._to_device(d)
        try:
            cfg = _lib.MbpsgdCfg(2, 1.0, kw["eta0"], kw["alpha0"], kw["alpha"], kw["beta"], 0.0, _lib.REG_L1,
                                 _lib.SCHED["optimal"], 1.0, mb, inner)
            it, ii, rl = C.c_int64(1), C.c_int64(0), C.c_double()
            _lib.check(lib.nimfm_fm_mbpsgd_epoch(ctx, h, ds.handle(), C.byref(cfg), mb, C.byref(it), C.byref(ii), None,
                                                 C.byref(rl)))
            fm._from_device(h)
        finally:
            lib.nimfm_fm_free(ctx, h)
        return rl.value, fm, it.value, ii.value

    # (a) the whole shard, 1 Mi-row minibatches
    mb = 1 << 20
    inner = (N_FULL - 1) // mb + 1
    loss_dev, fm, it_dev, ii_dev = device_epoch(mb, inner)
    threads = min(os.cpu_count() or 1, 32)
    Pf, w, b, cur, it, loss_sum = oracle.to_feature_major(P0), w0.copy(), 0.02, 0, 1, 0.0
    for t in range(inner):
        gP, gw, gb = None, None, 0.0
        left, start = mb, cur
        while left > 0:                                   # the cyclic cursor of epoch() (minibatch_psgd.nim:102-111)
            take = min(left, N_FULL - start)
            ls, gP1, gw1, gb1 = _oracle_grad_all_threads(oracle, csr, y, Pf, w, b, start, start + take, mb, k, threads)
            loss_sum += ls
            gP = gP1 if gP is None else gP + gP1
            gw = gw1 if gw is None else gw + gw1
            gb += gb1
            left -= take
            start = (start + take) % N_FULL
        cur = start
        etaP, etaW, etaB = (oracle.get_eta("optimal", kw["eta0"], 1.0, r, it) for r in (kw["beta"], kw["alpha"], kw["alpha0"]))
        Pf += (-etaP) * gP                                  # add, then scale by the reciprocal (two roundings)
        Pf *= 1.0 / (1.0 + etaP * kw["beta"])
        w += (-etaW) * gw
        w *= 1.0 / (1.0 + etaW * kw["alpha"])
        b += (-etaB) * gb
        b *= 1.0 / (1.0 + etaB * kw["alpha0"])
        it += 1
    loss_ref = loss_sum / (mb * inner)
    assert it_dev == it and ii_dev == cur
    assert abs(loss_dev - loss_ref) <= 1e-8 * abs(loss_ref)
    assert max_rel(fm.P, oracle.to_component_major(Pf)) <= 1e-8 and max_rel(fm.w, w) <= 1e-8
    assert abs(fm.intercept - b) <= 1e-9

    # (b) the reference-default minibatch: the first 40 minibatches == the oracle solver on those rows
    mb = max(d * N_FULL // (N_FULL * 39), 1)
    inner = min(40, (N_FULL - 1) // mb + 1)
    rows = mb * inner
    if rows <= N_FULL:
        loss_dev, fm, it_dev, _ = device_epoch(mb, inner)
        sub = CSR(csr.data[:rows * 39], csr.indices[:rows * 39], csr.indptr[:rows + 1], rows, d)
        ref = oracle.mbpsgd_fit(sub, y[:rows], P0, w0, 0.02, 2, "logistic", max_iter=1, gamma=0.0, reg="l1",
                                mini_batch_size=mb, max_iter_inner=inner, it=1, **kw)
        assert it_dev == ref["it"]
        assert abs(loss_dev - ref["epoch_loss"][0]) <= 1e-8 * abs(ref["epoch_loss"][0])
        assert max_rel(fm.P, ref["P"]) <= 1e-8 and max_rel(fm.w, ref["w"]) <= 1e-8
        assert abs(fm.intercept - ref["intercept"]) <= 1e-9


def test_c2_cd_full_size(oracle):
    """C1 / C2: ML-100K shape (100 k rows, 943 + 1682 one-hot features), rank 30, CD with squared loss: the
    objective after 5 epochs and the parameters against the oracle (cd.nim:110-194) <= 1e-8"""
    data, idx, ptr, y, d = bench.gen_ml100k()
    n = len(y)
    csc = oracle.csr_to_csc(CSR(data, idx, ptr, n, d))
    ds = nf.newCSCDataset(csc.data, csc.indices, csc.indptr, n, d)
    kw = dict(alpha0=1e-10, alpha=1e-10, beta=1e-3)
    try:
        for degree in (2, 3):
            P = np.random.default_rng(1).standard_normal((degree - 1, 30, d)) * 0.01
            fm = nf.newFactorizationMachine(nf.regression, degree=degree, nComponents=30, warmStart=True)
            fm.P, fm.w, fm.intercept, fm.isInitialized = P.copy(), np.zeros(d), 0.0, True
            opt = nf.newCD(maxIter=5, verbose=0, tol=0.0, **kw)
            opt.fit(ds, y, fm)
            ref = oracle.cd_fit(csc, y, P, np.zeros(d), 0.0, degree, "squared", max_iter=5, **kw)
            obj_dev = opt.history[-1][1] + opt.history[-1][2]
            obj_ref = ref["loss"][-1] + ref["reg"][-1]
            assert abs(obj_dev - obj_ref) <= 1e-8 * abs(obj_ref)
            np.testing.assert_allclose([h_[0] for h_ in opt.history], ref["viol"], rtol=1e-7)
            assert max_rel(fm.P, ref["P"]) <= 1e-7 and max_rel(fm.w, ref["w"]) <= 1e-7
    finally:
        ds.free()


def test_c5_ffm_full_shape(oracle):
    """C5 at its BASELINE size (10 M rows, 39 fields, one feature per field, 1 M features, rank 8): halves add
    up, a random sample matches the oracle"""
    n = int(os.environ.get("NIMFM_FULLSIZE_FFM_ROWS", 10_000_000))
    data, idx, ptr, fields, y = bench.gen_ffm_rows(n, 77)
    d = bench.D_FEATURES
    ds = nf.newCSRFieldDataset(data, idx, ptr, fields, n, d, 39)
    ds.set_targets(y)
    rng = np.random.default_rng(3)
    m = nf.newFieldAwareFactorizationMachine(nf.classification, nComponents=8, warmStart=True)
    m.P, m.w, m.intercept, m.isInitialized = rng.standard_normal((39, d, 8)) * 0.01, rng.standard_normal(d) * 0.01, 0.0, True
    lib, ctx = _lib.load(), _lib.ctx()
    h = m._to_device(ds)
    try:
        def grad(b, cnt, rows=None, mb=n):
            ls = C.c_double()
            ids = None if rows is None else _lib.i64(rows)
            _lib.check(lib.nimfm_ffm_loss_grad(ctx, h, ds.handle(), 2, 1.0, b, cnt if rows is None else len(rows),
                                               _lib.ptr(ids), mb, 1, 0, C.byref(ls)))
            gP = np.zeros_like(m.P)
            _lib.check(lib.nimfm_ffm_get_grads(ctx, h, _lib.ptr(gP), None, None))
            return ls.value, gP
        ls, gP = grad(0, n)
        lsa, gPa = grad(0, n // 2)
        lsb, gPb = grad(n // 2, n - n // 2)
        assert abs(ls - (lsa + lsb)) <= 1e-10 * abs(ls) and sums_agree(gPa + gPb, gP)
        rows = rng.integers(0, n, 300)
        lsr, gPr = grad(0, 0, rows=rows, mb=len(rows))
        csr = CSR(data, idx, ptr, n, d, fields=fields, n_fields=39)
        sub = oracle.csr_take_rows(csr, rows)
        sub.fields = np.concatenate([fields[ptr[r]:ptr[r + 1]] for r in rows]).astype(np.int64)
        sub.n_fields = 39
        ref = oracle.ffm_loss_grad(sub, y[rows], m.P, m.w, 0.0, "logistic")
        assert abs(lsr - ref["loss"]) <= 1e-10 * abs(ref["loss"]) and max_rel(gPr, ref["gP"]) <= 1e-9
        out = np.zeros(n)
        _lib.check(lib.nimfm_ffm_decision_function(ctx, h, ds.handle(), _lib.ptr(out)))
        assert max_rel(out[rows], oracle.ffm_decision_function(sub, m.P, m.w, 0.0)) <= 1e-10
    finally:
        lib.nimfm_ffm_free(ctx, h)
        ds.free()
