"""Parity at BASELINE.json's FULL sizes through size-independent properties (the oracle cannot run 10 M
rows in seconds): the C4 workload of bench.py -- 10 M rows x 39 nnz, 1 M features, HOFM degree 3 rank 32 --
and a C5-shaped FFM problem.  Each property ties the full-size device result either to itself computed a
second way (chunked host streaming, halves that must add up) or to the oracle on a random row sample."""
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
from helpers import max_rel  # noqa: E402

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
    assert max_rel(gPa + gPb, gP) <= 1e-9 and max_rel(gwa + gwb, gw) <= 1e-9
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
    assert abs(lsh.value - lsm) <= 1e-10 * abs(lsm) and max_rel(gPh, gPm) <= 1e-9


def test_c5_ffm_full_shape(oracle):
    """C5 row shape (39 fields, one feature per field, 1 M features, rank 8) at 1 M rows: halves add up,
    a random sample matches the oracle"""
    n = int(os.environ.get("NIMFM_FULLSIZE_FFM_ROWS", 1_000_000))
    data, idx, ptr, fields, y, d = bench_configs.gen_ffm_rows(n, 77)
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
        assert abs(ls - (lsa + lsb)) <= 1e-10 * abs(ls) and max_rel(gPa + gPb, gP) <= 1e-9
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
