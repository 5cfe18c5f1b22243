"""Parity of the coordinate-descent kernels (K7-K9) against the CPU oracle, through the C ABI.
Template of tests/test_cd.nim (n=50, d=6, k=4): fast CD vs a reference after 3 iterations for degree
2..4 x fitLower x fitLinear x fitIntercept, plus the ML-100K-shaped one-hot case where the
disjoint-column batching actually batches."""
import numpy as np
import pytest

import nimfm_b200 as nf
from nimfm_b200 import _lib
from oracle.oracle import CSR
from helpers import make_dense, make_fm_params, max_rel

pytestmark = pytest.mark.gpu
OBJ_TOL = 1e-8


def csc_ds(oracle, csr):
    csc = oracle.csr_to_csc(csr)
    return csc, nf.newCSCDataset(csc.data, csc.indices, csc.indptr, csr.n, csr.d)


def make_fm(degree, k, fit_lower, fit_linear, fit_intercept, P, w, b, task=nf.regression):
    fm = nf.newFactorizationMachine(task, degree=degree, nComponents=k, fitLower=fit_lower, fitLinear=fit_linear,
                                    fitIntercept=fit_intercept, warmStart=True)
    fm.P, fm.w, fm.intercept, fm.isInitialized = P.copy(), w.copy(), b, True
    return fm


def check(opt, fm, ref, degree):
    h = np.array(opt.history)
    np.testing.assert_allclose(h[:, 0], ref["viol"], rtol=1e-8)
    np.testing.assert_allclose(h[:, 1], ref["loss"], rtol=OBJ_TOL)
    # (atol: degree 4 / fitLower=none collapses P to ~1e-22, where reg ~1e-26 is rounding noise)
    np.testing.assert_allclose(h[:, 2], ref["reg"], rtol=OBJ_TOL, atol=1e-18)
    # objective after the fixed number of epochs (mean loss + regularization / n), <= 1e-8 relative
    obj_dev, obj_ref = h[-1, 1] + h[-1, 2], ref["loss"][-1] + ref["reg"][-1]
    assert abs(obj_dev - obj_ref) <= OBJ_TOL * abs(obj_ref)
    np.testing.assert_allclose(fm.P, ref["P"], rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(fm.w, ref["w"], rtol=1e-8, atol=1e-12)
    assert abs(fm.intercept - ref["intercept"]) <= 1e-9 * max(1.0, abs(ref["intercept"]))


@pytest.mark.parametrize("degree", [2, 3, 4])
@pytest.mark.parametrize("fit_lower", ["explicit", "none", "augment"])
@pytest.mark.parametrize("fit_linear,fit_intercept", [(True, True), (False, True), (True, False), (False, False)])
def test_cd_matches_oracle(oracle, degree, fit_lower, fit_linear, fit_intercept):
    n, d, k = 50, 6, 4                       # tests/test_cd.nim:10-13
    X = make_dense(n, d, 31 + degree, density=0.7, positive=False)
    y = np.random.default_rng(degree).standard_normal(n)
    csr = CSR.from_dense(X)
    csc, ds = csc_ds(oracle, csr)
    P, w, nA = make_fm_params(d, degree, k, fit_lower, fit_linear, seed=9, scale=0.1)
    kw = dict(alpha0=1e-6, alpha=1e-3, beta=1e-3)
    ref = oracle.cd_fit(csc, y, P, w, 0.0, degree, "squared", fit_linear, fit_intercept, max_iter=3, **kw)
    fm = make_fm(degree, k, fit_lower, fit_linear, fit_intercept, P, w, 0.0)
    opt = nf.newCD(maxIter=3, verbose=0, tol=0.0, **kw)
    opt.fit(ds, y, fm)
    check(opt, fm, ref, degree)


@pytest.mark.parametrize("loss_name", ["logistic", "squared_hinge"])
def test_cd_classification_losses(oracle, loss_name):
    n, d, k, degree = 60, 7, 3, 3
    X = make_dense(n, d, 5, density=0.6, positive=False)
    y = np.sign(np.random.default_rng(1).standard_normal(n))
    csr = CSR.from_dense(X)
    csc, ds = csc_ds(oracle, csr)
    P, w, _ = make_fm_params(d, degree, k, "explicit", True, seed=3, scale=0.1)
    ref = oracle.cd_fit(csc, y, P, w, 0.0, degree, loss_name, True, True, max_iter=3, alpha0=1e-6, alpha=1e-3, beta=1e-3)
    fm = make_fm(degree, k, "explicit", True, True, P, w, 0.0, task=nf.classification)
    loss = nf.Logistic() if loss_name == "logistic" else nf.SquaredHinge()
    opt = nf.newCD(maxIter=3, verbose=0, tol=0.0, loss=loss, alpha0=1e-6, alpha=1e-3, beta=1e-3)
    opt.fit(ds, y, fm)
    check(opt, fm, ref, degree)


def one_hot_user_item(n, n_users, n_items, seed):
    """ML-100K shape (BASELINE configs 1-2): row i = {user u_i, n_users + item v_i}, values 1.0"""
    rng = np.random.default_rng(seed)
    u = rng.integers(0, n_users, n)
    v = rng.integers(0, n_items, n) + n_users
    idx = np.stack([u, v], axis=1)
    csr = CSR(np.ones(2 * n), idx.ravel(), np.arange(n + 1) * 2, n, n_users + n_items)
    y = rng.integers(1, 6, n).astype(np.float64)
    return csr, y


@pytest.mark.parametrize("degree", [2, 3])
def test_cd_one_hot_batched_columns(oracle, degree):
    """user columns are pairwise row-disjoint and so are item columns: two batches per sweep; results
    must still equal the strictly sequential reference order"""
    n, nu, ni, k = 3000, 60, 90, 8
    csr, y = one_hot_user_item(n, nu, ni, 7)
    csc, ds = csc_ds(oracle, csr)
    rng = np.random.default_rng(2)
    P = rng.standard_normal((degree - 1, k, nu + ni)) * 0.01
    w = np.zeros(nu + ni)
    kw = dict(alpha0=1e-10, alpha=1e-10, beta=1e-3)    # benchmarks/ml100k/factorization_machine.nim:13-14
    ref = oracle.cd_fit(csc, y, P, w, 0.0, degree, "squared", True, True, max_iter=3, **kw)
    fm = make_fm(degree, k, "explicit", True, True, P, w, 0.0)
    opt = nf.newCD(maxIter=3, verbose=0, tol=0.0, **kw)
    opt.fit(ds, y, fm)
    check(opt, fm, ref, degree)
    # warm start == continuing (tests/test_cd.nim:58-90): 3 more iterations from the fitted model
    ref2 = oracle.cd_fit(csc, y, ref["P"], ref["w"], ref["intercept"], degree, "squared", True, True, max_iter=2, **kw)
    opt2 = nf.newCD(maxIter=2, verbose=0, tol=0.0, **kw)
    opt2.fit(ds, y, fm)
    np.testing.assert_allclose(fm.P, ref2["P"], rtol=1e-7, atol=1e-11)
    np.testing.assert_allclose(np.array(opt2.history)[:, 1], ref2["loss"], rtol=OBJ_TOL)


def test_cd_callback_and_tol(oracle):
    n, d, k, degree = 40, 5, 2, 2
    X = make_dense(n, d, 3, density=0.8)
    y = np.random.default_rng(0).standard_normal(n)
    csc, ds = csc_ds(oracle, CSR.from_dense(X))
    P, w, _ = make_fm_params(d, degree, k, "explicit", True, seed=1)
    fm = make_fm(degree, k, "explicit", True, True, P, w, 0.0)
    seen = []
    opt = nf.newCD(maxIter=50, verbose=0, tol=1.0)
    opt.fit(ds, y, fm, callback=lambda o, m: seen.append(m.P.copy()))
    assert 1 <= len(opt.history) < 50 and opt.history[-1][0] < 1.0      # stopped by viol < tol (cd.nim:186-189)
    assert len(seen) == len(opt.history) and np.allclose(seen[-1], fm.P)
    with pytest.raises(TypeError):
        opt.fit(nf.newCSRDataset([1.0], [0], [0, 1], 1, 5), [1.0], fm)


# ---------------------------------------------------------------- PCD (optimizer/pcd.nim)
def make_reg(reg):
    return {"l1": nf.newL1, "squaredl12": nf.newSquaredL12,
            "squaredl12_rows": lambda: nf.newSquaredL12(transpose=False)}[reg]()


def check_pcd(opt, fm, ref, gamma, reg):
    h = np.array(opt.history)
    np.testing.assert_allclose(h[:, 0], ref["viol"], rtol=1e-8)
    np.testing.assert_allclose(h[:, 1], ref["loss"], rtol=OBJ_TOL)
    # verbose=0: the host does not add gamma*reg.eval (pcd.nim:178-185 computes it only when printing);
    # compare the full objective by adding it here from the fitted parameters of the LAST epoch
    extra = sum(gamma * make_reg(reg).eval(np.asarray(fm.P[o]).T, 2) for o in range(fm.P.shape[0]))
    obj_dev, obj_ref = h[-1, 1] + h[-1, 2] + extra, ref["loss"][-1] + ref["reg"][-1]
    assert abs(obj_dev - obj_ref) <= OBJ_TOL * abs(obj_ref)
    np.testing.assert_allclose(fm.P, ref["P"], rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(fm.w, ref["w"], rtol=1e-8, atol=1e-12)
    assert np.array_equal(fm.P == 0.0, ref["P"] == 0.0)            # the same coordinates are exactly zero
    assert abs(fm.intercept - ref["intercept"]) <= 1e-9 * max(1.0, abs(ref["intercept"]))


@pytest.mark.parametrize("degree,fit_lower,reg", [(2, "explicit", "squaredl12"), (2, "augment", "squaredl12"),
                                                  (2, "explicit", "squaredl12_rows"), (2, "explicit", "l1"),
                                                  (3, "explicit", "l1"), (3, "augment", "l1"), (4, "none", "l1")])
@pytest.mark.parametrize("fit_linear,fit_intercept", [(True, True), (False, False)])
def test_pcd_matches_oracle(oracle, degree, fit_lower, reg, fit_linear, fit_intercept):
    n, d, k = 50, 6, 4
    X = make_dense(n, d, 41 + degree, density=0.7, positive=False)
    y = np.random.default_rng(degree).standard_normal(n)
    csr = CSR.from_dense(X)
    csc, ds = csc_ds(oracle, csr)
    P, w, nA = make_fm_params(d, degree, k, fit_lower, fit_linear, seed=19, scale=0.1)
    gamma = 2e-2
    kw = dict(alpha0=1e-6, alpha=1e-3, beta=1e-3, gamma=gamma)
    ref = oracle.pcd_fit(csc, y, P, w, 0.0, degree, "squared", fit_linear, fit_intercept, max_iter=4, reg=reg, **kw)
    assert np.count_nonzero(ref["P"] == 0.0) > 0                    # the prox acts
    fm = make_fm(degree, k, fit_lower, fit_linear, fit_intercept, P, w, 0.0)
    opt = nf.newPCD(maxIter=4, verbose=0, tol=0.0, reg=make_reg(reg), **kw)
    opt.fit(ds, y, fm)
    if degree == 2:
        check_pcd(opt, fm, ref, gamma, reg)
    else:
        h = np.array(opt.history)
        np.testing.assert_allclose(h[:, 0], ref["viol"], rtol=1e-8)
        np.testing.assert_allclose(h[:, 1], ref["loss"], rtol=OBJ_TOL)
        np.testing.assert_allclose(fm.P, ref["P"], rtol=1e-8, atol=1e-12)
        assert np.array_equal(fm.P == 0.0, ref["P"] == 0.0)


@pytest.mark.parametrize("reg", ["squaredl12", "squaredl12_rows", "l1"])
def test_pcd_one_hot_batched_columns(oracle, reg):
    """ML-100K shape: the chained SquaredL12 prox resolves each batch of row-disjoint columns in the
    reference's sequential order (the running |P[s]|_1 cache links every coordinate of a sweep)"""
    n, nu, ni, k = 3000, 60, 90, 8
    csr, y = one_hot_user_item(n, nu, ni, 17)
    csc, ds = csc_ds(oracle, csr)
    rng = np.random.default_rng(3)
    P = rng.standard_normal((1, k, nu + ni)) * 0.05
    w = np.zeros(nu + ni)
    gamma = 5e-3
    kw = dict(alpha0=1e-10, alpha=1e-10, beta=1e-3, gamma=gamma)
    ref = oracle.pcd_fit(csc, y, P, w, 0.0, 2, "squared", True, True, max_iter=3, reg=reg, **kw)
    fm = make_fm(2, k, "explicit", True, True, P, w, 0.0)
    opt = nf.newPCD(maxIter=3, verbose=0, tol=0.0, reg=make_reg(reg), **kw)
    opt.fit(ds, y, fm)
    check_pcd(opt, fm, ref, gamma, reg)


def test_pcd_default_regulariser_rules(oracle):
    X = make_dense(20, 5, 1)
    csc, ds = csc_ds(oracle, CSR.from_dense(X))
    P, w, _ = make_fm_params(5, 3, 2, "explicit", True, seed=1)
    fm = make_fm(3, 2, "explicit", True, True, P, w, 0.0)
    with pytest.raises(ValueError, match="SquaredL12 supports only degree=2"):   # squaredl12.nim:90-93
        nf.newPCD(maxIter=1, verbose=0).fit(ds, np.zeros(20), fm)
    with pytest.raises(TypeError):
        nf.newPCD(maxIter=1, verbose=0, reg=nf.newL21()).fit(ds, np.zeros(20), fm)
