"""Pins oracle/ref_cpu.c against the brute-force definitions restated from the reference's own
test helpers (tests/kernels_slow.nim, tests/model/*_slow.nim, tests/optimizer/*_slow.nim), on the
reference's test shapes and tolerances.  CPU only."""
import numpy as np
import pytest

from oracle import bruteforce as bf
from oracle.oracle import CSR
from helpers import make_dense, make_fm_params, make_field_csr


# tests/test_kernels.nim:7-46: n=20, d=10, k=10, degrees 2..5, 0..3 dummies, |fast-slow| < 1e-6
@pytest.mark.parametrize("is_csc", [False, True])
def test_anova_vs_subset_enumeration(oracle, is_csc):
    n, d, k, n_aug_max = 20, 10, 10, 3
    X = make_dense(n, d, 42)
    rng = np.random.default_rng(7)
    P = rng.standard_normal((k, d + n_aug_max))
    csr = CSR.from_dense(X)
    mat = oracle.csr_to_csc(csr) if is_csc else csr
    for m in range(n_aug_max + 1):
        for degree in range(2, 6):
            for s in range(0, k, 3):
                Ps = P[s, :d + m]
                A = oracle.anova(mat, Ps, degree, n_aug=m, is_csc=is_csc)
                for i in range(0, n, 4):
                    expect = bf.anova_slow(X[i], Ps, degree, d, m)
                    assert abs(A[i, degree] - expect) < 1e-6 * max(1.0, abs(expect))


# model/fm_slow.nim:42-60 -- decisionFunction incl. dummy features, all fitLower kinds
@pytest.mark.parametrize("degree", [2, 3, 4])
@pytest.mark.parametrize("fit_lower", ["explicit", "none", "augment"])
def test_fm_decision_function_vs_bruteforce(oracle, degree, fit_lower):
    n, d, k = 12, 6, 4
    X = make_dense(n, d, 3, density=0.7)
    P, w, nA = make_fm_params(d, degree, k, fit_lower, True, seed=degree)
    csr = CSR.from_dense(X)
    expect = bf.fm_decision_function(X, P, w, 0.3, degree)
    got_r = oracle.fm_decision_function(csr, P, w, 0.3, degree)
    got_c = oracle.fm_decision_function(oracle.csr_to_csc(csr), P, w, 0.3, degree, is_csc=True)
    np.testing.assert_allclose(got_r, expect, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(got_c, expect, rtol=1e-10, atol=1e-12)


# model/fm_slow.nim:111-134 -- gradient by combNotj enumeration vs sgd.nim:176-188 recurrences
@pytest.mark.parametrize("degree", [2, 3, 4, 5])
@pytest.mark.parametrize("fit_lower", ["explicit", "none", "augment"])
def test_fm_grad_vs_bruteforce(oracle, degree, fit_lower):
    n, d, k = 6, 6, 3
    X = make_dense(n, d, 11, density=0.8)
    y = np.random.default_rng(0).standard_normal(n)
    P, w, nA = make_fm_params(d, degree, k, fit_lower, True, seed=degree + 10)
    csr = CSR.from_dense(X)
    res = oracle.fm_loss_grad(csr, y, P, w, 0.1, degree, "squared", mini_batch_size=1)
    yhat = bf.fm_decision_function(X, P, w, 0.1, degree)
    np.testing.assert_allclose(res["y_pred"], yhat, rtol=1e-10, atol=1e-12)
    grad = np.zeros_like(P)
    gw = np.zeros(d)
    gb = 0.0
    for i in range(n):
        dL = yhat[i] - y[i]
        bf.fm_grad(X, i, P, degree, dL, grad)
        gw += dL * X[i]
        gb += dL
    np.testing.assert_allclose(res["gP"], grad, rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(res["gw"], gw, rtol=1e-10, atol=1e-12)
    assert abs(res["gb"] - gb) < 1e-10


# model/ffm_slow.nim:38-56,110-127; shape of tests/test_sgd_ffm.nim:10-14 (d=20, 5 fields, k=4)
def test_ffm_vs_bruteforce(oracle):
    n, d, nF, k = 10, 20, 5, 4
    X, csr, field_of = make_field_csr(n, d, nF, 5)
    rng = np.random.default_rng(1)
    P = rng.standard_normal((nF, d, k)) * 0.3
    w = rng.standard_normal(d) * 0.1
    y = rng.standard_normal(n)
    expect = bf.ffm_decision_function(X, field_of, P, w, -0.2)
    got = oracle.ffm_decision_function(csr, P, w, -0.2)
    np.testing.assert_allclose(got, expect, rtol=1e-10, atol=1e-12)
    res = oracle.ffm_loss_grad(csr, y, P, w, -0.2, "squared", mini_batch_size=1)
    np.testing.assert_allclose(res["y_pred"], expect, rtol=1e-10, atol=1e-12)
    grad = np.zeros_like(P)
    for i in range(n):
        bf.ffm_grad(X, field_of, i, P, expect[i] - y[i], grad)
    np.testing.assert_allclose(res["gP"], grad, rtol=1e-9, atol=1e-12)


# tests/test_cd.nim:93-128 -- fast CD vs CDSlow after 3 iterations (rtol 1e-6 / atol 1e-9)
@pytest.mark.parametrize("degree,fit_lower", [(2, "explicit"), (3, "explicit"), (3, "augment"),
                                              (3, "none"), (4, "explicit")])
@pytest.mark.parametrize("fit_linear,fit_intercept", [(True, True), (False, False)])
def test_cd_vs_slow(oracle, degree, fit_lower, fit_linear, fit_intercept):
    n, d, k = 14, 5, 2
    X = make_dense(n, d, 21, density=0.8)
    rng = np.random.default_rng(2)
    y = rng.standard_normal(n)
    P, w, nA = make_fm_params(d, degree, k, fit_lower, fit_linear, seed=5)
    csc = oracle.csr_to_csc(CSR.from_dense(X))
    fast = oracle.cd_fit(csc, y, P, w, 0.0, degree, "squared", fit_linear, fit_intercept,
                         max_iter=2, alpha0=1e-6, alpha=1e-3, beta=1e-3)
    Ps, ws, bs = bf.cd_slow_fit(X, y, P, w, 0.0, degree, fit_linear, fit_intercept, "squared", 2,
                                1e-6, 1e-3, 1e-3)
    np.testing.assert_allclose(fast["P"], Ps, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(fast["w"], ws, rtol=1e-6, atol=1e-9)
    assert abs(fast["intercept"] - bs) < 1e-7
    # the cached predictions must equal a from-scratch forward with the final parameters
    yp = oracle.fm_decision_function(csc, fast["P"], fast["w"], fast["intercept"], degree, is_csc=True)
    np.testing.assert_allclose(fast["y_pred"], yp, rtol=1e-9, atol=1e-11)


# tests/test_adagrad.nim:92-126 -- AdaGrad vs AdaGradSlow (5 epochs in the reference; 2 here)
@pytest.mark.parametrize("degree,fit_lower", [(2, "explicit"), (3, "explicit"), (3, "augment"), (4, "none")])
def test_adagrad_vs_slow(oracle, degree, fit_lower):
    n, d, k = 10, 5, 2
    X = make_dense(n, d, 31, density=0.7)
    y = np.random.default_rng(3).standard_normal(n)
    P, w, nA = make_fm_params(d, degree, k, fit_lower, True, seed=8)
    w = np.zeros(d)
    csr = CSR.from_dense(X)
    fast = oracle.adagrad_fit(csr, y, P, w, 0.0, degree, "squared", True, True, max_iter=2)
    Ps, ws, bs = bf.adagrad_slow_fit(X, y, P, w, 0.0, degree, True, True, "squared", 2, 0.1, 1e-6,
                                     1e-3, 1e-3, 1e-10)
    np.testing.assert_allclose(fast["P"], Ps, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(fast["w"], ws, rtol=1e-6, atol=1e-9)
    assert abs(fast["intercept"] - bs) < 1e-7


@pytest.mark.parametrize("degree,fit_lower", [(2, "explicit"), (3, "explicit"), (3, "augment")])
@pytest.mark.parametrize("B", [1, 4, 16])
@pytest.mark.parametrize("fit_linear,fit_intercept", [(True, True), (False, True), (True, False)])
def test_adagrad_minibatch_restatement_vs_naive_dense_definition(oracle, degree, fit_lower, B, fit_linear, fit_intercept):
    """oracle.adagrad_fit with miniBatchSize = B (lazy refresh of the minibatch's features from g_sum / g_norm, CSR,
    per-sample squares -- the definition the device's synchronous-minibatch AdaGrad is held to) against the naive
    dense definition bruteforce.adagrad_minibatch_slow_fit (every parameter refreshed before every minibatch);
    at B = 1 the latter is the reference's AdaGradSlow (tests/optimizer/adagrad_slow.nim:29-102)."""
    n, d, k = 37, 6, 3
    X = make_dense(n, d, 41, density=0.5, positive=False)
    y = np.sign(np.random.default_rng(6).standard_normal(n))
    csr = CSR.from_dense(X)
    P, w, _ = make_fm_params(d, degree, k, fit_lower, fit_linear, seed=2, scale=0.2)
    w = np.zeros(d)
    kw = dict(eta0=0.1, alpha0=1e-3, alpha=1e-2, beta=2e-2, eps=1e-10)
    got = oracle.adagrad_fit(csr, y, P, w, 0.0, degree, "logistic", fit_linear, fit_intercept, max_iter=3,
                             mini_batch_size=B, **kw)
    sP, sw, sb = bf.adagrad_minibatch_slow_fit(X, y, P, w, 0.0, degree, fit_linear, fit_intercept, "logistic", B, 3,
                                               kw["eta0"], kw["alpha0"], kw["alpha"], kw["beta"], kw["eps"])
    np.testing.assert_allclose(got["P"], sP, rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(got["w"], sw, rtol=1e-9, atol=1e-13)
    assert abs(got["intercept"] - sb) <= 1e-12
    if B == 1:
        qP, qw, qb = bf.adagrad_slow_fit(X, y, P, w, 0.0, degree, fit_linear, fit_intercept, "logistic", 3, kw["eta0"],
                                         kw["alpha0"], kw["alpha"], kw["beta"], kw["eps"])
        np.testing.assert_allclose(sP, qP, rtol=1e-9, atol=1e-13)
        np.testing.assert_allclose(sw, qw, rtol=1e-9, atol=1e-13)


def test_losses_match_the_definitions(oracle):
    """loss.nim:15-102 restated in C (ref_loss / ref_dloss / ref_mu) against the closed forms, and dloss against a
    centred finite difference of loss for the three smooth losses.  Huber is held to the reference AS WRITTEN:
    dloss returns y - p inside the threshold and +threshold outside (loss.nim:90-93), which is not the derivative of
    its loss -- the device reproduces that, so the quirk is pinned here rather than corrected."""
    rng = np.random.default_rng(0)
    ys = np.concatenate([np.sign(rng.standard_normal(60)), rng.standard_normal(60) * 2])
    ps = rng.standard_normal(120) * 3
    for kind in ("squared", "squared_hinge", "logistic", "huber"):
        for yv, pv in zip(ys, ps):
            assert abs(oracle.loss(kind, yv, pv) - bf.loss_val(kind, yv, pv)) <= 1e-14 * max(1.0, abs(bf.loss_val(kind, yv, pv)))
            assert abs(oracle.dloss(kind, yv, pv) - bf.dloss_val(kind, yv, pv)) <= 1e-14 * max(1.0, abs(bf.dloss_val(kind, yv, pv)))
            if kind != "huber" and not (kind == "squared_hinge" and abs(1 - pv * yv) < 1e-3):
                h = 1e-6
                fd = (bf.loss_val(kind, yv, pv + h) - bf.loss_val(kind, yv, pv - h)) / (2 * h)
                assert abs(oracle.dloss(kind, yv, pv) - fd) <= 1e-6 * max(1.0, abs(fd))
    # the quirk itself: inside the threshold the sign is flipped, outside it is +threshold on both sides
    assert oracle.dloss("huber", 0.0, 0.5) == -0.5 and oracle.dloss("huber", 0.0, -0.5) == 0.5
    assert oracle.dloss("huber", 0.0, 3.0) == 1.0 and oracle.dloss("huber", 0.0, -3.0) == 1.0
    assert oracle.loss("huber", 0.0, 3.0) == 2.5 and oracle.loss("huber", 0.0, 0.5) == 0.125
    # large margins: the two branches of the logistic loss stay finite
    assert oracle.loss("logistic", 1.0, 800.0) == 0.0 and abs(oracle.loss("logistic", 1.0, -800.0) - 800.0) < 1e-9
    assert oracle.dloss("logistic", 1.0, 800.0) == 0.0 and oracle.dloss("logistic", 1.0, -800.0) == -1.0


# tests/test_sgd_ffm.nim / test_adagrad_ffm.nim "Comparison to naive implementation" (n=80, d=20, 5 fields, k=4)
@pytest.mark.parametrize("fit_linear", [False, True])
@pytest.mark.parametrize("fit_intercept", [False, True])
def test_ffm_sgd_vs_slow(oracle, fit_linear, fit_intercept):
    """oracle.ffm_sgd_fit (sgd_ffm.nim: lazy L2 scaling) against the reference's naive SGDSlow for FFMSlow
    (tests/optimizer/sgd_ffm_slow.nim), as tests/test_sgd_ffm.nim does"""
    n, d, nF, k = 40, 12, 3, 3
    X, csr, field_of = make_field_csr(n, d, nF, 14)
    y = np.random.default_rng(3).standard_normal(n)
    rng = np.random.default_rng(11)
    P, w = rng.standard_normal((nF, d, k)) * 0.1, np.zeros(d)
    kw = dict(eta0=0.05, alpha0=1e-3, alpha=1e-2, beta=2e-2)
    got = oracle.ffm_sgd_fit(csr, y, P, w, 0.0, "squared", fit_linear, fit_intercept, max_iter=3, it=1, **kw)
    sP, sw, sb = bf.ffm_sgd_slow_fit(X, field_of, y, P, w, 0.0, fit_linear, fit_intercept, "squared", 3, kw["eta0"],
                                     kw["alpha0"], kw["alpha"], kw["beta"])
    np.testing.assert_allclose(got["P"], sP, rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(got["w"], sw, rtol=1e-8, atol=1e-12)
    assert abs(got["intercept"] - sb) <= 1e-10
    if not fit_linear:
        assert np.all(got["w"] == 0.0)
    if not fit_intercept:
        assert got["intercept"] == 0.0


@pytest.mark.parametrize("fit_linear", [False, True])
@pytest.mark.parametrize("fit_intercept", [False, True])
def test_ffm_adagrad_vs_slow(oracle, fit_linear, fit_intercept):
    """oracle.ffm_adagrad_fit (adagrad_ffm.nim: lazy refresh of the row's features over all fields) against the
    reference's naive AdaGradSlow for FFMSlow (tests/optimizer/adagrad_ffm_slow.nim), as tests/test_adagrad_ffm.nim
    does (atol 1e-7 there)"""
    n, d, nF, k = 40, 12, 3, 3
    X, csr, field_of = make_field_csr(n, d, nF, 15)
    y = np.random.default_rng(4).standard_normal(n)
    rng = np.random.default_rng(12)
    P, w = rng.standard_normal((nF, d, k)) * 0.1, np.zeros(d)
    kw = dict(eta0=0.1, alpha0=1e-3, alpha=1e-2, beta=2e-2, eps=1e-10)
    got = oracle.ffm_adagrad_fit(csr, y, P, w, 0.0, "squared", fit_linear, fit_intercept, max_iter=3, **kw)
    sP, sw, sb = bf.ffm_adagrad_slow_fit(X, field_of, y, P, w, 0.0, fit_linear, fit_intercept, "squared", 3, kw["eta0"],
                                         kw["alpha0"], kw["alpha"], kw["beta"], kw["eps"])
    np.testing.assert_allclose(got["P"], sP, rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(got["w"], sw, rtol=1e-7, atol=1e-9)
    assert abs(got["intercept"] - sb) <= 1e-8


# tests/test_sgd.nim:92-126 -- SGD (lazy scaling) vs SGDSlow (dense shrink)
@pytest.mark.parametrize("degree,fit_lower", [(2, "explicit"), (3, "explicit"), (3, "augment"), (4, "none")])
def test_sgd_vs_slow(oracle, degree, fit_lower):
    n, d, k = 10, 5, 2
    X = make_dense(n, d, 41, density=0.5)
    y = np.random.default_rng(4).standard_normal(n)
    P, w, nA = make_fm_params(d, degree, k, fit_lower, True, seed=9)
    csr = CSR.from_dense(X)
    fast = oracle.sgd_fit(csr, y, P, w, 0.0, degree, "squared", True, True, max_iter=2)
    Ps, ws, bs = bf.sgd_slow_fit(X, y, P, w, 0.0, degree, True, True, "squared", 2, 0.01, 1e-6,
                                 1e-3, 1e-3)
    np.testing.assert_allclose(fast["P"], Ps, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(fast["w"], ws, rtol=1e-6, atol=1e-9)
    assert abs(fast["intercept"] - bs) < 1e-7


# tests/test_dataset.nim:6-52 style: CSR<->CSC round trip + nnz equality + stable ordering
def test_transpose_bookkeeping(oracle):
    X = make_dense(17, 9, 51, density=0.4)
    csr = CSR.from_dense(X)
    csc = oracle.csr_to_csc(csr)
    assert csc.indptr[-1] == csr.indptr[-1] == np.count_nonzero(X)
    for j in range(9):
        rows = csc.indices[csc.indptr[j]:csc.indptr[j + 1]]
        assert np.all(np.diff(rows) > 0)           # stable in row order
        np.testing.assert_array_equal(rows, np.nonzero(X[:, j])[0])
    back = oracle.csc_to_csr(csc)
    np.testing.assert_array_equal(back.indptr, csr.indptr)
    np.testing.assert_array_equal(back.indices, csr.indices)
    np.testing.assert_array_equal(back.data, csr.data)
    sub = oracle.csr_take_rows(csr, [3, 0, 16])
    np.testing.assert_array_equal(sub.to_dense(), X[[3, 0, 16]])
    sl = oracle.csc_slice_rows(csc, 4, 11)
    np.testing.assert_array_equal(oracle.csc_to_csr(sl).to_dense(), X[4:12])


def test_mbpsgd_reduces_to_full_gradient_step(oracle):
    """With mb == n, maxIterInner == 1, one MBPSGD epoch is one (P - eta*g)/(1+eta*beta) step on the
    mean gradient (minibatch_psgd.nim:96-122, params.nim:90-98)."""
    n, d, k, degree = 9, 6, 3, 3
    X = make_dense(n, d, 61, density=0.6)
    y = np.sign(np.random.default_rng(5).standard_normal(n))
    P, w, nA = make_fm_params(d, degree, k, "explicit", True, seed=3)
    csr = CSR.from_dense(X)
    res = oracle.mbpsgd_fit(csr, y, P, w, 0.05, degree, "logistic", max_iter=1, mini_batch_size=n,
                            eta0=0.1, alpha0=1e-6, alpha=1e-3, beta=1e-4)
    g = oracle.fm_loss_grad(csr, y, P, w, 0.05, degree, "logistic", mini_batch_size=n)
    eta_P = 0.1 / (1 + 0.1 * 1e-4 * 1)
    eta_w = 0.1 / (1 + 0.1 * 1e-3 * 1)
    eta_b = 0.1 / (1 + 0.1 * 1e-6 * 1)
    np.testing.assert_allclose(res["P"], (P - eta_P * g["gP"]) / (1 + eta_P * 1e-4), rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(res["w"], (w - eta_w * g["gw"]) / (1 + eta_w * 1e-3), rtol=1e-12, atol=1e-15)
    assert abs(res["intercept"] - (0.05 - eta_b * g["gb"]) / (1 + eta_b * 1e-6)) < 1e-14
    assert abs(res["epoch_loss"][0] - g["loss"] / n) < 1e-14
    assert res["it"] == 2


@pytest.mark.parametrize("degree,fit_lower", [(2, "explicit"), (3, "explicit"), (3, "augment")])
@pytest.mark.parametrize("reg,gamma", [("identity", 0.0), ("l1", 0.05), ("l21", 0.05), ("squaredl12", 0.02),
                                       ("squaredl12_rows", 0.02)])
@pytest.mark.parametrize("mb,fit_linear", [(-1, True), (5, True), (4, False)])
def test_mbpsgd_restatement_vs_naive_dense_definition(oracle, degree, fit_lower, reg, gamma, mb, fit_linear):
    """oracle.mbpsgd_fit (line-by-line restatement of minibatch_psgd.nim:67-211 on CSR data: per-sample scatter,
    Params.step, the regulariser's prox) against bruteforce.mbpsgd_slow_fit, the naive dense definition (subset
    enumeration for yhat and its gradient, mean minibatch gradient, dense step, closed-form prox): iterates and
    epoch losses over three epochs, default and explicit minibatch sizes (a ragged last wrap included: 37 rows),
    every MBPSGD regulariser, and the intercept quirk of params.nim:47 with fitLinear = false."""
    n, d, k = 37, 6, 3
    X = make_dense(n, d, 35, density=0.5, positive=False)
    y = np.sign(np.random.default_rng(9).standard_normal(n))
    csr = CSR.from_dense(X)
    P, w, _ = make_fm_params(d, degree, k, fit_lower, fit_linear, seed=5, scale=0.3)
    kw = dict(eta0=0.2, alpha0=1e-3, alpha=1e-2, beta=2e-2)
    got = oracle.mbpsgd_fit(csr, y, P, w, 0.1, degree, "logistic", fit_linear, True, max_iter=3, gamma=gamma, reg=reg,
                            mini_batch_size=mb, **kw)
    sP, sw, sb, sl = bf.mbpsgd_slow_fit(X, y, P, w, 0.1, degree, fit_linear, True, "logistic", 3, kw["eta0"],
                                        kw["alpha0"], kw["alpha"], kw["beta"], gamma=gamma, reg=reg, mini_batch_size=mb)
    np.testing.assert_allclose(got["P"], sP, rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(got["w"], sw, rtol=1e-9, atol=1e-13)
    assert abs(got["intercept"] - sb) <= 1e-12
    np.testing.assert_allclose(got["epoch_loss"], sl, rtol=1e-10)
    if reg != "identity":
        assert np.array_equal(got["P"] == 0.0, sP == 0.0)      # the prox leaves the same zero pattern


# ---------------------------------------------------------------- proximal operators
@pytest.mark.parametrize("n,lam", [(1, 0.3), (7, 0.05), (50, 0.5), (200, 1e-3), (64, 10.0)])
def test_prox_squaredl12_matches_closed_form_and_optimality(oracle, n, lam):
    """proxSquaredL12 (squaredl12.nim:16-64) restated with a deterministic pivot stream == the sort-based
    closed form, and satisfies the optimality condition q = softthreshold(p, 2*lam*||q||_1)."""
    rng = np.random.default_rng(n)
    for trial in range(5):
        p = rng.standard_normal(n) * (rng.random(n) < 0.8)
        if trial == 4:
            p[: n // 2] = p[0]                          # ties
        q = oracle.prox_matrix(p.reshape(-1, 1), lam, "squaredl12").ravel()
        np.testing.assert_allclose(q, bf.prox_squaredl12_sorted(p, lam), rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(q, np.sign(p) * np.maximum(np.abs(p) - 2 * lam * np.sum(np.abs(q)), 0.0),
                                   rtol=1e-10, atol=1e-14)
        # transpose=false applies the same operator to every row
        M = rng.standard_normal((3, n))
        R = oracle.prox_matrix(M, lam, "squaredl12_rows")
        for j in range(3):
            np.testing.assert_allclose(R[j], bf.prox_squaredl12_sorted(M[j], lam), rtol=1e-12, atol=1e-15)


def test_prox_l21_and_l1_and_eval(oracle):
    rng = np.random.default_rng(9)
    M = rng.standard_normal((6, 5))
    M[2] *= 1e-3
    R = oracle.prox_matrix(M, 0.1, "l21")
    for j in range(6):
        np.testing.assert_allclose(R[j], bf.prox_l21_row(M[j], 0.1), rtol=1e-14, atol=0)
    assert np.all(R[2] == 0.0)
    np.testing.assert_allclose(oracle.prox_matrix(M, 0.2, "l1"), np.sign(M) * np.maximum(np.abs(M) - 0.2, 0), rtol=1e-15)
    np.testing.assert_allclose(oracle.reg_eval(M, "l1"), np.abs(M).sum(), rtol=1e-14)
    np.testing.assert_allclose(oracle.reg_eval(M, "l21"), np.sqrt((M * M).sum(1)).sum(), rtol=1e-14)
    np.testing.assert_allclose(oracle.reg_eval(M, "squaredl12"), (np.abs(M).sum(0) ** 2).sum(), rtol=1e-14)
    np.testing.assert_allclose(oracle.reg_eval(M, "squaredl12_rows"), (np.abs(M).sum(1) ** 2).sum(), rtol=1e-14)


def test_mbpsgd_with_squaredl12_decreases_and_sparsifies(oracle):
    """the default nimfm MBPSGD configuration (reg=SquaredL12, gamma>0, degree 2) runs in the oracle:
    a larger gamma gives a sparser P"""
    X = make_dense(60, 12, 3, density=0.6)
    csr = CSR.from_dense(X)
    y = np.random.default_rng(4).standard_normal(60)
    P, w, _ = make_fm_params(12, 2, 4, "explicit", True, 5)
    nz = []
    for gamma in (0.0, 1e-2, 1.0):
        r = oracle.mbpsgd_fit(csr, y, P, w, 0.0, 2, "squared", max_iter=5, gamma=gamma, reg="squaredl12",
                              mini_batch_size=10)
        nz.append(int(np.count_nonzero(r["P"])))
        assert np.all(np.isfinite(r["epoch_loss"]))
    assert nz[0] >= nz[1] >= nz[2] and nz[2] < nz[0]


# ---------------------------------------------------------------- PCD
def test_pcd_oracle_reduces_to_cd_and_minimises_coordinate_model(oracle):
    """pcd.nim with gamma = 0 is cd.nim (degree 2; for degree > 2 PCD adds the 1e-12 guard, which never
    fires here); with gamma > 0 the objective incl. gamma*reg decreases monotonically for the squared
    loss (each coordinate step minimises an upper bound of it)."""
    X = make_dense(50, 9, 21, density=0.5)
    csr = CSR.from_dense(X)
    csc = oracle.csr_to_csc(csr)
    y = np.random.default_rng(6).standard_normal(50)
    for degree, fl in ((2, "explicit"), (3, "explicit"), (3, "augment")):
        P, w, _ = make_fm_params(9, degree, 3, fl, True, 8)
        kw = dict(max_iter=4, alpha0=1e-6, alpha=1e-3, beta=1e-3)
        a = oracle.cd_fit(csc, y, P, w, 0.0, degree, "squared", **kw)
        for reg in ("l1", "squaredl12", "squaredl12_rows"):
            if reg != "l1" and degree != 2:
                continue
            b = oracle.pcd_fit(csc, y, P, w, 0.0, degree, "squared", gamma=0.0, reg=reg, **kw)
            np.testing.assert_allclose(b["P"], a["P"], rtol=1e-12, atol=1e-15)
            np.testing.assert_allclose(b["viol"], a["viol"], rtol=1e-12)
    P, w, _ = make_fm_params(9, 2, 3, "explicit", True, 8)
    for reg in ("l1", "squaredl12", "squaredl12_rows"):
        r = oracle.pcd_fit(csc, y, P, w, 0.0, 2, "squared", max_iter=8, gamma=5e-2, reg=reg, beta=1e-3)
        obj = r["loss"] + r["reg"]
        assert np.all(np.diff(obj) <= 1e-12), (reg, obj)
        assert np.count_nonzero(r["P"] == 0.0) > 0


@pytest.mark.parametrize("degree,fit_lower,reg", [(2, "explicit", "squaredl12"), (2, "augment", "squaredl12"),
                                                  (2, "none", "squaredl12_rows"), (2, "explicit", "l1"),
                                                  (3, "explicit", "l1")])
def test_pcd_oracle_matches_reference_slow_solver(oracle, degree, fit_lower, reg):
    """tests/test_pcd_squaredl12.nim / test_pcd_l1.nim compare pcd.nim with PCDSlow (pcd_slow.nim) at
    n=50, d=6, k=4; the same comparison pins the restatement (rtol 1e-6 / atol 1e-9 there)."""
    n, d, k = 30, 6, 3
    X = make_dense(n, d, 77, density=0.7, positive=False)
    y = np.random.default_rng(8).standard_normal(n)
    csc = oracle.csr_to_csc(CSR.from_dense(X))
    for fit_linear, fit_intercept in ((True, True), (False, False)):
        P, w, _ = make_fm_params(d, degree, k, fit_lower, fit_linear, 13, scale=0.3)
        kw = dict(alpha0=1e-6, alpha=1e-3, beta=1e-3, gamma=3e-2)
        r = oracle.pcd_fit(csc, y, P, w, 0.0, degree, "squared", fit_linear, fit_intercept, max_iter=2, reg=reg, **kw)
        Ps, ws, bs = bf.pcd_slow_fit(X, y, P, w, 0.0, degree, fit_linear, fit_intercept, "squared", 2, reg=reg, **kw)
        np.testing.assert_allclose(r["P"], Ps, rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(r["w"], ws, rtol=1e-6, atol=1e-9)
        assert abs(r["intercept"] - bs) < 1e-8
        assert np.count_nonzero(Ps == 0.0) > 0 and np.array_equal(r["P"] == 0.0, Ps == 0.0)


def test_prox_squaredl12_matches_reference_slow_definition(oracle):
    """tests/test_squaredl12.nim: proxSquaredL12 vs proxSquaredL12Slow, d = 100, values in [-2, 2], the
    reference's own lambda grid, tolerance 1e-10"""
    rng = np.random.default_rng(0)
    for _ in range(20):
        q = rng.integers(0, 401, 100) / 100 - 2.0
        for lam in [0.001, 0.002, 0.005, 0.01, 0.02, 0.05, 0.1, 0.2, 0.5, 1, 2, 3, 4]:
            p1 = oracle.prox_matrix(q.reshape(-1, 1), lam, "squaredl12").ravel()
            assert np.max(np.abs(p1 - bf.prox_squaredl12_slow(q, lam))) < 1e-10


# ---------------------------------------------------------------- PSGD
@pytest.mark.parametrize("degree,fit_lower,reg", [(2, "explicit", "l1"), (3, "explicit", "l1"), (3, "augment", "l21"),
                                                  (2, "none", "l21"), (2, "explicit", "squaredl12"),
                                                  (2, "augment", "squaredl12_rows")])
def test_psgd_oracle_vs_reference_slow_solver(oracle, degree, fit_lower, reg):
    """tests/test_psgd_{l1,l21,squaredl12}.nim compare psgd.nim (lazy bookkeeping) with the dense
    PSGDSlow at rtol 1e-6 / atol 1e-9 after a few epochs (tests/utils.nim:82-104); the same comparison
    pins the restatement.  (The lazy L1 / L21 protocols are close to, not identical with, an eager prox
    per step, hence the reference's tolerance.)"""
    n, d, k = 20, 6, 3
    X = make_dense(n, d, 55, density=0.6, positive=False)
    y = np.random.default_rng(9).standard_normal(n)
    csr = CSR.from_dense(X)
    for fit_linear, fit_intercept in ((True, True), (False, False)):
        P, w, _ = make_fm_params(d, degree, k, fit_lower, fit_linear, 17, scale=0.3)
        kw = dict(eta0=0.05, alpha0=1e-6, alpha=1e-3, beta=1e-4, gamma=2e-2)
        r = oracle.psgd_fit(csr, y, P, w, 0.0, degree, "squared", fit_linear, fit_intercept, max_iter=2, reg=reg, **kw)
        Ps, ws, bs = bf.psgd_slow_fit(X, y, P, w, 0.0, degree, fit_linear, fit_intercept, "squared", 2, reg=reg, **kw)
        np.testing.assert_allclose(r["P"], Ps, rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(r["w"], ws, rtol=1e-4, atol=1e-6)
        assert abs(r["intercept"] - bs) < 1e-6
        assert r["it"] == 1 + 2 * n
    # the prox acts: a larger gamma zeroes parameters
    r = oracle.psgd_fit(csr, y, P, w, 0.0, degree, "squared", max_iter=2, reg=reg, eta0=0.05, gamma=0.5)
    assert np.count_nonzero(r["P"] == 0.0) > 0


@pytest.mark.parametrize("degree,fit_lower", [(2, "explicit"), (3, "explicit"), (3, "augment"), (2, "none")])
@pytest.mark.parametrize("fit_linear,fit_intercept", [(True, True), (False, True), (True, False)])
def test_minibatch_sgd_restatement_with_one_sample_is_the_sequential_step(oracle, degree, fit_lower, fit_linear, fit_intercept):
    """sgd_minibatch_fit (the rule the device's synchronous-minibatch SGD follows) with B = 1 is step()
    (sgd.nim:205-258): same iterates, viol and loss as the line-by-line restatement of SGD.fit, up to the
    rounding of eager vs lazy scaling of the untouched features."""
    n, d, k = 60, 9, 4
    X = make_dense(n, d, 21, density=0.4, positive=False)
    y = np.random.default_rng(1).standard_normal(n)
    csr = CSR.from_dense(X)
    P, w, _ = make_fm_params(d, degree, k, fit_lower, fit_linear, seed=3, scale=0.1)
    kw = dict(eta0=0.05, alpha0=1e-3, alpha=1e-2, beta=2e-2, scheduling="optimal", power=1.0)
    perms = np.array([np.random.default_rng(5 + e).permutation(n) for e in range(3)])
    ref = oracle.sgd_fit(csr, y, P, w, 0.2, degree, "squared", fit_linear, fit_intercept, max_iter=3, perms=perms,
                         it=1, **kw)
    got = oracle.sgd_minibatch_fit(csr, y, P, w, 0.2, degree, "squared", B=1, max_iter=3, perms=perms, it=1,
                                   fit_linear=fit_linear, fit_intercept=fit_intercept, **kw)
    assert got["it"] == ref["it"]
    np.testing.assert_allclose(got["P"], ref["P"], rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(got["w"], ref["w"], rtol=1e-10, atol=1e-14)
    assert abs(got["intercept"] - ref["intercept"]) <= 1e-12
    np.testing.assert_allclose(got["viol"], ref["viol"], rtol=1e-9)
    np.testing.assert_allclose(got["loss"], ref["loss"], rtol=1e-10)


def test_minibatch_sgd_restatement_ffm_with_one_sample(oracle):
    n, d, nf, k = 40, 12, 3, 3
    _, csr, _ = make_field_csr(n, d, nf, 4)
    y = np.random.default_rng(2).standard_normal(n)
    rng = np.random.default_rng(7)
    P, w = rng.standard_normal((nf, d, k)) * 0.1, rng.standard_normal(d) * 0.1
    kw = dict(eta0=0.05, alpha0=1e-3, alpha=1e-2, beta=2e-2, scheduling="optimal", power=1.0)
    ref = oracle.ffm_sgd_fit(csr, y, P, w, 0.1, "squared", True, True, max_iter=2, it=1, **kw)
    got = oracle.sgd_minibatch_fit(csr, y, P, w, 0.1, loss_kind="squared", B=1, max_iter=2, it=1, ffm=True, **kw)
    np.testing.assert_allclose(got["P"], ref["P"], rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(got["w"], ref["w"], rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(got["viol"], ref["viol"], rtol=1e-9)
    np.testing.assert_allclose(got["loss"], ref["loss"], rtol=1e-10)


@pytest.mark.parametrize("degree,fit_lower", [(2, "explicit"), (3, "explicit"), (3, "augment")])
@pytest.mark.parametrize("B", [1, 3, 16])
def test_minibatch_sgd_restatement_vs_naive_dense_definition(oracle, degree, fit_lower, B):
    """oracle.sgd_minibatch_fit (touched / untouched feature sets, CSR) against the naive dense generalisation of
    the reference's SGDSlow (bruteforce.sgd_minibatch_slow_fit: every parameter shrunk at every step); at B = 1
    the latter is sgd_slow_fit itself (tests/optimizer/sgd_slow.nim:38-91)."""
    n, d, k = 37, 6, 3
    X = make_dense(n, d, 33, density=0.5, positive=False)
    y = np.random.default_rng(8).standard_normal(n)
    csr = CSR.from_dense(X)
    P, w, _ = make_fm_params(d, degree, k, fit_lower, True, seed=4, scale=0.1)
    kw = dict(eta0=0.05, alpha0=1e-3, alpha=1e-2, beta=2e-2)
    got = oracle.sgd_minibatch_fit(csr, y, P, w, 0.1, degree, "squared", B=B, max_iter=3, it=1, **kw)
    sP, sw, sb = bf.sgd_minibatch_slow_fit(X, y, P, w, 0.1, degree, True, True, "squared", B, 3, kw["eta0"],
                                           kw["alpha0"], kw["alpha"], kw["beta"])
    np.testing.assert_allclose(got["P"], sP, rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(got["w"], sw, rtol=1e-9, atol=1e-13)
    assert abs(got["intercept"] - sb) <= 1e-12
    if B == 1:
        qP, qw, qb = bf.sgd_slow_fit(X, y, P, w, 0.1, degree, True, True, "squared", 3, kw["eta0"], kw["alpha0"],
                                     kw["alpha"], kw["beta"])
        np.testing.assert_allclose(sP, qP, rtol=1e-10, atol=1e-14)
