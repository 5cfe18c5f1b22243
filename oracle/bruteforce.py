"""Brute-force (subset-enumeration) definitions of the ANOVA / FM / FFM quantities.

TEST INFRASTRUCTURE ONLY -- used to pin oracle/ref_cpu.c the same way the reference pins its own
fast code (differential tests against naive re-implementations; the reference holds no golden
vectors).  Restated from the reference's test helpers (paths relative to /root/reference/tests/):

  anova_slow            kernels_slow.nim:18-27 + comb.nim:1-9
  fm_decision_function  model/fm_slow.nim:42-60
  fm_grad               model/fm_slow.nim:111-134 (comb.nim:12-23 combNotj)
  ffm_decision_function model/ffm_slow.nim:38-56
  ffm_grad              model/ffm_slow.nim:110-127
  cd_slow_fit           optimizer/cd_slow.nim:44-139 + optimizer/fit_linear_slow.nim:4-45
  adagrad_slow_fit      optimizer/adagrad_slow.nim:29-102
  sgd_slow_fit          optimizer/sgd_slow.nim:38-91

Pure Python loops over itertools.combinations: only for the reference's tiny test shapes.
"""
from itertools import combinations
import math

import numpy as np


def anova_slow(x, p, degree, d, m):
    """sum over all index subsets of size `degree` of prod p_j x_j; features j >= d are dummies (x=1)."""
    res = 0.0
    for idx in combinations(range(d + m), degree):
        prod = 1.0
        for j in idx:
            prod *= p[j]
            if j < d:
                prod *= x[j]
        res += prod
    return res


def n_orders(degree, fit_lower):
    if degree == 1:
        return 0
    return degree - 1 if fit_lower == "explicit" else 1


def n_augments(degree, fit_lower, fit_linear):
    if fit_lower != "augment":
        return 0
    return degree - 2 if fit_linear else degree - 1


def fm_decision_function(X, P, w, intercept, degree):
    """X dense [n,d]; P model layout [nOrders,k,d+nAug]."""
    n, d = X.shape
    nO, k, dd = P.shape
    m = dd - d
    out = np.zeros(n)
    for i in range(n):
        r = intercept
        for j in range(d):
            r += w[j] * X[i, j]
        for o in range(nO):
            for s in range(k):
                r += anova_slow(X[i], P[o, s], degree - o, d, m)
        out[i] = r
    return out


def fm_grad(X, i, P, degree, dL, grad):
    """grad[o,s,j] += dL * d(yhat_i)/dP[o,s,j]   (fm_slow.nim:111-134)."""
    n, d = X.shape
    nO, k, dd = P.shape
    for o in range(nO):
        deg = degree - o
        for s in range(k):
            for j in range(dd):
                others = [q for q in range(dd) if q != j]
                tmp = 0.0
                if deg - 1 == 0:
                    tmp = 1.0
                else:
                    for idx in combinations(others, deg - 1):
                        prod = 1.0
                        for j2 in idx:
                            prod *= P[o, s, j2]
                            if j2 < d:
                                prod *= X[i, j2]
                        tmp += prod
                if j < d:
                    tmp *= X[i, j]
                grad[o, s, j] += dL * tmp


def ffm_decision_function(X, fields, P, w, intercept):
    """X dense [n,d]; fields[j] = field of feature j; P [nFields, d, k] (ffm_slow.nim:38-56)."""
    n, d = X.shape
    out = np.zeros(n)
    for i in range(n):
        r = intercept
        for j in range(d):
            r += w[j] * X[i, j]
        for j1 in range(d):
            for j2 in range(j1 + 1, d):
                inter = X[i, j1] * X[i, j2]
                if inter == 0.0:
                    continue
                r += inter * float(np.dot(P[fields[j2], j1], P[fields[j1], j2]))
        out[i] = r
    return out


def ffm_grad(X, fields, i, P, dL, grad):
    n, d = X.shape
    for j1 in range(d):
        for j2 in range(j1 + 1, d):
            inter = X[i, j1] * X[i, j2]
            if inter == 0.0:
                continue
            grad[fields[j2], j1] += dL * P[fields[j1], j2] * inter
            grad[fields[j1], j2] += dL * P[fields[j2], j1] * inter


# ---------------------------------------------------------------------------------------------
# losses (loss.nim), in Python so the slow solvers are independent of the C oracle
# ---------------------------------------------------------------------------------------------
def loss_val(kind, y, p):
    if kind == "squared":
        return 0.5 * (y - p) ** 2
    if kind == "squared_hinge":
        return max(1 - p * y, 0.0) ** 2
    if kind == "logistic":
        z = p * y
        return math.log(1 + math.exp(-z)) if z > 0 else math.log(math.exp(z) + 1) - z
    if kind == "huber":                      # loss.nim:84-87, threshold 1.0
        z = abs(y - p)
        return 0.5 * z ** 2 if z < 1.0 else 1.0 * (z - 0.5)
    raise ValueError(kind)


def dloss_val(kind, y, p):
    if kind == "squared":
        return p - y
    if kind == "squared_hinge":
        z = 1 - p * y
        return -2 * y * z if z > 0 else 0.0
    if kind == "logistic":
        z = p * y
        return -y * math.exp(-z) / (1 + math.exp(-z)) if z > 0 else -y / (math.exp(z) + 1)
    if kind == "huber":                      # loss.nim:90-93 AS WRITTEN: y - p inside the threshold (the negative of
        z = abs(y - p)                       # d loss / d p) and +threshold outside whatever the sign of y - p
        return y - p if z < 1.0 else 1.0
    raise ValueError(kind)


MU = {"squared": 1.0, "squared_hinge": 2.0, "logistic": 0.25}


def cd_slow_fit(X, y, P, w, intercept, degree, fit_linear, fit_intercept, loss, max_iter, alpha0,
                alpha, beta):
    """CDSlow.fit (cd_slow.nim:97-139): naive gradients, full re-prediction after every update."""
    n, d = X.shape
    nO, k, dd = P.shape
    m = dd - d
    P = P.copy()
    w = w.copy()
    alpha0, alpha, beta = alpha0 * n, alpha * n, beta * n
    mu = MU[loss]
    col_norm_sq = (X ** 2).sum(axis=0)
    y_pred = fm_decision_function(X, P, w, intercept, degree)
    for _ in range(max_iter):
        if fit_intercept:  # fit_linear_slow / fit_linear.nim:28-38
            upd = alpha0 * intercept + sum(dloss_val(loss, y[i], y_pred[i]) for i in range(n))
            upd /= mu * n + alpha0
            intercept -= upd
            y_pred = fm_decision_function(X, P, w, intercept, degree)
        if fit_linear:
            for j in range(d):
                upd = alpha * w[j]
                for i in range(n):
                    upd += dloss_val(loss, y[i], y_pred[i]) * X[i, j]
                inv = mu * col_norm_sq[j] + alpha
                if inv < 1e-12:
                    continue
                upd /= inv
                w[j] -= upd
                for i in range(n):
                    y_pred[i] -= upd * X[i, j]
            y_pred = fm_decision_function(X, P, w, intercept, degree)
        for o in range(nO):
            deg = degree - o
            for s in range(k):
                for j in range(dd):
                    dA = np.zeros(n)
                    others = [q for q in range(dd) if q != j]
                    for i in range(n):
                        acc = 0.0
                        if deg - 1 == 0:
                            acc = 1.0
                        else:
                            for idx in combinations(others, deg - 1):
                                prod = 1.0
                                for j2 in idx:
                                    prod *= P[o, s, j2]
                                    if j2 < d:
                                        prod *= X[i, j2]
                                acc += prod
                        if j < d:
                            acc *= X[i, j]
                        dA[i] = acc
                    inv = float((dA ** 2).sum()) * mu + beta
                    g = beta * P[o, s, j]
                    for i in range(n):
                        g += dloss_val(loss, y[i], y_pred[i]) * dA[i]
                    upd = g / inv
                    P[o, s, j] -= upd
                    y_pred = fm_decision_function(X, P, w, intercept, degree)
    return P, w, intercept


def adagrad_slow_fit(X, y, P, w, intercept, degree, fit_linear, fit_intercept, loss, max_iter, eta0,
                     alpha0, alpha, beta, eps):
    """AdaGradSlow.fit (adagrad_slow.nim:29-102), shuffle=false: dense dual-averaging update."""
    n, d = X.shape
    P = P.copy()
    w = w.copy()
    gsP = np.zeros_like(P)
    gnP = np.zeros_like(P) + eps
    gsw = np.zeros(d)
    gnw = np.zeros(d) + eps
    gsb, gnb = 0.0, eps
    it = 1
    for _ in range(max_iter):
        for i in range(n):
            y_pred = fm_decision_function(X[i:i + 1], P, w, intercept, degree)[0]
            dL = dloss_val(loss, y[i], y_pred)
            grad = np.zeros_like(P)
            fm_grad(X, i, P, degree, dL, grad)
            t = float(it)
            if fit_intercept:
                gsb += dL
                gnb += dL ** 2
                intercept = -eta0 * gsb / (math.sqrt(gnb) + eta0 * t * alpha0)
            if fit_linear:
                den = eta0 * t * alpha
                for j in range(d):
                    gsw[j] += dL * X[i, j]
                    gnw[j] += (dL * X[i, j]) ** 2
                    w[j] = -eta0 * gsw[j] / (den + math.sqrt(gnw[j]))
            den = eta0 * t * beta
            gsP += grad
            gnP += grad ** 2
            P = -eta0 * gsP / (den + np.sqrt(gnP))
            it += 1
    return P, w, intercept


def adagrad_minibatch_slow_fit(X, y, P, w, intercept, degree, fit_linear, fit_intercept, loss, B, max_iter, eta0,
                               alpha0, alpha, beta, eps):
    """Naive dense definition of the synchronous-minibatch dual-averaging AdaGrad (the `miniBatchSize > 1` variant
    this repository defines; B = 1 is AdaGradSlow.fit above, adagrad_slow.nim:29-102): before a minibatch EVERY
    parameter is set from the accumulators, theta = -eta0 G / (eta0 t reg + sqrt(N)) with t = samples seen so far
    (none before the very first minibatch, which runs on the given parameters); the B samples are evaluated at that
    snapshot; G += g_i and N += g_i^2 per SAMPLE (adagrad.nim:119-124 squares the per-sample gradient); after the
    last epoch every parameter is set once more (finalize, adagrad.nim:65-84).  Dense X, subset enumeration --
    independent of oracle/ref_cpu.c (lazy refresh of the touched features, CSR) and of the device code."""
    n, d = X.shape
    P = P.copy()
    w = w.copy()
    gsP, gnP = np.zeros_like(P), np.zeros_like(P) + eps
    gsw, gnw = np.zeros(d), np.zeros(d) + eps
    gsb, gnb = 0.0, eps
    seen = 0

    def refresh():
        nonlocal P, w, intercept
        t = float(seen)
        P = -eta0 * gsP / (eta0 * t * beta + np.sqrt(gnP))
        if fit_linear:
            w = -eta0 * gsw / (eta0 * t * alpha + np.sqrt(gnw))
        if fit_intercept:
            intercept = -eta0 * gsb / (math.sqrt(gnb) + eta0 * t * alpha0)

    for _ in range(max_iter):
        for q0 in range(0, n, B):
            rows = range(q0, min(n, q0 + B))
            if seen > 0:
                refresh()
            y_pred = fm_decision_function(X[q0:q0 + len(rows)], P, w, intercept, degree)
            for t_, i in enumerate(rows):
                dL = dloss_val(loss, y[i], y_pred[t_])
                g = np.zeros_like(P)
                fm_grad(X, i, P, degree, dL, g)
                gsP += g
                gnP += g ** 2
                if fit_linear:
                    gsw += dL * X[i]
                    gnw += (dL * X[i]) ** 2
                if fit_intercept:
                    gsb += dL
                    gnb += dL ** 2
            seen += len(rows)
    refresh()
    return P, w, intercept


def sgd_slow_fit(X, y, P, w, intercept, degree, fit_linear, fit_intercept, loss, max_iter, eta0,
                 alpha0, alpha, beta, power=1.0):
    """SGDSlow.fit (sgd_slow.nim:38-91), shuffle=false, scheduling=optimal: dense updates with the
    L2 shrink applied to every parameter at every step."""
    n, d = X.shape
    P = P.copy()
    w = w.copy()
    it = 1

    def eta(reg):
        return eta0 / (1.0 + eta0 * reg * it) ** power

    for _ in range(max_iter):
        for i in range(n):
            y_pred = fm_decision_function(X[i:i + 1], P, w, intercept, degree)[0]
            dL = dloss_val(loss, y[i], y_pred)
            grad = np.zeros_like(P)
            fm_grad(X, i, P, degree, dL, grad)
            P -= eta(beta) * (grad + beta * P)
            if fit_linear:
                w -= eta(alpha) * (dL * X[i] + alpha * w)
            if fit_intercept:
                intercept -= eta(alpha0) * (dL + alpha0 * intercept)
            it += 1
    return P, w, intercept


def sgd_minibatch_slow_fit(X, y, P, w, intercept, degree, fit_linear, fit_intercept, loss, B, max_iter, eta0,
                           alpha0, alpha, beta, power=1.0):
    """The synchronous-minibatch generalisation of SGDSlow.fit (sgd_slow.nim:38-91: dense updates, the L2 shrink
    applied to every parameter at every step): the B samples of a minibatch are evaluated at the same
    parameters, then their B dense updates are applied at once with the step sizes of the minibatch's first
    iteration -- theta <- (1 - eta reg)^B theta - eta sum_i grad_i.  B = 1 is sgd_slow_fit.  Independent of
    oracle.sgd_minibatch_fit (which works on touched / untouched feature sets) and of the device code."""
    n, d = X.shape
    P = P.copy()
    w = w.copy()
    it = 1

    def eta(reg):
        return eta0 / (1.0 + eta0 * reg * it) ** power

    def spow(base, m):
        s = 1.0
        for _ in range(m):
            s *= base
        return s

    for _ in range(max_iter):
        for q0 in range(0, n, B):
            rows = range(q0, min(n, q0 + B))
            m = len(rows)
            y_pred = fm_decision_function(X[q0:q0 + m], P, w, intercept, degree)
            grad = np.zeros_like(P)
            gw, gb = np.zeros(d), 0.0
            for t, i in enumerate(rows):
                dL = dloss_val(loss, y[i], y_pred[t])
                fm_grad(X, i, P, degree, dL, grad)
                gw += dL * X[i]
                gb += dL
            eP, eW, eB = eta(beta), eta(alpha), eta(alpha0)
            P = spow(1.0 - eP * beta, m) * P - eP * grad
            if fit_linear:
                w = spow(1.0 - eW * alpha, m) * w - eW * gw
            if fit_intercept:
                intercept = spow(1.0 - eB * alpha0, m) * intercept - eB * gb
            it += m
    return P, w, intercept


def ffm_sgd_slow_fit(X, fields, y, P, w, intercept, fit_linear, fit_intercept, loss, max_iter, eta0, alpha0, alpha,
                     beta, power=1.0):
    """SGDSlow.fit for FFMSlow (tests/optimizer/sgd_ffm_slow.nim:8-58), shuffle = false, scheduling = optimal: dense
    updates, the L2 shrink applied to every parameter at every step.  P is [nFields, d, k]."""
    n, d = X.shape
    P = P.copy()
    w = w.copy()
    it = 1

    def eta(reg):
        return eta0 / (1.0 + eta0 * reg * it) ** power

    for _ in range(max_iter):
        for i in range(n):
            y_pred = ffm_decision_function(X[i:i + 1], fields, P, w, intercept)[0]
            dL = dloss_val(loss, y[i], y_pred)
            grad = np.zeros_like(P)
            ffm_grad(X, fields, i, P, dL, grad)
            if fit_intercept:
                intercept -= eta(alpha0) * (dL + alpha0 * intercept)
            if fit_linear:
                w -= eta(alpha) * (dL * X[i] + alpha * w)
            P -= eta(beta) * (grad + beta * P)
            it += 1
    return P, w, intercept


def ffm_adagrad_slow_fit(X, fields, y, P, w, intercept, fit_linear, fit_intercept, loss, max_iter, eta0, alpha0,
                         alpha, beta, eps):
    """AdaGradSlow.fit for FFMSlow (tests/optimizer/adagrad_ffm_slow.nim:8-40 + adagrad_slow.nim:29-69), shuffle =
    false: dense dual-averaging update of every parameter after every sample."""
    n, d = X.shape
    P = P.copy()
    w = w.copy()
    gsP, gnP = np.zeros_like(P), np.zeros_like(P) + eps
    gsw, gnw = np.zeros(d), np.zeros(d) + eps
    gsb, gnb = 0.0, eps
    it = 1
    for _ in range(max_iter):
        for i in range(n):
            y_pred = ffm_decision_function(X[i:i + 1], fields, P, w, intercept)[0]
            dL = dloss_val(loss, y[i], y_pred)
            grad = np.zeros_like(P)
            ffm_grad(X, fields, i, P, dL, grad)
            t = float(it)
            if fit_intercept:
                gsb += dL
                gnb += dL ** 2
                intercept = -eta0 * gsb / (math.sqrt(gnb) + eta0 * t * alpha0)
            if fit_linear:
                gsw += dL * X[i]
                gnw += (dL * X[i]) ** 2
                w = -eta0 * gsw / (eta0 * t * alpha + np.sqrt(gnw))
            gsP += grad
            gnP += grad ** 2
            P = -eta0 * gsP / (eta0 * t * beta + np.sqrt(gnP))
            it += 1
    return P, w, intercept


# ---------------------------------------------------------------- proximal operators (definitions)
def prox_squaredl12_sorted(p, lam):
    """argmin_q 0.5*||q - p||^2 + lam * ||q||_1^2 in closed form: with a = sort(|p|, descending) and
    prefix sums c_m, theta = the largest m with a_m > 2*lam*c_m / (1 + 2*lam*m), S = c_theta / (1 +
    2*lam*theta), q = softthreshold(p, 2*lam*S) -- what proxSquaredL12 (regularizer/squaredl12.nim:16-64)
    finds by randomised selection."""
    p = np.asarray(p, dtype=np.float64)
    a = np.sort(np.abs(p))[::-1]
    c = np.cumsum(a)
    m = np.arange(1, len(a) + 1)
    ok = a > 2 * lam * c / (1.0 + 2.0 * lam * m)
    theta = int(np.max(m[ok])) if np.any(ok) else 0
    S = (c[theta - 1] if theta else 0.0) / (1.0 + 2.0 * lam * theta)
    return np.sign(p) * np.maximum(np.abs(p) - 2 * lam * S, 0.0)


def prox_l21_row(p, lam):
    """regularizer/l21.nim:25-29"""
    p = np.asarray(p, dtype=np.float64)
    nrm = np.sqrt(np.sum(p * p))
    return p * (1.0 - lam / nrm) if nrm > lam else np.zeros_like(p)


def prox_squaredl12_slow(p, lam):
    """proxSquaredL12Slow, tests/regularizer/squaredl12_slow.nim:10-25 (the reference's own test definition)"""
    p = np.asarray(p, dtype=np.float64).copy()
    n = len(p)
    absp = np.sort(np.abs(p))[::-1]
    S = 2.0 * lam * np.cumsum(absp)
    for i in range(n):
        S[i] /= 1.0 + 2.0 * lam * (i + 1.0)
    theta = 0
    for i in range(n):
        if absp[i] - S[i] < 0:
            break
        theta += 1
    for i in range(n):
        if abs(p[i]) < absp[theta - 1]:
            p[i] = 0.0
        else:
            p[i] = np.sign(p[i]) * max(abs(p[i]) - S[theta - 1], 0.0)
    return p


def mbpsgd_slow_fit(X, y, P, w, intercept, degree, fit_linear, fit_intercept, loss, max_iter, eta0, alpha0, alpha,
                    beta, gamma=0.0, reg="identity", mini_batch_size=-1, max_iter_inner=-1, power=1.0):
    """Naive dense definition of minibatch proximal SGD, the way the reference's *_slow.nim helpers define their
    solvers (the reference has no MBPSGDSlow; minibatch_psgd.nim has no test): dense X, yhat and its gradient by
    subset enumeration (fm_decision_function / fm_grad above), a minibatch = the next `mb` rows of the cyclic order
    0, 1, ..., n-1 (shuffle = false),
        g = (1/mb) sum_i dloss_i * grad yhat_i,   theta <- (theta - eta g) / (1 + eta * l2)      (params.nim:90-98)
    with eta = eta0 / (1 + eta0 * l2 * it)^power per parameter group (sgd.nim:60-65, 'optimal'), then the
    regulariser's proximal operator with lam = gamma * eta_P / (1 + eta_P * beta) on every order
    (minibatch_psgd.nim:96-122), it += 1 per minibatch.  The intercept takes its gradient step only when fitLinear
    holds as well (params.nim:47: `self.fitIntercept and grad.fitLinear`) -- the reference's quirk, part of the
    definition here.  Default sizes as minibatch_psgd.nim:157-164.  Returns (P, w, intercept, epoch losses).
    Independent of oracle/ref_cpu.c (CSR, per-sample scatter, the solver layout) and of the device code."""
    n, d = X.shape
    P = P.copy()
    w = w.copy()
    nnz = int(np.count_nonzero(X))
    mb = mini_batch_size if mini_batch_size > 0 else max((d * n) // nnz, 1)
    inner = max_iter_inner if max_iter_inner > 0 else max((n - 1) // mb + 1, 1)
    it, ii, losses = 1, 0, []

    def eta(l2):
        return eta0 / (1.0 + eta0 * l2 * it) ** power

    def soft(v, lam):
        return np.sign(v) * np.maximum(np.abs(v) - lam, 0.0)

    for _ in range(max_iter):
        run = 0.0
        for _ in range(inner):
            gP, gw, gb = np.zeros_like(P), np.zeros(d), 0.0
            for _ in range(mb):
                i = ii
                y_pred = fm_decision_function(X[i:i + 1], P, w, intercept, degree)[0]
                run += loss_val(loss, y[i], y_pred)
                dL = dloss_val(loss, y[i], y_pred) / mb
                fm_grad(X, i, P, degree, dL, gP)
                gw += dL * X[i]
                gb += dL
                ii = (ii + 1) % n
            eP, eW, eB = eta(beta), eta(alpha), eta(alpha0)
            P = (P - eP * gP) / (1.0 + eP * beta)
            if fit_linear:
                w = (w - eW * gw) / (1.0 + eW * alpha)
            if fit_intercept:
                if fit_linear:
                    intercept -= eB * gb
                intercept /= 1.0 + eB * alpha0
            lam = gamma * eP / (1.0 + eP * beta)
            for o in range(P.shape[0]):          # P[o] is [k, d + nAug] here; the reference's prox sees its transpose
                if reg == "l1":
                    P[o] = soft(P[o], lam)
                elif reg == "l21":               # one vector per feature (l21.nim:25-35)
                    for j in range(P.shape[2]):
                        P[o, :, j] = prox_l21_row(P[o, :, j], lam)
                elif reg == "squaredl12":        # transpose = true: one vector per component over all features
                    for s_ in range(P.shape[1]):
                        P[o, s_, :] = prox_squaredl12_sorted(P[o, s_, :], lam)
                elif reg == "squaredl12_rows":   # transpose = false: one vector per feature
                    for j in range(P.shape[2]):
                        P[o, :, j] = prox_squaredl12_sorted(P[o, :, j], lam)
                elif reg != "identity":
                    raise ValueError(reg)
            it += 1
        losses.append(run / (mb * inner))
    return P, w, intercept, losses


def pcd_slow_fit(X, y, P, w, intercept, degree, fit_linear, fit_intercept, loss, max_iter, alpha0,
                 alpha, beta, gamma, reg):
    """PCDSlow.fit (tests/optimizer/pcd_slow.nim:27-113): the naive dense solver the reference tests
    pcd.nim against -- gradient step on one coordinate, then the regulariser's coordinate prox
    (tests/regularizer/l1_slow.nim:28-31, squaredl12_slow.nim:43-56), full re-prediction after every
    update.  reg in l1 / squaredl12 (transpose=true) / squaredl12_rows."""
    n, d = X.shape
    nO, k, dd = P.shape
    P = P.copy()
    w = w.copy()
    alpha0, alpha, beta, gamma = alpha0 * n, alpha * n, beta * n, gamma * n
    mu = MU[loss]
    col_norm_sq = (X ** 2).sum(axis=0)
    y_pred = fm_decision_function(X, P, w, intercept, degree)
    for _ in range(max_iter):
        if fit_intercept:
            upd = alpha0 * intercept + sum(dloss_val(loss, y[i], y_pred[i]) for i in range(n))
            upd /= mu * n + alpha0
            intercept -= upd
            y_pred = fm_decision_function(X, P, w, intercept, degree)
        if fit_linear:
            for j in range(d):
                upd = alpha * w[j]
                for i in range(n):
                    upd += dloss_val(loss, y[i], y_pred[i]) * X[i, j]
                inv = mu * col_norm_sq[j] + alpha
                if inv < 1e-12:
                    continue
                upd /= inv
                w[j] -= upd
                for i in range(n):
                    y_pred[i] -= upd * X[i, j]
            y_pred = fm_decision_function(X, P, w, intercept, degree)
        for o in range(nO):
            deg = degree - o
            for s in range(k):
                for j in range(dd):
                    dA = np.zeros(n)
                    others = [q for q in range(dd) if q != j]
                    for i in range(n):
                        acc = 1.0 if deg - 1 == 0 else 0.0
                        if deg - 1 > 0:
                            for idx in combinations(others, deg - 1):
                                prod = 1.0
                                for j2 in idx:
                                    prod *= P[o, s, j2]
                                    if j2 < d:
                                        prod *= X[i, j2]
                                acc += prod
                        if j < d:
                            acc *= X[i, j]
                        dA[i] = acc
                    inv = float((dA ** 2).sum()) * mu + beta
                    if inv < 1e-12:
                        continue
                    g = beta * P[o, s, j]
                    for i in range(n):
                        g += dloss_val(loss, y[i], y_pred[i]) * dA[i]
                    P[o, s, j] -= g / inv
                    lam = gamma / inv
                    psj = P[o, s, j]
                    if reg == "l1":
                        P[o, s, j] = np.sign(psj) * max(abs(psj) - lam, 0.0)
                    else:
                        strength = (np.abs(P[o, s, :]).sum() if reg == "squaredl12" else np.abs(P[o, :, j]).sum()) - abs(psj)
                        P[o, s, j] = np.sign(psj / (1 + 2 * lam)) * max(abs(psj / (1 + 2 * lam)) - 2 * lam * strength / (1 + 2 * lam), 0.0)
                    y_pred = fm_decision_function(X, P, w, intercept, degree)
    return P, w, intercept


def psgd_slow_fit(X, y, P, w, intercept, degree, fit_linear, fit_intercept, loss, max_iter, eta0, alpha0,
                  alpha, beta, gamma, reg, it=1):
    """PSGDSlow.fit (tests/optimizer/psgd_slow.nim:42-99), shuffle=false, scheduling=optimal, power=1:
    dense gradient step on every parameter, then the regulariser's full prox, every sample.
    P model layout [nOrders, k, d+nAug]; reg in l1 / l21 / squaredl12 / squaredl12_rows."""
    n, d = X.shape
    nO, k, dd = P.shape
    P = P.copy()
    w = w.copy()

    def eta(r):
        return eta0 / (1.0 + eta0 * r * it)

    for _ in range(max_iter):
        for i in range(n):
            y_pred = fm_decision_function(X[i:i + 1], P, w, intercept, degree)[0]
            dL = dloss_val(loss, y[i], y_pred)
            grad = np.zeros_like(P)
            fm_grad(X, i, P, degree, dL, grad)
            eta_w, eta_P = eta(alpha), eta(beta)
            eta_s = eta_P / (1.0 + eta_P * beta)
            if fit_intercept:
                upd = eta(alpha0) * (dL + alpha0 * intercept)
                intercept -= upd / (1.0 + eta(alpha0) * alpha0)
            if fit_linear:
                for j in range(d):
                    upd = eta_w * (dL * X[i, j] + alpha * w[j])
                    w[j] -= upd / (1.0 + eta_w * alpha)
            for o in range(nO):
                P[o] -= eta_s * (grad[o] + beta * P[o])
                lam = gamma * eta_s
                if reg == "l1":
                    P[o] = np.sign(P[o]) * np.maximum(np.abs(P[o]) - lam, 0.0)
                elif reg == "l21":
                    for j in range(dd):
                        P[o][:, j] = prox_l21_row(P[o][:, j], lam)
                elif reg == "squaredl12":
                    for s in range(k):
                        P[o][s] = prox_squaredl12_slow(P[o][s], lam)
                else:
                    for j in range(dd):
                        P[o][:, j] = prox_squaredl12_slow(P[o][:, j], lam)
            it += 1
    return P, w, intercept
