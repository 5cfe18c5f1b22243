"""Shared synthetic-data helpers for the test-suite (NumPy RNG; never Nim's RNG)."""
import numpy as np

from oracle import bruteforce as bf
from oracle.oracle import CSR


def make_dense(n, d, seed, density=1.0, positive=True):
    rng = np.random.default_rng(seed)
    X = rng.random((n, d)) if positive else rng.standard_normal((n, d))
    if density < 1.0:
        X = X * (rng.random((n, d)) < density)
    return X


def make_fm_params(d, degree, k, fit_lower, fit_linear, seed, scale=0.1):
    rng = np.random.default_rng(seed + 1000)
    nO = bf.n_orders(degree, fit_lower)
    nA = bf.n_augments(degree, fit_lower, fit_linear)
    P = rng.standard_normal((nO, k, d + nA)) * scale
    w = rng.standard_normal(d) * scale if fit_linear else np.zeros(d)
    return P, w, nA


def make_field_csr(n, d, n_fields, seed, density=0.6):
    """Dense-ish field data like tests/test_sgd_ffm.nim (n=80, d=20, 5 fields): feature j belongs
    to field j % n_fields; indices sorted within a row."""
    rng = np.random.default_rng(seed)
    X = rng.random((n, d)) * (rng.random((n, d)) < density)
    csr = CSR.from_dense(X)
    field_of = np.arange(d) % n_fields
    csr.fields = field_of[csr.indices].astype(np.int64)
    csr.n_fields = n_fields
    return X, csr, field_of


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = np.maximum(np.abs(b), 1e-300)
    return float(np.max(np.abs(a - b) / den)) if a.size else 0.0


def max_rel(a, b, floor=1.0):
    """max |a-b| / max(|b|, floor*max|b|*1e-6): relative to the element, guarded near zero."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    scale = np.maximum(np.abs(b), np.max(np.abs(b)) * 1e-6 + 1e-300)
    return float(np.max(np.abs(a - b) / scale))


def sums_agree(a, b, norm_tol=1e-10, elem_tol=1e-7):
    """Two evaluations of the SAME sum of millions of signed FP64 terms in different orders (RED arrival order,
    halves added on the host).  The rounding error of such a sum is bounded by eps * sum|terms|, not by
    eps * |sum|: an element whose terms nearly cancel can sit 1e-6 below the largest one and still carry the
    absolute error of the large partial sums, so the element-wise ratio of `max_rel` is only meaningful at a
    looser bar (1e-7; 1.8e-9 was observed on the 10 M-row FFM gradient) while the error relative to the largest
    element is held tight (1e-10; a random walk of 1e7 roundings is ~4e-13)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return True
    top = float(np.max(np.abs(b))) + 1e-300
    return float(np.max(np.abs(a - b))) <= norm_tol * top and max_rel(a, b) <= elem_tol
