"""ctypes/numpy front-end of the CPU oracle (oracle/ref_cpu.c, oracle/ref_hogwild.c).

TEST INFRASTRUCTURE ONLY.  May be imported from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- never from nimfm_b200/.  See ref_cpu.c for the parity
status ("pinned against brute-force definitions + committed fixtures; RNG paths unpinned").
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libnimfm_oracle.so")

LOSS = {"squared": 0, "squared_hinge": 1, "logistic": 2, "huber": 3}
SCHED = {"constant": 0, "optimal": 1, "invscaling": 2, "pegasos": 3}

c_i64 = C.c_int64
c_dbl = C.c_double
c_int = C.c_int
PD = C.POINTER(C.c_double)
PI = C.POINTER(C.c_int64)


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("ref_cpu.c", "ref_hogwild.c", "ref_pcd.c", "ref_psgd.c")]
    if (not force and os.path.exists(_SO)
            and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in srcs)):
        return _SO
    subprocess.check_call(["make", "-C", _HERE, "-s", "clean", "all"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        for name in ("ref_loss", "ref_dloss", "ref_regularization", "ref_predict_with_grad",
                     "ref_fm_loss_grad", "ref_ffm_predict_with_grad", "ref_ffm_loss_grad",
                     "ref_hogwild_adagrad_epoch", "ref_reg_eval"):
            getattr(_lib, name).restype = c_dbl
        _lib.ref_mu.restype = c_dbl
    return _lib


def _d(a):
    return None if a is None else a.ctypes.data_as(PD)


def _i(a):
    return None if a is None else a.ctypes.data_as(PI)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


class CSR:
    """Plain host CSR/CSC triple with the reference's dtypes (f64 data, i64 indices/indptr)."""

    def __init__(self, data, indices, indptr, n, d, fields=None, n_fields=0):
        self.data, self.indices, self.indptr = f64(data), i64(indices), i64(indptr)
        self.n, self.d = int(n), int(d)
        self.fields = None if fields is None else i64(fields)
        self.n_fields = int(n_fields)

    @staticmethod
    def from_dense(X):
        X = np.asarray(X, dtype=np.float64)
        n, d = X.shape
        data, indices, indptr = [], [], [0]
        for i in range(n):
            nz = np.nonzero(X[i])[0]
            indices.extend(nz.tolist())
            data.extend(X[i, nz].tolist())
            indptr.append(len(indices))
        return CSR(data, indices, indptr, n, d)

    def to_dense(self):
        X = np.zeros((self.n, self.d))
        for i in range(self.n):
            for jj in range(self.indptr[i], self.indptr[i + 1]):
                X[i, self.indices[jj]] = self.data[jj]
        return X


def loss(kind, y, p, thr=1.0):
    return lib().ref_loss(LOSS[kind], c_dbl(thr), c_dbl(y), c_dbl(p))


def dloss(kind, y, p, thr=1.0):
    return lib().ref_dloss(LOSS[kind], c_dbl(thr), c_dbl(y), c_dbl(p))


def loss_vec(kind, y, p, thr=1.0):
    return np.array([loss(kind, a, b, thr) for a, b in zip(y, p)])


def csr_to_csc(X):
    nnz = len(X.data)
    od, oi, op = np.zeros(nnz), np.zeros(nnz, np.int64), np.zeros(X.d + 1, np.int64)
    lib().ref_csr_to_csc(c_i64(X.n), c_i64(X.d), _d(X.data), _i(X.indices), _i(X.indptr), _d(od),
                         _i(oi), _i(op))
    return CSR(od, oi, op, X.n, X.d)


def csc_to_csr(Xc):
    nnz = len(Xc.data)
    od, oi, op = np.zeros(nnz), np.zeros(nnz, np.int64), np.zeros(Xc.n + 1, np.int64)
    lib().ref_csc_to_csr(c_i64(Xc.n), c_i64(Xc.d), _d(Xc.data), _i(Xc.indices), _i(Xc.indptr),
                         _d(od), _i(oi), _i(op))
    return CSR(od, oi, op, Xc.n, Xc.d)


def csr_take_rows(X, rows):
    rows = i64(rows)
    op = np.zeros(len(rows) + 1, np.int64)
    lib().ref_csr_take_rows(_d(X.data), _i(X.indices), _i(X.indptr), _i(rows), c_i64(len(rows)),
                            None, None, _i(op))
    nnz = int(op[-1])
    od, oi = np.zeros(nnz), np.zeros(nnz, np.int64)
    lib().ref_csr_take_rows(_d(X.data), _i(X.indices), _i(X.indptr), _i(rows), c_i64(len(rows)),
                            _d(od), _i(oi), _i(op))
    return CSR(od, oi, op, len(rows), X.d)


def csc_slice_rows(Xc, a, b):
    op = np.zeros(Xc.d + 1, np.int64)
    lib().ref_csc_slice_rows(c_i64(Xc.d), _d(Xc.data), _i(Xc.indices), _i(Xc.indptr), c_i64(a),
                             c_i64(b), None, None, _i(op))
    nnz = int(op[-1])
    od, oi = np.zeros(nnz), np.zeros(nnz, np.int64)
    lib().ref_csc_slice_rows(c_i64(Xc.d), _d(Xc.data), _i(Xc.indices), _i(Xc.indptr), c_i64(a),
                             c_i64(b), _d(od), _i(oi), _i(op))
    return CSR(od, oi, op, b - a + 1, Xc.d)


def csr_vstack(parts):
    """vstack of CSRMatrix / CSRFieldMatrix (tensor/sparse.nim:564-583, 612-640): data / indices
    [/ fields] concatenated, indptr[1..] of each later part shifted by the nnz so far."""
    data, indices, indptr = parts[0].data.copy(), parts[0].indices.copy(), parts[0].indptr.copy()
    fields = None if parts[0].fields is None else parts[0].fields.copy()
    n = parts[0].n
    for X in parts[1:]:
        if X.d != parts[0].d:
            raise ValueError("All matrics must have the same shape[1].")
        nnz = len(data)
        data = np.concatenate([data, X.data])
        indices = np.concatenate([indices, X.indices])
        if fields is not None:
            fields = np.concatenate([fields, X.fields])
        indptr = np.concatenate([indptr, X.indptr[1:] + nnz])
        n += X.n
    return CSR(data, indices, indptr, n, parts[0].d, fields, parts[0].n_fields)


def csc_vstack(parts):
    """vstack of CSCMatrix (tensor/sparse.nim:586-609): per column j, the entries of every part in
    turn with row ids offset by the rows of the parts before it."""
    d = parts[0].d
    for X in parts[1:]:
        if X.d != d:
            raise ValueError("All matrics must have the same shape[1].")
    data, indices, indptr = [], [], np.zeros(d + 1, np.int64)
    for j in range(d):
        offset = 0
        for X in parts:
            for q in range(X.indptr[j], X.indptr[j + 1]):
                data.append(X.data[q])
                indices.append(X.indices[q] + offset)
                indptr[j + 1] += 1
            offset += X.n
    return CSR(data, indices, np.cumsum(indptr), sum(X.n for X in parts), d)


def anova(X, Ps, degree, n_aug=0, is_csc=False):
    """kernels.anova for one component; returns A [n, degree+1]."""
    A = np.zeros((X.n, degree + 1))
    fn = lib().ref_anova_csc if is_csc else lib().ref_anova_csr
    fn(c_i64(X.n), c_i64(X.d), c_int(n_aug), _d(X.data), _i(X.indices), _i(X.indptr), _d(f64(Ps)),
       _d(A), c_int(degree + 1), c_int(degree))
    return A


def fm_decision_function(X, P, w, intercept, degree, is_csc=False, lams=None):
    """P: model layout [nOrders, k, d+nAug]."""
    P = f64(P)
    nO, k, dd = P.shape
    out = np.zeros(X.n)
    lib().ref_fm_decision_function(c_int(int(is_csc)), c_i64(X.n), c_i64(X.d), _d(X.data),
                                   _i(X.indices), _i(X.indptr), c_int(degree), c_int(k), c_int(nO),
                                   c_int(dd - X.d), _d(P), _d(f64(w)), c_dbl(intercept),
                                   _d(None if lams is None else f64(lams)), _d(out))
    return out


def to_feature_major(P):
    return np.ascontiguousarray(np.transpose(f64(P), (0, 2, 1)))


def to_component_major(Pf):
    return np.ascontiguousarray(np.transpose(f64(Pf), (0, 2, 1)))


def fm_loss_grad(X, y, P, w, intercept, degree, loss_kind="squared", row_begin=0, row_end=None,
                 fit_linear=True, fit_intercept=True, mini_batch_size=None, thr=1.0):
    """updateGradient over rows [row_begin,row_end). P: model layout. Returns dict with loss sum,
    yPred, grads in the MODEL layout (gP [nOrders,k,d+aug]), gw, gb."""
    P = f64(P)
    nO, k, dd = P.shape
    Pf = to_feature_major(P)
    if row_end is None:
        row_end = X.n
    mb = mini_batch_size if mini_batch_size is not None else (row_end - row_begin)
    gP, gw, gb = np.zeros_like(Pf), np.zeros(X.d), C.c_double(0.0)
    ypred = np.zeros(row_end - row_begin)
    dA = np.zeros_like(Pf)
    ls = lib().ref_fm_loss_grad(c_i64(X.d), _d(X.data), _i(X.indices), _i(X.indptr), _d(f64(y)),
                                c_i64(row_begin), c_i64(row_end), c_int(degree), c_int(k), c_int(nO),
                                c_int(dd - X.d), c_int(int(fit_linear)), c_int(int(fit_intercept)),
                                _d(Pf), _d(f64(w)), c_dbl(intercept), c_int(LOSS[loss_kind]),
                                c_dbl(thr), c_i64(mb), _d(gP), _d(gw), C.byref(gb), _d(ypred), _d(dA))
    return dict(loss=ls, y_pred=ypred, gP=to_component_major(gP), gw=gw, gb=gb.value)


REG = {"identity": 0, "l1": 1, "squaredl12": 2, "squaredl12_rows": 3, "l21": 4}


def prox_matrix(Pjs, lam, reg):
    """reg.prox on one order's solver-layout matrix P[j][s] (returns a new array)."""
    out = f64(Pjs).copy()
    lib().ref_prox_matrix(_d(out), c_i64(out.shape[0]), c_int(out.shape[1]), c_dbl(lam), c_int(REG[reg]))
    return out


def reg_eval(Pjs, reg):
    Pjs = f64(Pjs)
    return lib().ref_reg_eval(_d(Pjs), c_i64(Pjs.shape[0]), c_int(Pjs.shape[1]), c_int(REG[reg]))


def mbpsgd_fit(X, y, P, w, intercept, degree, loss_kind="squared", fit_linear=True,
               fit_intercept=True, max_iter=10, eta0=0.1, alpha0=1e-6, alpha=1e-3, beta=1e-4,
               gamma=0.0, reg="identity", mini_batch_size=-1, max_iter_inner=-1,
               scheduling="optimal", power=1.0, tol=0.0, perms=None, it=1, thr=1.0):
    P = f64(P).copy()
    w = f64(w).copy()
    nO, k, dd = P.shape
    b = C.c_double(intercept)
    itc = C.c_int64(it)
    el = np.zeros(max_iter)
    if perms is not None:
        perms = i64(perms).reshape(-1, X.n)
    ne = lib().ref_mbpsgd_fit(
        c_i64(X.n), c_i64(X.d), _d(X.data), _i(X.indices), _i(X.indptr), _d(f64(y)), c_int(degree),
        c_int(k), c_int(nO), c_int(dd - X.d), c_int(int(fit_linear)), c_int(int(fit_intercept)),
        _d(P), _d(w), C.byref(b), c_int(LOSS[loss_kind]), c_dbl(thr), c_int(max_iter), c_dbl(eta0),
        c_dbl(alpha0), c_dbl(alpha), c_dbl(beta), c_dbl(gamma), c_int(REG[reg]),
        c_i64(mini_batch_size), c_i64(max_iter_inner), c_int(SCHED[scheduling]), c_dbl(power),
        c_dbl(tol), _i(perms), c_i64(0 if perms is None else perms.shape[0]), C.byref(itc), _d(el))
    return dict(P=P, w=w, intercept=b.value, it=itc.value, epoch_loss=el[:ne], epochs=ne)


def _adagrad_state(nP_shape, d, state):
    if state is None:
        return (np.zeros(nP_shape), np.zeros(nP_shape), np.zeros(d), np.zeros(d), C.c_double(0.0),
                C.c_double(0.0))
    return (f64(state["gsP"]).copy(), f64(state["gnP"]).copy(), f64(state["gsw"]).copy(),
            f64(state["gnw"]).copy(), C.c_double(state["gsb"]), C.c_double(state["gnb"]))


def adagrad_fit(X, y, P, w, intercept, degree, loss_kind="squared", fit_linear=True,
                fit_intercept=True, max_iter=10, eta0=0.1, alpha0=1e-6, alpha=1e-3, beta=1e-3,
                eps=1e-10, tol=0.0, mini_batch_size=1, perms=None, it=1, state=None, thr=1.0):
    """state (g_sum/g_norm) uses the SOLVER layout [nOrders, d+aug, k]."""
    P = f64(P).copy()
    w = f64(w).copy()
    nO, k, dd = P.shape
    b, itc = C.c_double(intercept), C.c_int64(it)
    gsP, gnP, gsw, gnw, gsb, gnb = _adagrad_state((nO, dd, k), X.d, state)
    viol, ls = np.zeros(max_iter), np.zeros(max_iter)
    if perms is not None:
        perms = i64(perms).reshape(-1, X.n)
        assert perms.shape[0] >= max_iter
    ne = lib().ref_adagrad_fit(
        c_i64(X.n), c_i64(X.d), _d(X.data), _i(X.indices), _i(X.indptr), _d(f64(y)), c_int(degree),
        c_int(k), c_int(nO), c_int(dd - X.d), c_int(int(fit_linear)), c_int(int(fit_intercept)),
        _d(P), _d(w), C.byref(b), c_int(LOSS[loss_kind]), c_dbl(thr), c_int(max_iter), c_dbl(eta0),
        c_dbl(alpha0), c_dbl(alpha), c_dbl(beta), c_dbl(eps), c_dbl(tol), c_i64(mini_batch_size),
        _i(perms), C.byref(itc), _d(gsP), _d(gnP), _d(gsw), _d(gnw), C.byref(gsb), C.byref(gnb),
        _d(viol), _d(ls))
    st = dict(gsP=gsP, gnP=gnP, gsw=gsw, gnw=gnw, gsb=gsb.value, gnb=gnb.value)
    return dict(P=P, w=w, intercept=b.value, it=itc.value, viol=viol[:ne], loss=ls[:ne], epochs=ne,
                state=st)


def sgd_fit(X, y, P, w, intercept, degree, loss_kind="squared", fit_linear=True, fit_intercept=True,
            max_iter=10, eta0=0.01, alpha0=1e-6, alpha=1e-3, beta=1e-3, scheduling="optimal",
            power=1.0, tol=0.0, perms=None, it=1, thr=1.0):
    P = f64(P).copy()
    w = f64(w).copy()
    nO, k, dd = P.shape
    b, itc = C.c_double(intercept), C.c_int64(it)
    viol, ls = np.zeros(max_iter), np.zeros(max_iter)
    if perms is not None:
        perms = i64(perms).reshape(-1, X.n)
    ne = lib().ref_sgd_fit(
        c_i64(X.n), c_i64(X.d), _d(X.data), _i(X.indices), _i(X.indptr), _d(f64(y)), c_int(degree),
        c_int(k), c_int(nO), c_int(dd - X.d), c_int(int(fit_linear)), c_int(int(fit_intercept)),
        _d(P), _d(w), C.byref(b), c_int(LOSS[loss_kind]), c_dbl(thr), c_int(max_iter), c_dbl(eta0),
        c_dbl(alpha0), c_dbl(alpha), c_dbl(beta), c_int(SCHED[scheduling]), c_dbl(power), c_dbl(tol),
        _i(perms), C.byref(itc), _d(viol), _d(ls))
    return dict(P=P, w=w, intercept=b.value, it=itc.value, viol=viol[:ne], loss=ls[:ne], epochs=ne)


def cd_fit(Xc, y, P, w, intercept, degree, loss_kind="squared", fit_linear=True, fit_intercept=True,
           max_iter=10, alpha0=1e-6, alpha=1e-3, beta=1e-3, tol=0.0, thr=1.0):
    """Xc: CSC triple (CSR class holding CSC arrays, n rows, d cols). P: model layout."""
    P = f64(P).copy()
    w = f64(w).copy()
    nO, k, dd = P.shape
    b = C.c_double(intercept)
    viol, ls, rg = np.zeros(max_iter), np.zeros(max_iter), np.zeros(max_iter)
    yp = np.zeros(Xc.n)
    ni = lib().ref_cd_fit(
        c_i64(Xc.n), c_i64(Xc.d), _d(Xc.data), _i(Xc.indices), _i(Xc.indptr), _d(f64(y)),
        c_int(degree), c_int(k), c_int(nO), c_int(dd - Xc.d), c_int(int(fit_linear)),
        c_int(int(fit_intercept)), _d(P), _d(w), C.byref(b), c_int(LOSS[loss_kind]), c_dbl(thr),
        c_int(max_iter), c_dbl(alpha0), c_dbl(alpha), c_dbl(beta), c_dbl(tol), _d(viol), _d(ls),
        _d(rg), _d(yp))
    return dict(P=P, w=w, intercept=b.value, viol=viol[:ni], loss=ls[:ni], reg=rg[:ni], iters=ni,
                y_pred=yp)


def pcd_fit(Xc, y, P, w, intercept, degree, loss_kind="squared", fit_linear=True, fit_intercept=True,
            max_iter=10, alpha0=1e-6, alpha=1e-3, beta=1e-4, gamma=1e-4, reg="squaredl12", tol=0.0, thr=1.0):
    """pcd.fit (optimizer/pcd.nim:108-200). Xc: CSC triple; P: model layout; reg in l1 / squaredl12 /
    squaredl12_rows."""
    P = f64(P).copy()
    w = f64(w).copy()
    nO, k, dd = P.shape
    b = C.c_double(intercept)
    viol, ls, rg = np.zeros(max_iter), np.zeros(max_iter), np.zeros(max_iter)
    yp = np.zeros(Xc.n)
    ni = lib().ref_pcd_fit(
        c_i64(Xc.n), c_i64(Xc.d), _d(Xc.data), _i(Xc.indices), _i(Xc.indptr), _d(f64(y)),
        c_int(degree), c_int(k), c_int(nO), c_int(dd - Xc.d), c_int(int(fit_linear)),
        c_int(int(fit_intercept)), _d(P), _d(w), C.byref(b), c_int(LOSS[loss_kind]), c_dbl(thr),
        c_int(max_iter), c_dbl(alpha0), c_dbl(alpha), c_dbl(beta), c_dbl(gamma), c_int(REG[reg]), c_dbl(tol),
        _d(viol), _d(ls), _d(rg), _d(yp))
    return dict(P=P, w=w, intercept=b.value, viol=viol[:ni], loss=ls[:ni], reg=rg[:ni], iters=ni,
                y_pred=yp)


def psgd_fit(X, y, P, w, intercept, degree, loss_kind="squared", fit_linear=True, fit_intercept=True,
             max_iter=10, eta0=0.01, alpha0=1e-6, alpha=1e-3, beta=1e-4, gamma=1e-4, reg="squaredl12",
             scheduling="optimal", power=1.0, tol=0.0, perms=None, it=1, thr=1.0):
    """PSGD.fit (optimizer/psgd.nim:76-215), X CSR; reg in l1 / l21 / squaredl12 / squaredl12_rows."""
    P = f64(P).copy()
    w = f64(w).copy()
    nO, k, dd = P.shape
    b, itc = C.c_double(intercept), C.c_int64(it)
    el = np.zeros(max_iter)
    if perms is not None:
        perms = i64(perms).reshape(-1, X.n)
        assert perms.shape[0] >= max_iter
    ne = lib().ref_psgd_fit(
        c_i64(X.n), c_i64(X.d), _d(X.data), _i(X.indices), _i(X.indptr), _d(f64(y)), c_int(degree),
        c_int(k), c_int(nO), c_int(dd - X.d), c_int(int(fit_linear)), c_int(int(fit_intercept)),
        _d(P), _d(w), C.byref(b), c_int(LOSS[loss_kind]), c_dbl(thr), c_int(max_iter), c_dbl(eta0),
        c_dbl(alpha0), c_dbl(alpha), c_dbl(beta), c_dbl(gamma), c_int(REG[reg]), c_int(SCHED[scheduling]),
        c_dbl(power), c_dbl(tol), _i(perms), C.byref(itc), _d(el))
    return dict(P=P, w=w, intercept=b.value, it=itc.value, epoch_loss=el[:ne], epochs=ne)


def regularization(P, w, intercept, alpha0, alpha, beta):
    P, w = f64(P), f64(w)
    return lib().ref_regularization(_d(P), c_i64(P.size), _d(w), c_i64(w.size), c_dbl(intercept),
                                    c_dbl(alpha0), c_dbl(alpha), c_dbl(beta))


# ------------------------------------------------------------------ FFM
def ffm_decision_function(X, P, w, intercept):
    P = f64(P)
    nF, d, k = P.shape
    out = np.zeros(X.n)
    lib().ref_ffm_decision_function(c_i64(X.n), c_i64(X.d), c_int(nF), c_int(k), _d(X.data),
                                    _i(X.indices), _i(X.indptr), _i(X.fields), _d(P), _d(f64(w)),
                                    c_dbl(intercept), _d(out))
    return out


def ffm_loss_grad(X, y, P, w, intercept, loss_kind="squared", row_begin=0, row_end=None,
                  fit_linear=True, fit_intercept=True, mini_batch_size=None, thr=1.0):
    P = f64(P)
    nF, d, k = P.shape
    if row_end is None:
        row_end = X.n
    mb = mini_batch_size if mini_batch_size is not None else (row_end - row_begin)
    gP, gw, gb = np.zeros_like(P), np.zeros(X.d), C.c_double(0.0)
    ypred = np.zeros(row_end - row_begin)
    dA = np.zeros_like(P)
    ls = lib().ref_ffm_loss_grad(c_i64(X.d), c_int(nF), c_int(k), _d(X.data), _i(X.indices),
                                 _i(X.indptr), _i(X.fields), _d(f64(y)), c_i64(row_begin),
                                 c_i64(row_end), c_int(int(fit_linear)), c_int(int(fit_intercept)),
                                 _d(P), _d(f64(w)), c_dbl(intercept), c_int(LOSS[loss_kind]),
                                 c_dbl(thr), c_i64(mb), _d(gP), _d(gw), C.byref(gb), _d(ypred), _d(dA))
    return dict(loss=ls, y_pred=ypred, gP=gP, gw=gw, gb=gb.value)


def ffm_adagrad_fit(X, y, P, w, intercept, loss_kind="squared", fit_linear=True, fit_intercept=True,
                    max_iter=10, eta0=0.1, alpha0=1e-6, alpha=1e-3, beta=1e-3, eps=1e-10, tol=0.0,
                    mini_batch_size=1, perms=None, it=1, state=None, thr=1.0):
    P = f64(P).copy()
    w = f64(w).copy()
    nF, d, k = P.shape
    b, itc = C.c_double(intercept), C.c_int64(it)
    gsP, gnP, gsw, gnw, gsb, gnb = _adagrad_state((nF, d, k), X.d, state)
    viol, ls = np.zeros(max_iter), np.zeros(max_iter)
    if perms is not None:
        perms = i64(perms).reshape(-1, X.n)
    ne = lib().ref_ffm_adagrad_fit(
        c_i64(X.n), c_i64(X.d), c_int(nF), c_int(k), _d(X.data), _i(X.indices), _i(X.indptr),
        _i(X.fields), _d(f64(y)), c_int(int(fit_linear)), c_int(int(fit_intercept)), _d(P), _d(w),
        C.byref(b), c_int(LOSS[loss_kind]), c_dbl(thr), c_int(max_iter), c_dbl(eta0), c_dbl(alpha0),
        c_dbl(alpha), c_dbl(beta), c_dbl(eps), c_dbl(tol), c_i64(mini_batch_size), _i(perms),
        C.byref(itc), _d(gsP), _d(gnP), _d(gsw), _d(gnw), C.byref(gsb), C.byref(gnb), _d(viol), _d(ls))
    st = dict(gsP=gsP, gnP=gnP, gsw=gsw, gnw=gnw, gsb=gsb.value, gnb=gnb.value)
    return dict(P=P, w=w, intercept=b.value, it=itc.value, viol=viol[:ne], loss=ls[:ne], epochs=ne,
                state=st)


def ffm_sgd_fit(X, y, P, w, intercept, loss_kind="squared", fit_linear=True, fit_intercept=True,
                max_iter=10, eta0=0.01, alpha0=1e-6, alpha=1e-3, beta=1e-3, scheduling="optimal",
                power=1.0, tol=0.0, perms=None, it=1, thr=1.0):
    P = f64(P).copy()
    w = f64(w).copy()
    nF, d, k = P.shape
    b, itc = C.c_double(intercept), C.c_int64(it)
    viol, ls = np.zeros(max_iter), np.zeros(max_iter)
    if perms is not None:
        perms = i64(perms).reshape(-1, X.n)
    ne = lib().ref_ffm_sgd_fit(
        c_i64(X.n), c_i64(X.d), c_int(nF), c_int(k), _d(X.data), _i(X.indices), _i(X.indptr),
        _i(X.fields), _d(f64(y)), c_int(int(fit_linear)), c_int(int(fit_intercept)), _d(P), _d(w),
        C.byref(b), c_int(LOSS[loss_kind]), c_dbl(thr), c_int(max_iter), c_dbl(eta0), c_dbl(alpha0),
        c_dbl(alpha), c_dbl(beta), c_int(SCHED[scheduling]), c_dbl(power), c_dbl(tol), _i(perms),
        C.byref(itc), _d(viol), _d(ls))
    return dict(P=P, w=w, intercept=b.value, it=itc.value, viol=viol[:ne], loss=ls[:ne], epochs=ne)


def hogwild_adagrad_epoch(X, y, Pf, w, intercept, degree, n_threads, is_ffm=False,
                          loss_kind="logistic", fit_linear=True, fit_intercept=True, eta0=0.1,
                          alpha0=1e-6, alpha=1e-3, beta=1e-3, eps=1e-10, n_rows=None, it=1):
    """TIMING ONLY (racy). Pf: solver layout (FM [nOrders,d+aug,k]; FFM [nFields,d,k]); modified in
    place together with w. Returns (loss_sum, viol)."""
    nO, dd, k = Pf.shape
    n_rows = X.n if n_rows is None else n_rows
    gsP, gnP = np.zeros_like(Pf), np.full_like(Pf, eps)
    gsw, gnw = np.zeros(X.d), np.full(X.d, eps)
    gsb, gnb, b = C.c_double(0.0), C.c_double(eps), C.c_double(intercept)
    itc, viol = C.c_int64(it), C.c_double(0.0)
    ls = lib().ref_hogwild_adagrad_epoch(
        c_int(int(is_ffm)), c_i64(n_rows), c_i64(X.d), _d(X.data), _i(X.indices), _i(X.indptr),
        _i(X.fields), _d(f64(y)), c_int(degree), c_int(k), c_int(0 if is_ffm else nO),
        c_int(0 if is_ffm else dd - X.d), c_int(nO if is_ffm else 0), c_int(int(fit_linear)),
        c_int(int(fit_intercept)), _d(Pf), _d(w), C.byref(b), c_int(LOSS[loss_kind]), c_dbl(1.0),
        c_dbl(eta0), c_dbl(alpha0), c_dbl(alpha), c_dbl(beta), C.byref(itc), _d(gsP), _d(gnP),
        _d(gsw), _d(gnw), C.byref(gsb), C.byref(gnb), None, c_int(n_threads), C.byref(viol))
    return ls, viol.value


# ---------------------------------------------------------------- text loaders (dataset.nim:562-990)
def _tokens(line):
    """the reference walks the line with parseFloat / parseInt, skipping ONE separator character after
    every number (dataset.nim:573-582); for well-formed lines that is a split on blanks and colons"""
    return line.replace(":", " ").split()


def load_svmlight(path, n_features=-1):
    """loadSVMLightFile -> CSR (dataset.nim:562-632): returns (CSR, y)"""
    data, indices, indptr, y = [], [], [0], []
    min_index, max_index = 1, 0                                   # :569-570
    for line in open(os.path.expanduser(path)).read().split("\n"):
        tok = _tokens(line)
        if not tok:
            continue
        y.append(float(tok[0]))
        for a in range(1, len(tok), 2):
            j = int(tok[a])
            min_index, max_index = min(j, min_index), max(j, max_index)
            indices.append(j)
            data.append(float(tok[a + 1]))
        indptr.append(len(indices))
    if min_index < 0:
        raise ValueError("Negative index is included.")           # :583-584
    offset = 0 if min_index == 0 else 1                           # :585
    pred = max_index + 1 - offset                                 # :586
    if n_features > 0 and pred > n_features:
        raise ValueError(f"nFeatures is {n_features} but dataset has at least {pred} features.")
    return CSR(data, np.array(indices, np.int64) - offset, indptr, len(y), max(pred, n_features)), f64(y)


def load_ffm(path, n_features=-1, n_fields=-1):
    """loadFFMFile (dataset.nim:696-790): returns (CSR with fields, y)"""
    data, indices, fields, indptr, y = [], [], [], [0], []
    min_index, max_index, min_field, max_field = 1, 0, 1, 1       # :705-708
    for line in open(os.path.expanduser(path)).read().split("\n"):
        tok = _tokens(line)
        if not tok:
            continue
        y.append(float(tok[0]))
        for a in range(1, len(tok), 3):
            f, j = int(tok[a]), int(tok[a + 1])
            min_field, max_field = min(f, min_field), max(f, max_field)
            min_index, max_index = min(j, min_index), max(j, max_index)
            fields.append(f)
            indices.append(j)
            data.append(float(tok[a + 2]))
        indptr.append(len(indices))
    if min_index < 0:
        raise ValueError("Negative index is included.")
    offset = 0 if min_index == 0 else 1
    offset_field = 0 if min_field == 0 else 1                     # :736
    pred, fpred = max_index + 1 - offset, max_field + 1 - offset_field
    if n_fields > 0 and fpred > n_fields:
        raise ValueError(f"nFields is {n_fields} but dataset has at least {fpred} fields.")
    if n_features > 0 and pred > n_features:
        raise ValueError(f"nFeatures is {n_features} but dataset has at least {pred} features.")
    return CSR(data, np.array(indices, np.int64) - offset, indptr, len(y), max(pred, n_features),
               fields=np.array(fields, np.int64) - offset_field, n_fields=max(fpred, n_fields)), f64(y)


def load_user_item_rating(path):
    """loadUserItemRatingFile -> CSR (dataset.nim:840-898)"""
    import re
    users, items, y = [], [], []
    for line in open(os.path.expanduser(path)).read().split("\n"):
        if len(line) < 5:                                         # :869-870
            continue
        nums = re.findall(r"\d+(?:\.\d+)?(?:[eE][-+]?\d+)?", line)
        users.append(int(nums[0]))
        items.append(int(nums[1]))
        y.append(float(nums[2]))
    n = len(y)
    min_user, max_user = min([1] + users), max([0] + users)       # :862-866
    min_item, max_item = min([1] + items), max([0] + items)
    n_users, n_items = max_user - min_user + 1, max_item - min_item + 1
    idx = np.empty(2 * n, np.int64)
    idx[0::2] = np.array(users, np.int64) - min_user
    idx[1::2] = np.array(items, np.int64) + n_users - min_item
    return CSR(np.ones(2 * n), idx, np.arange(n + 1) * 2, n, n_users + n_items if n else 0), f64(y)


# ---------------------------------------------------------------- synchronous-minibatch SGD
def get_eta(scheduling, eta0, power, reg, it):
    """getEta (optimizer/sgd.nim:60-69)"""
    if scheduling == "constant":
        return eta0
    if scheduling == "optimal":
        return eta0 / (1.0 + eta0 * reg * float(it)) ** power
    if scheduling == "invscaling":
        return eta0 / float(it) ** power
    return 1.0 / (reg * float(it))      # pegasos


def _shrink_pow(base, B):
    s = 1.0
    for _ in range(B):
        s *= base
    return s


def sgd_minibatch_fit(X, y, P, w, intercept, degree=2, loss_kind="squared", B=1, max_iter=3, eta0=0.01, alpha0=1e-6,
                      alpha=1e-3, beta=1e-3, scheduling="optimal", power=1.0, it=1, perms=None, fit_linear=True,
                      fit_intercept=True, ffm=False, thr=1.0):
    """The deterministic device analogue of the reference's Hogwild SGD (sgd_multi.nim:40-120,
    sgd_ffm_multi.nim:31-103), restated: the B samples of a minibatch see the same parameters (predictWithGrad,
    sgd.nim:191-202 / sgd_ffm.nim:11-30, summed with coef = dloss), then
        touched features (those occurring in the minibatch; all orders / fields, as update() sgd.nim:214-222):
            p <- (1 - eta_P beta)^B p - eta_P sum_i dL_i dA_i,   viol += |p_new - p|
        untouched:  p <- (1 - eta_P beta)^B p                    (the lazy scaling of sgd.nim:231-239)
        w likewise with (eta_w, alpha); b <- (1 - eta_b alpha0)^B b - eta_b sum_i dL_i;   it += B
    with the step sizes of the minibatch's first iteration.  B = 1 is step() itself.
    P: model layout ([nOrders, k, d+aug] for FM, [nFields, d, k] for FFM).  Returns P, w, intercept, it and the
    per-epoch viol / mean loss."""
    P, w, b = f64(P).copy(), f64(w).copy(), float(intercept)
    y = f64(y)
    n, d = X.n, X.d
    viols, losses = [], []
    order = np.arange(n)
    for ep in range(max_iter):
        if perms is not None:
            order = i64(perms[ep])
        viol, loss_sum = 0.0, 0.0
        for q0 in range(0, n, B):
            rows = order[q0:q0 + B]
            Bm = len(rows)
            sub = csr_take_rows(X, rows)
            if ffm:
                sub.fields = np.concatenate([X.fields[X.indptr[r]:X.indptr[r + 1]] for r in rows]).astype(np.int64) \
                    if Bm else np.zeros(0, np.int64)
                sub.n_fields = X.n_fields
                g = ffm_loss_grad(sub, y[rows], P, w, b, loss_kind, mini_batch_size=1, thr=thr)
            else:
                g = fm_loss_grad(sub, y[rows], P, w, b, degree, loss_kind, mini_batch_size=1, thr=thr)
            loss_sum += g["loss"]
            etaP, etaW, etaB = (get_eta(scheduling, eta0, power, r_, it) for r_ in (beta, alpha, alpha0))
            sP, sW, sB = _shrink_pow(1.0 - etaP * beta, Bm), _shrink_pow(1.0 - etaW * alpha, Bm), \
                _shrink_pow(1.0 - etaB * alpha0, Bm)
            touched = np.zeros(P.shape[1] if ffm else P.shape[2], bool)
            touched[np.unique(sub.indices)] = True
            if not ffm:
                touched[d:] = True                      # dummy features occur in every row (dataset.nim:182-189)
            if ffm:
                Pn = sP * P
                upd = sP * P[:, touched, :] - etaP * g["gP"][:, touched, :]
                viol += float(np.abs(upd - P[:, touched, :]).sum())
                Pn[:, touched, :] = upd
            else:
                Pn = sP * P
                upd = sP * P[:, :, touched] - etaP * g["gP"][:, :, touched]
                viol += float(np.abs(upd - P[:, :, touched]).sum())
                Pn[:, :, touched] = upd
            P = Pn
            if fit_linear:
                tw = touched[:d]
                wn = sW * w
                updw = sW * w[tw] - etaW * g["gw"][tw]
                viol += float(np.abs(updw - w[tw]).sum())
                wn[tw] = updw
                w = wn
            if fit_intercept:
                bn = sB * b - etaB * g["gb"]
                viol += abs(bn - b)
                b = bn
            it += Bm
        viols.append(viol)
        losses.append(loss_sum / n)
    return dict(P=P, w=w, intercept=b, it=it, viol=np.array(viols), loss=np.array(losses))
