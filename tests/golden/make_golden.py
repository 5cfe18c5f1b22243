"""Generates tests/golden/*.npz: known-answer vectors for the hot path.

The reference holds no golden vectors (its tests are differential) and cannot be
run here (no Nim toolchain), so the committed answers come from the BRUTE-FORCE definitions restated
from the reference's own test helpers (oracle/bruteforce.py: subset enumeration for ANOVA / FM / FFM,
naive dense solvers for CD / AdaGrad / SGD / MBPSGD / minibatch AdaGrad) -- independent of both oracle/ref_cpu.c and the CUDA
code, which are then both checked against these files.  Inputs use NumPy seeds; re-running this
script must reproduce the files bit for bit.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import bruteforce as bf  # noqa: E402
from helpers import make_dense, make_fm_params  # noqa: E402


def fm_cases():
    out = {}
    cid = 0
    for degree, fit_lower in [(2, "explicit"), (3, "explicit"), (3, "augment"), (4, "none"), (5, "explicit")]:
        n, d, k = 10, 7, 3
        X = make_dense(n, d, 100 + cid, density=0.7, positive=False)
        P, w, nA = make_fm_params(d, degree, k, fit_lower, True, seed=200 + cid, scale=0.3)
        b = 0.1 * (cid + 1)
        y = np.random.default_rng(300 + cid).standard_normal(n)
        yhat = bf.fm_decision_function(X, P, w, b, degree)
        grad = np.zeros_like(P)
        gw, gb = np.zeros(d), 0.0
        for i in range(n):
            dL = yhat[i] - y[i]
            bf.fm_grad(X, i, P, degree, dL / n, grad)
            gw += dL / n * X[i]
            gb += dL / n
        loss = float(np.sum(0.5 * (y - yhat) ** 2))
        pre = f"fm{cid}_"
        out.update({pre + "X": X, pre + "P": P, pre + "w": w, pre + "b": b, pre + "y": y, pre + "degree": degree,
                    pre + "fit_lower": fit_lower, pre + "yhat": yhat, pre + "gP": grad, pre + "gw": gw,
                    pre + "gb": gb, pre + "loss": loss})
        cid += 1
    out["n_cases"] = cid
    return out


def ffm_case():
    rng = np.random.default_rng(7)
    n, d, nF, k = 9, 12, 4, 3
    X = rng.random((n, d)) * (rng.random((n, d)) < 0.6)
    fields = np.arange(d) % nF
    P = rng.standard_normal((nF, d, k)) * 0.3
    w = rng.standard_normal(d) * 0.1
    y = rng.standard_normal(n)
    yhat = bf.ffm_decision_function(X, fields, P, w, -0.3)
    grad = np.zeros_like(P)
    for i in range(n):
        bf.ffm_grad(X, fields, i, P, (yhat[i] - y[i]) / n, grad)
    return dict(X=X, fields=fields, P=P, w=w, b=-0.3, y=y, yhat=yhat, gP=grad)


def solver_cases():
    out = {}
    n, d, k = 12, 5, 2
    X = make_dense(n, d, 400, density=0.7, positive=False)
    y = np.random.default_rng(401).standard_normal(n)
    for tag, degree, fit_lower in [("a", 2, "explicit"), ("b", 3, "explicit"), ("c", 3, "augment")]:
        P, w, nA = make_fm_params(d, degree, k, fit_lower, True, seed=402, scale=0.1)
        cdP, cdw, cdb = bf.cd_slow_fit(X, y, P, w, 0.0, degree, True, True, "squared", 2, 1e-6, 1e-3, 1e-3)
        agP, agw, agb = bf.adagrad_slow_fit(X, y, P, w * 0, 0.0, degree, True, True, "squared", 2, 0.1, 1e-6, 1e-3, 1e-3, 1e-10)
        sgP, sgw, sgb = bf.sgd_slow_fit(X, y, P, w, 0.0, degree, True, True, "squared", 2, 0.01, 1e-6, 1e-3, 1e-3)
        # synchronous minibatches of 5 samples (the device analogue of Hogwild fit(..., maxThreads=5))
        mbP, mbw, mbb = bf.sgd_minibatch_slow_fit(X, y, P, w, 0.0, degree, True, True, "squared", 5, 2, 0.01, 1e-6,
                                                  1e-3, 1e-3)
        out.update({f"{tag}_mbP": mbP, f"{tag}_mbw": mbw, f"{tag}_mbb": mbb})
        out.update({f"{tag}_degree": degree, f"{tag}_fit_lower": fit_lower, f"{tag}_P0": P, f"{tag}_w0": w,
                    f"{tag}_cdP": cdP, f"{tag}_cdw": cdw, f"{tag}_cdb": cdb,
                    f"{tag}_agP": agP, f"{tag}_agw": agw, f"{tag}_agb": agb,
                    f"{tag}_sgP": sgP, f"{tag}_sgw": sgw, f"{tag}_sgb": sgb})
    out.update(X=X, y=y)
    return out


def minibatch_solver_cases():
    """MBPSGD (L1 and no regulariser, minibatches of 5 over 12 rows: the cyclic order wraps) and the synchronous-
    minibatch AdaGrad (minibatches of 4) from their naive dense definitions"""
    out = {}
    n, d, k = 12, 5, 2
    X = make_dense(n, d, 410, density=0.7, positive=False)
    y = np.sign(np.random.default_rng(411).standard_normal(n))
    for tag, degree, fit_lower in [("a", 2, "explicit"), ("b", 3, "explicit"), ("c", 3, "augment")]:
        P, w, nA = make_fm_params(d, degree, k, fit_lower, True, seed=412, scale=0.3)
        for rtag, reg, gamma in (("id", "identity", 0.0), ("l1", "l1", 0.05)):
            mP, mw, mb, ml = bf.mbpsgd_slow_fit(X, y, P, w, 0.1, degree, True, True, "logistic", 3, 0.2, 1e-3, 1e-2,
                                                2e-2, gamma=gamma, reg=reg, mini_batch_size=5)
            out.update({f"{tag}_{rtag}_P": mP, f"{tag}_{rtag}_w": mw, f"{tag}_{rtag}_b": mb,
                        f"{tag}_{rtag}_loss": np.array(ml)})
        aP, aw, ab = bf.adagrad_minibatch_slow_fit(X, y, P, w * 0, 0.0, degree, True, True, "logistic", 4, 3, 0.1, 1e-3,
                                                   1e-2, 2e-2, 1e-10)
        out.update({f"{tag}_degree": degree, f"{tag}_fit_lower": fit_lower, f"{tag}_P0": P, f"{tag}_w0": w,
                    f"{tag}_agP": aP, f"{tag}_agw": aw, f"{tag}_agb": ab})
    out.update(X=X, y=y)
    return out


if __name__ == "__main__":
    np.savez(os.path.join(HERE, "minibatch_golden.npz"), **minibatch_solver_cases())
    np.savez(os.path.join(HERE, "fm_golden.npz"), **fm_cases())
    np.savez(os.path.join(HERE, "ffm_golden.npz"), **ffm_case())
    np.savez(os.path.join(HERE, "solver_golden.npz"), **solver_cases())
    print("wrote", [f for f in os.listdir(HERE) if f.endswith(".npz")])
