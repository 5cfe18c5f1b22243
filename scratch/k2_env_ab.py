"""A/B of a row-kernel switch (default NIMFM_ROW_PREFETCH=0/1): K2 / K1 on the C4 and C3 shapes (whole shard, a row
list, one reference-default minibatch), the C4 AdaGrad epoch; the gradient of every setting is compared with the first."""
import ctypes as C, sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import nimfm_b200 as nf
from nimfm_b200 import _lib

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 6_000_000
var = sys.argv[2] if len(sys.argv) > 2 else "NIMFM_ROW_PREFETCH"
settings = sys.argv[3].split(",") if len(sys.argv) > 3 else ["0", "1", "0", "1"]
lib, ctx = _lib.load(), _lib.ctx()
data, indices, indptr, y = bench.gen_criteo_rows(rows, 1000)
ds = nf.newCSRDataset(data, indices, indptr, rows, bench.D_FEATURES); ds.set_targets(y)


def nrm(a, b):
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300))


def k2(h, shape, n, mb, reps, ref):
    for v in settings:
        os.environ[var] = v
        out = {"shape": shape, "rows": n, var: v}
        for grad in (1, 0):
            ms = C.c_float()
            _lib.check(lib.nimfm_fm_time_loss_grad(ctx, h, ds.handle(), 2, n, mb, reps, grad, C.byref(ms)))
            _lib.check(lib.nimfm_fm_time_loss_grad(ctx, h, ds.handle(), 2, n, mb, reps, grad, C.byref(ms)))
            out["grad_Mrows_s" if grad else "fwd_Mrows_s"] = round(n / ms.value / 1e3, 2)
        if ref is not None:
            ls = C.c_double()
            _lib.check(lib.nimfm_fm_loss_grad(ctx, h, ds.handle(), 2, 1.0, 0, n, None, n, 1, 0, C.byref(ls)))
            gP, gw, gb = np.zeros(ref["shapeP"]), np.zeros(bench.D_FEATURES), C.c_double()
            _lib.check(lib.nimfm_fm_get_grads(ctx, h, _lib.ptr(gP), _lib.ptr(gw), C.byref(gb)))
            if "gP" not in ref:
                ref.update(gP=gP, gw=gw, ls=ls.value)
            else:
                out.update(gP_err=nrm(gP, ref["gP"]), gw_err=nrm(gw, ref["gw"]), loss_rel=abs(ls.value - ref["ls"]) / abs(ref["ls"]))
        print(json.dumps(out), flush=True)


P, w, b = bench.model_params(7)
fm = nf.newFactorizationMachine(nf.classification, degree=3, nComponents=32)
fm.P, fm.w, fm.intercept, fm.isInitialized = P, w, b, True
h = fm._to_device(bench.D_FEATURES)
k2(h, "C4", rows, rows, 3, {"shapeP": P.shape})
k2(h, "C4 mb", 25641, 25641, 50, None)
lib.nimfm_fm_free(ctx, h)

rng = np.random.default_rng(2)
P3 = rng.standard_normal((1, 16, bench.D_FEATURES)) * 0.01
fm3 = nf.newFactorizationMachine(nf.classification, degree=2, nComponents=16)
fm3.P, fm3.w, fm3.intercept, fm3.isInitialized = P3, w, b, True
h3 = fm3._to_device(bench.D_FEATURES)
k2(h3, "C3", rows, rows, 3, {"shapeP": P3.shape})
k2(h3, "C3 mb", 25641, 25641, 50, None)
lib.nimfm_fm_free(ctx, h3)

for v in settings[:2]:
    os.environ[var] = v
    fa = nf.newFactorizationMachine(nf.classification, degree=3, nComponents=32, warmStart=True)
    fa.P, fa.w, fa.intercept, fa.isInitialized = P.copy(), w.copy(), 0.0, True
    opt = nf.newAdaGrad(maxIter=3, eta0=1e-4, loss=nf.Logistic(), verbose=0, tol=0.0, shuffle=False, miniBatchSize=1 << 19)
    opt.fit(ds, y, fa)
    print(json.dumps({"shape": "C4 adagrad", var: v, "Msamples_s": round(rows / min(opt.epoch_seconds) / 1e6, 2),
                      "loss": opt.history[-1] if opt.history else None, "P_abs_sum": float(np.abs(fa.P).sum())}), flush=True)
    f3 = nf.newFactorizationMachine(nf.classification, degree=2, nComponents=16, warmStart=True)
    f3.P, f3.w, f3.intercept, f3.isInitialized = P3.copy(), np.zeros(bench.D_FEATURES), 0.0, True
    o3 = nf.newMBPSGD(maxIter=3, eta0=0.1, alpha0=1e-6, alpha=1e-3, beta=1e-4, gamma=0.0, loss=nf.Logistic(), miniBatchSize=-1,
                      verbose=0, tol=0.0, shuffle=True)
    o3.fit(ds, y, f3)
    print(json.dumps({"shape": "C3 mbpsgd default mb (shuffled)", var: v, "Msamples_s": round(rows / min(o3.epoch_seconds) / 1e6, 2),
                      "loss": o3.history[-1] if o3.history else None, "P_abs_sum": float(np.abs(f3.P).sum())}), flush=True)
