"""FFM predict+grad and the AdaGrad epoch: RED route vs TMA bulk-reduction route (NIMFM_FFM_BULK) on the C5 shape,
gradients of both compared."""
import sys, os, json, ctypes as C, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, nimfm_b200 as nf
from nimfm_b200 import _lib
lib, ctx = _lib.load(), _lib.ctx()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
data, idx, ptr, fields, y = bench.gen_ffm_rows(n, 6000)
ds = nf.newCSRFieldDataset(data, idx, ptr, fields, n, bench.D_FEATURES, 39); ds.set_targets(y); ds.handle()
rng = np.random.default_rng(3)
P = rng.standard_normal((39, bench.D_FEATURES, 8)) * 0.01
w0 = rng.standard_normal(bench.D_FEATURES) * 0.01
m = nf.newFieldAwareFactorizationMachine(nf.classification, nComponents=8, warmStart=True)
m.P, m.w, m.intercept, m.isInitialized = P, w0, 0.0, True
h = m._to_device(ds)
ls = C.c_double()
def step(rows=n):
    _lib.check(lib.nimfm_ffm_loss_grad(ctx, h, ds.handle(), 2, 1.0, 0, rows, None, rows, 1, 0, C.byref(ls)))
ref = None
for bulk in ("0", "1", "0", "1"):
    os.environ["NIMFM_FFM_BULK"] = bulk
    for rows in (n, min(n, 1 << 16)):
        step(rows); step(rows)
        ms = C.c_float(); _lib.check(lib.nimfm_timer_start(ctx)); step(rows); step(rows); _lib.check(lib.nimfm_timer_stop(ctx, C.byref(ms)))
        out = {"bulk": bulk, "rows": rows, "ms": round(ms.value / 2, 3), "Mrows_s": round(rows / (ms.value / 2e3) / 1e6, 2), "loss": ls.value}
        if rows == n:
            step(rows)
            gP = np.zeros_like(P); gw = np.zeros(bench.D_FEATURES)
            _lib.check(lib.nimfm_ffm_get_grads(ctx, h, _lib.ptr(gP), _lib.ptr(gw), None))
            if ref is None:
                ref = (gP, gw)
            else:
                out["gP_err"] = float(np.max(np.abs(gP - ref[0])) / np.max(np.abs(ref[0])))
                out["gw_err"] = float(np.max(np.abs(gw - ref[1])) / np.max(np.abs(ref[1])))
        print(json.dumps(out), flush=True)
lib.nimfm_ffm_free(ctx, h)
for bulk in ("0", "1"):
    os.environ["NIMFM_FFM_BULK"] = bulk
    ma = nf.newFieldAwareFactorizationMachine(nf.classification, nComponents=8, warmStart=True)
    ma.P, ma.w, ma.intercept, ma.isInitialized = P.copy(), np.zeros(bench.D_FEATURES), 0.0, True
    opt = nf.newAdaGrad(maxIter=2, eta0=1e-3, eps=1e-10, loss=nf.Logistic(), verbose=0, tol=0.0, shuffle=False, miniBatchSize=1 << 19)
    opt.fit(ds, y, ma)
    print(json.dumps({"adagrad bulk": bulk, "Msamples_s": round(n / min(opt.epoch_seconds) / 1e6, 2), "history": opt.history,
                      "P_abs_sum": float(np.abs(ma.P).sum())}), flush=True)
