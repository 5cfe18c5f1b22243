"""C3-shape K2 / K1 and MBPSGD epochs with whichever library is installed (A/B by swapping the .so)"""
import ctypes as C, sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import nimfm_b200 as nf
from nimfm_b200 import _lib
rows = 6_000_000
lib, ctx = _lib.load(), _lib.ctx()
data, indices, indptr, y = bench.gen_criteo_rows(rows, 1000)
ds = nf.newCSRDataset(data, indices, indptr, rows, bench.D_FEATURES); ds.set_targets(y)
rng = np.random.default_rng(2)
P3 = rng.standard_normal((1, 16, bench.D_FEATURES)) * 0.01
fm3 = nf.newFactorizationMachine(nf.classification, degree=2, nComponents=16)
fm3.P, fm3.w, fm3.intercept, fm3.isInitialized = P3, np.zeros(bench.D_FEATURES), 0.0, True
h3 = fm3._to_device(bench.D_FEATURES)
out = {"tag": sys.argv[1]}
for n, mb, reps, name in ((rows, rows, 3, "full"), (25641, 25641, 50, "mb25641"), (262144, 262144, 10, "mb256Ki")):
    for grad in (1, 0):
        ms = C.c_float()
        _lib.check(lib.nimfm_fm_time_loss_grad(ctx, h3, ds.handle(), 2, n, mb, reps, grad, C.byref(ms)))
        _lib.check(lib.nimfm_fm_time_loss_grad(ctx, h3, ds.handle(), 2, n, mb, reps, grad, C.byref(ms)))
        out[f"{name}_{'grad' if grad else 'fwd'}"] = round(n / ms.value / 1e3, 1)
lib.nimfm_fm_free(ctx, h3)
for mb in (-1, 1 << 18, 1 << 20):
    f3 = nf.newFactorizationMachine(nf.classification, degree=2, nComponents=16, warmStart=True)
    f3.P, f3.w, f3.intercept, f3.isInitialized = P3.copy(), np.zeros(bench.D_FEATURES), 0.0, True
    o3 = nf.newMBPSGD(maxIter=3, eta0=0.1, alpha0=1e-6, alpha=1e-3, beta=1e-4, gamma=0.0, loss=nf.Logistic(), miniBatchSize=mb,
                      verbose=0, tol=0.0, shuffle=False)
    o3.fit(ds, y, f3)
    out[f"epoch_mb{mb}"] = round(rows / min(o3.epoch_seconds) / 1e6, 1)
    out[f"loss_mb{mb}"] = o3.history[-1]
print(json.dumps(out), flush=True)
