"""FFM pair-block kernel: backward pass from shared memory (NIMFM_FFM_STAGE=1, default) vs re-gather (=0)"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench_configs, nimfm_b200 as nf
from nimfm_b200 import _lib
n = 400_000
data, idx, ptr, fields, y, d = bench_configs.gen_ffm_rows(n, 4000)
ds = nf.newCSRFieldDataset(data, idx, ptr, fields, n, d, 39)
ds.set_targets(y)
rng = np.random.default_rng(3)
m = nf.newFieldAwareFactorizationMachine(nf.classification, nComponents=8, warmStart=True)
m.P, m.w, m.intercept, m.isInitialized = rng.standard_normal((39, d, 8)) * 0.01, np.zeros(d), 0.0, True
lib, ctx = _lib.load(), _lib.ctx()
h = m._to_device(ds)
for rep in range(2):
    for st in ("1", "0"):
        os.environ["NIMFM_FFM_STAGE"] = st
        ms = C.c_float()
        _lib.check(lib.nimfm_ffm_time_loss_grad(ctx, h, ds.handle(), 2, n, n, 5, 1, C.byref(ms)))
        print(f"stage={st}: grad {ms.value:.3f} ms  {n / ms.value / 1e3:.2f} M rows/s", flush=True)
for st in ("1", "0"):
    os.environ["NIMFM_FFM_STAGE"] = st
    opt = nf.newAdaGrad(maxIter=2, eta0=1e-3, loss=nf.Logistic(), verbose=0, tol=0.0, shuffle=False, miniBatchSize=65536)
    m2 = nf.newFieldAwareFactorizationMachine(nf.classification, nComponents=8, warmStart=True)
    m2.P, m2.w, m2.intercept, m2.isInitialized = m.P.copy(), np.zeros(d), 0.0, True
    opt.fit(ds, y, m2)
    print(f"stage={st}: adagrad {n / min(opt.epoch_seconds) / 1e6:.2f} M samples/s loss {opt.history[-1][1]:.12f}", flush=True)
