"""FFM predict+grad: RED route vs column route on the C5 shape (scratch A/B driver)"""
import sys, os, ctypes as C, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, nimfm_b200 as nf
from nimfm_b200 import _lib
lib, ctx = _lib.load(), _lib.ctx()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
data, idx, ptr, fields, y = bench.gen_ffm_rows(n, 6000)
ds = nf.newCSRFieldDataset(data, idx, ptr, fields, n, bench.D_FEATURES, 39); ds.set_targets(y); ds.handle()
base = np.random.default_rng(3).standard_normal((bench.D_FEATURES, 8)) * 0.01
P = np.stack([base * (1 + 0.01 * f) for f in range(39)])
m = nf.newFieldAwareFactorizationMachine(nf.classification, nComponents=8, warmStart=True)
m.P, m.w, m.intercept, m.isInitialized = P, np.zeros(bench.D_FEATURES), 0.0, True
h = m._to_device(ds)
ls = C.c_double()
def step(rows=n):
    _lib.check(lib.nimfm_ffm_loss_grad(ctx, h, ds.handle(), 2, 1.0, 0, rows, None, rows, 1, 0, C.byref(ls)))
for route in ("red", "cols"):
    os.environ["NIMFM_FFM_GRAD"] = route
    for rows in (n, min(n, 1 << 19), min(n, 1 << 16)):
        step(rows); step(rows)
        ms = C.c_float(); _lib.check(lib.nimfm_timer_start(ctx)); step(rows); step(rows); _lib.check(lib.nimfm_timer_stop(ctx, C.byref(ms)))
        print(route, "rows", rows, "ms/step %.2f" % (ms.value / 2), "M rows/s %.2f" % (rows / (ms.value / 2e3) / 1e6), "loss", ls.value, flush=True)
