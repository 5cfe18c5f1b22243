"""host-fed decisionFunction on pageable arrays in and out: per-call times"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import nimfm_b200 as nf
from nimfm_b200 import _lib
P, w, b = bench.model_params(2)
fm = nf.newFactorizationMachine(nf.classification, degree=3, nComponents=32, warmStart=True)
fm.P, fm.w, fm.intercept, fm.isInitialized = P, w, 0.0, True
lib, ctx = _lib.load(), _lib.ctx()
h = fm._to_device(bench.D_FEATURES)
n = int(os.environ.get("ROWS", "10000000"))
data, indices, indptr, y = bench.gen_criteo_rows(n, 1)
pp = [_lib.ptr(a) for a in (data, indices, indptr, y)]
for tag, out in (("pageable out", np.empty(n)), ("pinned out", torch.empty(n, dtype=torch.float64).pin_memory().numpy())):
    ts = []
    for _ in range(6):
        t0 = time.perf_counter()
        _lib.check(lib.nimfm_fm_decision_function_host(ctx, h, n, bench.D_FEATURES, pp[0], pp[1], pp[2], 0, _lib.ptr(out)))
        ts.append(time.perf_counter() - t0)
    print(tag, " ".join(f"{t*1e3:.0f}ms" for t in ts), f"-> {n/min(ts)/1e6:.1f} M rows/s best", flush=True)
ls = C.c_double()
ts = []
for _ in range(5):
    t0 = time.perf_counter()
    _lib.check(lib.nimfm_fm_loss_grad_host(ctx, h, n, bench.D_FEATURES, pp[0], pp[1], pp[2], pp[3], 2, 1.0, n, 0, 1, 0, C.byref(ls)))
    ts.append(time.perf_counter() - t0)
print("grad pageable", " ".join(f"{t*1e3:.0f}ms" for t in ts), flush=True)
