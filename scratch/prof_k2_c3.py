"""Profiling driver: K2 / K1 on the C3 shape (FM degree 2 rank 16) at a reduced row count (for ncu)."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import nimfm_b200 as nf
from nimfm_b200 import _lib
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
lib, ctx = _lib.load(), _lib.ctx()
data, indices, indptr, y = bench.gen_criteo_rows(rows, 1000)
ds = nf.newCSRDataset(data, indices, indptr, rows, bench.D_FEATURES); ds.set_targets(y)
rng = np.random.default_rng(2)
P3 = rng.standard_normal((1, 16, bench.D_FEATURES)) * 0.01
fm = nf.newFactorizationMachine(nf.classification, degree=2, nComponents=16)
fm.P, fm.w, fm.intercept, fm.isInitialized = P3, np.zeros(bench.D_FEATURES), 0.0, True
h = fm._to_device(bench.D_FEATURES)
for grad in (1, 0):
    ms = C.c_float()
    _lib.check(lib.nimfm_fm_time_loss_grad(ctx, h, ds.handle(), 2, rows, rows, reps, grad, C.byref(ms)))
    bts = bench.B3_GRAD if grad else bench.B3_FWD
    print(f"C3 rows={rows} grad={grad}: {ms.value:.3f} ms/launch  {rows/ms.value/1e3:.1f} M rows/s  {bts*rows/ms.value/1e6:.0f} GB/s algorithmic")
