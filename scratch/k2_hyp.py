"""Which addresses make the busiest L2 slice of K2?  Same shape, the hot spots moved:
A criteo | B categorical head removed (rank += 256) | C the 13 always-present columns spread over 64 ids each |
D both.  Less locality everywhere but A: if B/C/D run FASTER, contention on a few lines bounds A."""
import ctypes as C, sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import nimfm_b200 as nf
from nimfm_b200 import _lib
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
lib, ctx = _lib.load(), _lib.ctx()
data, indices, indptr, y = bench.gen_criteo_rows(rows, 1000)
idx0 = indices.reshape(rows, bench.Z)
R = (bench.D_FEATURES - bench.N_NUM) // bench.N_CAT
P, w, b = bench.model_params(7)
fm = nf.newFactorizationMachine(nf.classification, degree=3, nComponents=32)
fm.P, fm.w, fm.intercept, fm.isInitialized = P, w, b, True
h = fm._to_device(bench.D_FEATURES)
def variant(name):
    idx = idx0.copy()
    if name in ("B", "D"):
        cat = idx[:, bench.N_NUM:] - bench.N_NUM
        slot, rank = cat // R, cat % R
        idx[:, bench.N_NUM:] = bench.N_NUM + slot * R + np.minimum(rank + 256, R - 1)
    if name in ("C", "D"):
        r = (np.arange(rows) % 64)[:, None]
        idx[:, :bench.N_NUM] = bench.N_NUM + np.arange(bench.N_NUM)[None, :] * R + (R - 1 - r)
    return idx.reshape(-1)
for name in ("A", "B", "C", "D", "A"):
    ds = nf.newCSRDataset(data, variant(name), indptr, rows, bench.D_FEATURES); ds.set_targets(y)
    out = {"variant": name, "info": {k: v for k, v in ds.info().items() if "ot" in k}}
    for grad in (1, 0):
        ms = C.c_float()
        _lib.check(lib.nimfm_fm_time_loss_grad(ctx, h, ds.handle(), 2, rows, rows, 3, grad, C.byref(ms)))
        out["grad" if grad else "fwd"] = round(rows / ms.value / 1e3, 1)
    print(json.dumps(out), flush=True)
    ds.free()
