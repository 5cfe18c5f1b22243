"""Profiling driver: the deterministic predict+grad route (stash-forward row kernel + column kernel) on the C4 workload."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import nimfm_b200 as nf
from nimfm_b200 import _lib
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
os.environ["NIMFM_DETERMINISTIC"] = "1"
lib, ctx = _lib.load(), _lib.ctx()
data, indices, indptr, y = bench.gen_criteo_rows(rows, 1000)
ds = nf.newCSRDataset(data, indices, indptr, rows, bench.D_FEATURES); ds.set_targets(y)
P, w, b = bench.model_params(7)
fm = nf.newFactorizationMachine(nf.classification, degree=3, nComponents=32)
fm.P, fm.w, fm.intercept, fm.isInitialized = P, w, b, True
h = fm._to_device(bench.D_FEATURES)
ls = C.c_double()
for _ in range(2):
    _lib.check(lib.nimfm_fm_loss_grad(ctx, h, ds.handle(), 2, 1.0, 0, rows, None, rows, 1, 0, C.byref(ls)))
print("deterministic loss", ls.value)
