"""end-to-end host-fed calls: device narrowing vs host staging team at several thread counts"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import nimfm_b200 as nf
from nimfm_b200 import _lib
n = int(os.environ.get("ROWS", 4_000_000))
data, indices, indptr, y = bench.gen_criteo_rows(n, 1)
P, w, b = bench.model_params(2)
fm = nf.newFactorizationMachine(nf.classification, degree=3, nComponents=32, warmStart=True)
fm.P, fm.w, fm.intercept, fm.isInitialized = P, w, 0.0, True
lib, ctx = _lib.load(), _lib.ctx()
h = fm._to_device(bench.D_FEATURES)
hb = [torch.from_numpy(a).pin_memory() for a in (data, indices, indptr, y)]
hp = [C.c_void_p(t.data_ptr()) for t in hb]
out = torch.empty(n, dtype=torch.float64).pin_memory()
ref_loss, ref_out = None, None
for chunk in (0, 1 << 17):
  for thr in ("0", "8"):
    os.environ["NIMFM_HOST_THREADS"] = thr
    ls = C.c_double()
    def grad():
        _lib.check(lib.nimfm_fm_loss_grad_host(ctx, h, n, bench.D_FEATURES, hp[0], hp[1], hp[2], hp[3], 2, 1.0, n, chunk, 1, 0, C.byref(ls)))
    def pred():
        _lib.check(lib.nimfm_fm_decision_function_host(ctx, h, n, bench.D_FEATURES, hp[0], hp[1], hp[2], chunk, C.c_void_p(out.data_ptr())))
    res = {}
    for name, f in (("grad", grad), ("pred", pred)):
        f(); f()
        t0 = time.perf_counter()
        for _ in range(3): f()
        res[name] = 3 * n / (time.perf_counter() - t0) / 1e6
    if ref_loss is None: ref_loss, ref_out = ls.value, out.numpy().copy()
    a, bb, t = C.c_int64(), C.c_int64(), C.c_int32()
    lib.nimfm_stream_stats(ctx, C.byref(a), C.byref(bb), C.byref(t))
    print(f"chunk {chunk:7d} threads {thr:>2s}: grad {res['grad']:6.1f} M/s  pred {res['pred']:6.1f} M/s  h2d(pred) {a.value/n:.0f} B/row  loss_rel {abs(ls.value-ref_loss)/abs(ref_loss):.1e} pred_same {np.array_equal(out.numpy(), ref_out)}", flush=True)
