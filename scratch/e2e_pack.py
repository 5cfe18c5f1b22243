"""host-fed predict+grad from caller-pinned arrays: raw vs packed value transport at several team sizes and batch sizes"""
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
print("cores", os.cpu_count(), flush=True)
for n in (int(x) for x in os.environ.get("ROWS", "4000000,10000000").split(",")):
    data, indices, indptr, y = bench.gen_criteo_rows(n, 1)
    hb = [torch.from_numpy(a).pin_memory() for a in (data, indices, indptr, y)]
    hp = [C.c_void_p(t.data_ptr()) for t in hb]
    for pack, thr in (("0", "8"), ("0", "14"), ("1", "8"), ("1", "12"), ("1", "14"), ("1", "16")):
        os.environ["NIMFM_HOST_PACK"], os.environ["NIMFM_HOST_THREADS"] = pack, thr
        ls = C.c_double()
        def grad():
            _lib.check(lib.nimfm_fm_loss_grad_host(ctx, h, n, bench.D_FEATURES, hp[0], hp[1], hp[2], hp[3], 2, 1.0, n, 0, 1, 0, C.byref(ls)))
        grad(); grad()
        ts = []
        for _ in range(4):
            t0 = time.perf_counter(); grad(); ts.append(time.perf_counter() - t0)
        a, bb, t = C.c_int64(), C.c_int64(), C.c_int32()
        lib.nimfm_stream_stats(ctx, C.byref(a), C.byref(bb), C.byref(t))
        print(f"rows {n:9d} pack {pack} threads {thr:>2s}: best {n/min(ts)/1e6:6.1f} median {n/np.median(ts)/1e6:6.1f} M rows/s  h2d {a.value/n:.0f} B/row loss {ls.value:.6f}", flush=True)
    del hb
