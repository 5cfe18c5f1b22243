"""host-fed predict+grad from caller-pinned arrays: chunk size sweep (packed values, 14 / 16 staging threads)"""
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
n = 10_000_000
data, indices, indptr, y = bench.gen_criteo_rows(n, 1)
hb = [torch.from_numpy(a).pin_memory() for a in (data, indices, indptr, y)]
hp = [C.c_void_p(t.data_ptr()) for t in hb]
for thr in ("14", "16"):
    for chunk in (1 << 15, 1 << 16, 1 << 17, 1 << 18, 1 << 19):
        os.environ["NIMFM_HOST_PACK"], os.environ["NIMFM_HOST_THREADS"] = "1", thr
        ls = C.c_double()
        def grad():
            _lib.check(lib.nimfm_fm_loss_grad_host(ctx, h, n, bench.D_FEATURES, hp[0], hp[1], hp[2], hp[3], 2, 1.0, n, chunk, 1, 0, C.byref(ls)))
        grad(); grad()
        ts = []
        for _ in range(4):
            t0 = time.perf_counter(); grad(); ts.append(time.perf_counter() - t0)
        print(f"threads {thr} chunk {chunk:7d}: best {n/min(ts)/1e6:6.1f} median {n/np.median(ts)/1e6:6.1f} M rows/s loss {ls.value:.6f}", flush=True)
