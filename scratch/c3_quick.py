"""C3 MBPSGD epoch timing at the reference-default minibatch and per-kernel device time"""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, nimfm_b200 as nf
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
data, idx, ptr, y = bench.gen_criteo_rows(n, 2000)
ds = nf.newCSRDataset(data, idx, ptr, n, bench.D_FEATURES)
rng = np.random.default_rng(2)
P = rng.standard_normal((1, 16, bench.D_FEATURES)) * 0.01
for lazy, mb in ((None, -1), ("0", -1), ("1", 1 << 13), ("0", 1 << 13), ("1", 1 << 15), ("0", 1 << 15), ("1", 1 << 17), ("0", 1 << 17), (None, 1 << 20)):
    os.environ.pop("NIMFM_MBPSGD_LAZY", None)
    if lazy is not None:
        os.environ["NIMFM_MBPSGD_LAZY"] = lazy
    fm = nf.newFactorizationMachine(nf.classification, degree=2, nComponents=16, warmStart=True)
    fm.P, fm.w, fm.intercept, fm.isInitialized = P.copy(), np.zeros(bench.D_FEATURES), 0.0, True
    opt = nf.newMBPSGD(maxIter=3, eta0=0.1, alpha0=1e-6, alpha=1e-3, beta=1e-4, gamma=0.0, loss=nf.Logistic(),
                       miniBatchSize=mb, verbose=0, tol=0.0, shuffle=False)
    rmb, inner = opt.resolve_sizes(ds)
    opt.fit(ds, y, fm)
    ep = float(np.min(opt.epoch_seconds))
    print(json.dumps({"lazy": lazy, "mb": rmb, "inner": inner, "s_per_epoch": ep, "us_per_minibatch": ep / inner * 1e6,
                      "samples_per_s": rmb * inner / ep, "loss": opt.history[-1], "P_abs_sum": float(np.abs(fm.P).sum())}), flush=True)
