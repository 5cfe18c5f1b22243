"""NIMFM_LAZY_FUSED=0/1: the C3 MBPSGD epoch at the reference-default minibatch (and 8 192 rows), the C4 minibatch SGD at
maxThreads<0 (4 096 rows); iterates of both settings compared"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, nimfm_b200 as nf
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
data, idx, ptr, y = bench.gen_criteo_rows(n, 2000)
ds = nf.newCSRDataset(data, idx, ptr, n, bench.D_FEATURES)
rng = np.random.default_rng(2)
P3 = rng.standard_normal((1, 16, bench.D_FEATURES)) * 0.01
P4, w4, b4 = bench.model_params(3)
res = {}
for fused in ("0", "1", "0", "1"):
    os.environ["NIMFM_LAZY_FUSED"] = fused
    for mb in (-1, 1 << 13):
        fm = nf.newFactorizationMachine(nf.classification, degree=2, nComponents=16, warmStart=True)
        fm.P, fm.w, fm.intercept, fm.isInitialized = P3.copy(), np.zeros(bench.D_FEATURES), 0.0, True
        opt = nf.newMBPSGD(maxIter=3, eta0=0.1, alpha0=1e-6, alpha=1e-3, beta=1e-4, gamma=0.0, loss=nf.Logistic(),
                           miniBatchSize=mb, verbose=0, tol=0.0, shuffle=False)
        rmb, inner = opt.resolve_sizes(ds)
        opt.fit(ds, y, fm)
        ep = float(np.min(opt.epoch_seconds))
        key = ("c3", mb)
        out = {"cfg": "C3 MBPSGD", "fused": fused, "mb": rmb, "us_per_minibatch": round(ep / inner * 1e6, 2),
               "Msamples_s": round(rmb * inner / ep / 1e6, 2), "loss": opt.history[-1]}
        if key in res:
            out["P_maxdiff"] = float(np.max(np.abs(fm.P - res[key][0]))); out["w_maxdiff"] = float(np.max(np.abs(fm.w - res[key][1])))
            out["b_diff"] = abs(fm.intercept - res[key][2])
        else:
            res[key] = (fm.P.copy(), fm.w.copy(), fm.intercept)
        print(json.dumps(out), flush=True)
    fm = nf.newFactorizationMachine(nf.classification, degree=3, nComponents=32, warmStart=True)
    fm.P, fm.w, fm.intercept, fm.isInitialized = P4.copy(), w4.copy(), 0.0, True
    opt = nf.newSGD(maxIter=2, eta0=1e-4, loss=nf.Logistic(), verbose=0, tol=0.0, shuffle=False)
    opt.fit(ds, y, fm, maxThreads=-1)
    out = {"cfg": "C4 SGD minibatch 4096", "fused": fused, "Msamples_s": round(n / min(opt.epoch_seconds) / 1e6, 2), "hist": opt.history[-1]}
    key = ("c4",)
    if key in res:
        out["P_maxdiff"] = float(np.max(np.abs(fm.P - res[key][0]))); out["w_maxdiff"] = float(np.max(np.abs(fm.w - res[key][1])))
    else:
        res[key] = (fm.P.copy(), fm.w.copy(), fm.intercept)
    print(json.dumps(out), flush=True)
