"""per-sample (sequential) solvers on the C4 shape: SGD and AdaGrad(miniBatchSize=1) samples/s"""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, nimfm_b200 as nf
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000
data, idx, ptr, y = bench.gen_criteo_rows(n, 2000)
ds = nf.newCSRDataset(data, idx, ptr, n, bench.D_FEATURES)
P, w, b = bench.model_params(7)
for name, mk in (("SGD", lambda: nf.newSGD(maxIter=2, eta0=0.01, loss=nf.Logistic(), verbose=0, tol=0.0, shuffle=False)),
                 ("AdaGrad mb=1", lambda: nf.newAdaGrad(maxIter=2, eta0=0.1, loss=nf.Logistic(), verbose=0, tol=0.0, shuffle=False, miniBatchSize=1)),
                 ("AdaGrad mb=256", lambda: nf.newAdaGrad(maxIter=2, eta0=0.01, loss=nf.Logistic(), verbose=0, tol=0.0, shuffle=False, miniBatchSize=256)),
                 ("AdaGrad mb=4096", lambda: nf.newAdaGrad(maxIter=2, eta0=0.01, loss=nf.Logistic(), verbose=0, tol=0.0, shuffle=False, miniBatchSize=4096))):
    fm = nf.newFactorizationMachine(nf.classification, degree=3, nComponents=32, warmStart=True)
    fm.P, fm.w, fm.intercept, fm.isInitialized = P.copy(), w.copy(), b, True
    opt = mk()
    opt.fit(ds, y, fm)
    ep = float(np.min(opt.epoch_seconds))
    print(json.dumps({"solver": name, "rows": n, "s_per_epoch": ep, "samples_per_s": n / ep, "hist": opt.history[-1]}), flush=True)
for name, reg in (("PSGD L1", nf.newL1), ("PSGD L21", nf.newL21)):
    fm = nf.newFactorizationMachine(nf.classification, degree=3, nComponents=32, warmStart=True)
    fm.P, fm.w, fm.intercept, fm.isInitialized = P.copy(), w.copy(), b, True
    opt = nf.newPSGD(maxIter=2, eta0=0.01, gamma=1e-5, loss=nf.Logistic(), reg=reg(), verbose=0, tol=0.0, shuffle=False)
    opt.fit(ds, y, fm)
    ep = float(np.min(opt.epoch_seconds))
    print(json.dumps({"solver": name, "rows": n, "s_per_epoch": ep, "samples_per_s": n / ep, "hist": opt.history[-1]}), flush=True)
