"""one lazy / dense MBPSGD epoch at the reference-default minibatch on the C3 shape (for an ncu launch list)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, nimfm_b200 as nf
n = 1_000_000
data, idx, ptr, y = bench.gen_criteo_rows(n, 2000)
ds = nf.newCSRDataset(data, idx, ptr, n, bench.D_FEATURES)
rng = np.random.default_rng(2)
P = rng.standard_normal((1, 16, bench.D_FEATURES)) * 0.01
for lazy in ("1", "0"):
    os.environ["NIMFM_MBPSGD_LAZY"] = lazy
    fm = nf.newFactorizationMachine(nf.classification, degree=2, nComponents=16, warmStart=True)
    fm.P, fm.w, fm.intercept, fm.isInitialized = P.copy(), np.zeros(bench.D_FEATURES), 0.0, True
    opt = nf.newMBPSGD(maxIter=1, eta0=0.1, alpha0=1e-6, alpha=1e-3, beta=1e-4, gamma=0.0, loss=nf.Logistic(),
                       miniBatchSize=25641, verbose=0, tol=0.0, shuffle=False)
    opt.fit(ds, y, fm)
    print(lazy, opt.history)
