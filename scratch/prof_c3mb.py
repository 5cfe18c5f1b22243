"""Profiling driver: one MBPSGD epoch slice on the C3 workload (FM degree 2 rank 16) at the
reference-default minibatch, for ncu (launch list / full capture of the row kernel)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, nimfm_b200 as nf
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
mb = int(sys.argv[2]) if len(sys.argv) > 2 else 25641
data, idx, ptr, y = bench.gen_criteo_rows(n, 2000)
ds = nf.newCSRDataset(data, idx, ptr, n, bench.D_FEATURES)
rng = np.random.default_rng(2)
fm = nf.newFactorizationMachine(nf.classification, degree=2, nComponents=16, warmStart=True)
fm.P, fm.w, fm.intercept, fm.isInitialized = rng.standard_normal((1, 16, bench.D_FEATURES)) * 0.01, np.zeros(bench.D_FEATURES), 0.0, True
opt = nf.newMBPSGD(maxIter=2, eta0=0.1, alpha0=1e-6, alpha=1e-3, beta=1e-4, gamma=0.0, loss=nf.Logistic(),
                   miniBatchSize=mb, verbose=0, tol=0.0, shuffle=False)
opt.fit(ds, y, fm)
print(opt.history, opt.epoch_seconds)
