"""Profiling driver: FFM AdaGrad minibatch epochs on the C5 shape (for the ncu launch list)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench_configs, nimfm_b200 as nf
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
mb = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
data, idx, ptr, fields, y, d = bench_configs.gen_ffm_rows(n, 4000)
ds = nf.newCSRFieldDataset(data, idx, ptr, fields, n, d, 39)
rng = np.random.default_rng(3)
m = nf.newFieldAwareFactorizationMachine(nf.classification, nComponents=8, warmStart=True)
m.P, m.w, m.intercept, m.isInitialized = rng.standard_normal((39, d, 8)) * 0.01, np.zeros(d), 0.0, True
opt = nf.newAdaGrad(maxIter=2, eta0=1e-3, eps=1e-10, loss=nf.Logistic(), verbose=0, tol=0.0, shuffle=False, miniBatchSize=mb)
opt.fit(ds, y, m)
print(opt.epoch_seconds, opt.history)
