"""C4 AdaGrad synchronous minibatch: one epoch for an ncu launch list; prints live epoch time first"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, nimfm_b200 as nf
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
mb = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
data, idx, ptr, y = bench.gen_criteo_rows(n, 2000)
ds = nf.newCSRDataset(data, idx, ptr, n, bench.D_FEATURES)
P, w, b = bench.model_params(3)
fm = nf.newFactorizationMachine(nf.classification, degree=3, nComponents=32, warmStart=True)
fm.P, fm.w, fm.intercept, fm.isInitialized = P, w, 0.0, True
opt = nf.newAdaGrad(maxIter=2, eta0=1e-4, loss=nf.Logistic(), verbose=0, tol=0.0, shuffle=False, miniBatchSize=mb)
opt.fit(ds, y, fm)
print("epoch_s", opt.epoch_seconds, "Msamples/s", n / min(opt.epoch_seconds) / 1e6)
