"""synchronous-minibatch SGD (fit(..., maxThreads=T)) on the C4 / C5 shapes vs the sequential kernel"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, bench_configs, nimfm_b200 as nf
n = 2_000_000
data, idx, ptr, y = bench.gen_criteo_rows(n, 2000)
ds = nf.newCSRDataset(data, idx, ptr, n, bench.D_FEATURES)
P, w, b = bench.model_params(3)
for T in (-1, 65536, 1 << 20):
    fm = nf.newFactorizationMachine(nf.classification, degree=3, nComponents=32, warmStart=True)
    fm.P, fm.w, fm.intercept, fm.isInitialized = P.copy(), w.copy(), 0.0, True
    opt = nf.newSGD(maxIter=2, eta0=1e-4, loss=nf.Logistic(), verbose=0, tol=0.0, shuffle=False)
    opt.fit(ds, y, fm, maxThreads=T)
    print(json.dumps({"config": "C4 SGD minibatch", "maxThreads": T, "samples_per_s": n / min(opt.epoch_seconds), "loss": opt.history[-1][1]}), flush=True)
ns = 20000
fm = nf.newFactorizationMachine(nf.classification, degree=3, nComponents=32, warmStart=True)
fm.P, fm.w, fm.intercept, fm.isInitialized = P.copy(), w.copy(), 0.0, True
opt = nf.newSGD(maxIter=1, eta0=1e-4, loss=nf.Logistic(), verbose=0, tol=0.0, shuffle=False)
opt.fit(ds[0:ns], y[:ns], fm)
print(json.dumps({"config": "C4 SGD sequential", "samples_per_s": ns / min(opt.epoch_seconds)}), flush=True)
del ds
nf_ = 400_000
data, idx, ptr, fields, y, d = bench_configs.gen_ffm_rows(nf_, 4000)
dsf = nf.newCSRFieldDataset(data, idx, ptr, fields, nf_, d, 39)
rng = np.random.default_rng(3)
for T in (4096, 65536):
    m = nf.newFieldAwareFactorizationMachine(nf.classification, nComponents=8, warmStart=True)
    m.P, m.w, m.intercept, m.isInitialized = rng.standard_normal((39, d, 8)) * 0.01, np.zeros(d), 0.0, True
    opt = nf.newSGD(maxIter=2, eta0=1e-4, loss=nf.Logistic(), verbose=0, tol=0.0, shuffle=False)
    opt.fit(dsf, y, m, maxThreads=T)
    print(json.dumps({"config": "C5 FFM SGD minibatch", "maxThreads": T, "samples_per_s": nf_ / min(opt.epoch_seconds), "loss": opt.history[-1][1]}), flush=True)
