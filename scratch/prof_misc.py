"""Profiling driver for the secondary kernels: K3 dense step + SquaredL12 prox (MBPSGD, C3 shape), K4/K5
AdaGrad minibatch (C4 shape), CD column kernels (C1 shape)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, bench_configs, nimfm_b200 as nf
from oracle import oracle as orc
from oracle.oracle import CSR
what = sys.argv[1] if len(sys.argv) > 1 else "all"
if what in ("all", "mbpsgd"):
    n = 300_000
    data, idx, ptr, y = bench.gen_criteo_rows(n, 2000)
    ds = nf.newCSRDataset(data, idx, ptr, n, bench.D_FEATURES)
    rng = np.random.default_rng(2)
    fm = nf.newFactorizationMachine(nf.classification, degree=2, nComponents=16, warmStart=True)
    fm.P, fm.w, fm.intercept, fm.isInitialized = rng.standard_normal((1, 16, bench.D_FEATURES)) * 0.01, np.zeros(bench.D_FEATURES), 0.0, True
    nf.newMBPSGD(maxIter=1, eta0=0.1, gamma=1e-3, loss=nf.Logistic(), miniBatchSize=1 << 17, verbose=0, tol=0.0,
                 shuffle=False).fit(ds, y, fm)          # default reg: SquaredL12 column prox
if what in ("all", "adagrad"):
    n = 600_000
    data, idx, ptr, y = bench.gen_criteo_rows(n, 3000)
    ds = nf.newCSRDataset(data, idx, ptr, n, bench.D_FEATURES)
    P, w, b = bench.model_params(7)
    fm = nf.newFactorizationMachine(nf.classification, degree=3, nComponents=32, warmStart=True)
    fm.P, fm.w, fm.intercept, fm.isInitialized = P, w, b, True
    nf.newAdaGrad(maxIter=1, eta0=1e-4, loss=nf.Logistic(), verbose=0, tol=0.0, shuffle=False, miniBatchSize=200_000).fit(ds, y, fm)
if what in ("all", "cd"):
    os.environ["NIMFM_CD_GRAPH"] = "0"
    data, idx, ptr, y, d = bench_configs.gen_ml100k()
    n = len(y)
    csc = orc.csr_to_csc(CSR(data, idx, ptr, n, d))
    ds = nf.newCSCDataset(csc.data, csc.indices, csc.indptr, n, d)
    fm = nf.newFactorizationMachine(nf.regression, degree=3, nComponents=4, warmStart=True)
    fm.P, fm.w, fm.intercept, fm.isInitialized = np.random.default_rng(1).standard_normal((2, 4, d)) * 0.01, np.zeros(d), 0.0, True
    nf.newCD(maxIter=1, verbose=0, tol=0.0, alpha0=1e-10, alpha=1e-10, beta=1e-3).fit(ds, y, fm)
print("done")
