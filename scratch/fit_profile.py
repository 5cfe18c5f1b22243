"""where the wall time of a whole fit() goes (host side): cProfile of MBPSGD.fit on 2 M Criteo-shaped rows"""
import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, nimfm_b200 as nf
from nimfm_b200 import _lib
n = 2_000_000
data, idx, ptr, y = bench.gen_criteo_rows(n, 2000)
_lib.ctx()
rng = np.random.default_rng(2)
P = rng.standard_normal((1, 16, bench.D_FEATURES)) * 0.01
def run():
    ds = nf.newCSRDataset(data, idx, ptr, n, bench.D_FEATURES)
    fm = nf.newFactorizationMachine(nf.classification, degree=2, nComponents=16, warmStart=True)
    fm.P, fm.w, fm.intercept, fm.isInitialized = P.copy(), np.zeros(bench.D_FEATURES), 0.0, True
    opt = nf.newMBPSGD(maxIter=2, eta0=0.1, gamma=0.0, loss=nf.Logistic(), miniBatchSize=1 << 20, verbose=0, tol=0.0, shuffle=False)
    t0 = time.perf_counter(); opt.fit(ds, y, fm); print("fit wall", time.perf_counter() - t0, "epochs", opt.epoch_seconds)
    t0 = time.perf_counter(); p = fm.decisionFunction(ds); print("decisionFunction wall", time.perf_counter() - t0)
run()
pr = cProfile.Profile(); pr.enable(); run(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
