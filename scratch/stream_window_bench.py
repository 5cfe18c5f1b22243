"""StreamCSR file of the Criteo shape: whole-file resident load vs cacheSize windows (decisionFunction, one
MBPSGD epoch), rows/s from a file in the page cache"""
import os, sys, time, json, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, nimfm_b200 as nf
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
Z = 39
data, idx, ptr, y = bench.gen_criteo_rows(n, 77)
rec = np.zeros(n, dtype=[("cnt", "<i8"), ("e", [("val", "<f8"), ("id", "<i8")], (Z,))])
rec["cnt"] = Z
rec["e"]["val"] = data.reshape(n, Z)
rec["e"]["id"] = idx.reshape(n, Z)
d = tempfile.mkdtemp()
fx, fy = os.path.join(d, "c.bin"), os.path.join(d, "c.lab")
with open(fx, "wb") as f:
    f.write(b"STREAMCSR")
    f.write(np.array([n, bench.D_FEATURES, n * Z], dtype="<i8").tobytes())
    f.write(np.array([1.0, 0.0], dtype="<f8").tobytes())
    rec.tofile(f)
y.astype("<f8").tofile(fy)
size = os.path.getsize(fx)
P, w, b = bench.model_params(5)
fm = nf.newFactorizationMachine(nf.classification, degree=3, nComponents=32, warmStart=True)
fm.P, fm.w, fm.intercept, fm.isInitialized = P, w, 0.0, True
t0 = time.perf_counter(); whole = nf.newStreamCSRDataset(fx, resident=True); t_res = time.perf_counter() - t0
ref = fm.decisionFunction(whole)
t0 = time.perf_counter(); ref = fm.decisionFunction(whole); t_dev = time.perf_counter() - t0
out = {"rows": n, "file_GB": size / 1e9, "resident_load_s": t_res, "resident_load_GBps": size / t_res / 1e9,
       "decisionFunction_resident_Mrows_s": n / t_dev / 1e6}
for cache in (200, 1000):
    t0 = time.perf_counter(); win = nf.newStreamCSRDataset(fx, cacheSize=cache, resident=False); t_open = time.perf_counter() - t0
    got = fm.decisionFunction(win)
    t0 = time.perf_counter(); got = fm.decisionFunction(win); t_win = time.perf_counter() - t0
    out[f"windowed_{cache}MB"] = {"open_s": t_open, "windows": sum(1 for _ in [0] for _ in range(0)) or -(-size // (cache << 20)),
                                  "decisionFunction_Mrows_s": n / t_win / 1e6, "file_GBps": size / t_win / 1e9,
                                  "bit_equal": bool(np.array_equal(got, ref))}
    win.close()
# one MBPSGD epoch (C3 model) windowed vs resident
rng = np.random.default_rng(2)
P3 = rng.standard_normal((1, 16, bench.D_FEATURES)) * 0.01
res = {}
for name, X in (("resident", whole), ("windowed_200MB", nf.newStreamCSRDataset(fx, cacheSize=200, resident=False))):
    f3 = nf.newFactorizationMachine(nf.classification, degree=2, nComponents=16, warmStart=True)
    f3.P, f3.w, f3.intercept, f3.isInitialized = P3.copy(), np.zeros(bench.D_FEATURES), 0.0, True
    opt = nf.newMBPSGD(maxIter=2, eta0=0.1, gamma=0.0, loss=nf.Logistic(), miniBatchSize=1 << 17, verbose=0, tol=0.0, shuffle=False)
    opt.fit(X, y, f3)
    res[name] = {"epoch_s": float(np.min(opt.epoch_seconds)), "Msamples_s": n / float(np.min(opt.epoch_seconds)) / 1e6, "loss": opt.history[-1]}
out["mbpsgd_epoch"] = res
print(json.dumps(out))
os.remove(fx); os.remove(fy)
