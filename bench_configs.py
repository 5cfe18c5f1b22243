#!/usr/bin/env python
"""bench_configs.py -- the OTHER BASELINE.json configs (bench.py carries the headline C4 line):

  C1/C2  ML-100K-shaped one-hot user+item CSC (100k rows, 943+1682 features), FM degree 2 / HOFM
         degree 3, rank 30, CD, squared loss: seconds per CD epoch, GPU vs the oracle port (1 thread)
  C3     Criteo-shaped CSR, FM degree 2 rank 16, MBPSGD logistic: samples/s of a full epoch at the
         reference-default minibatch (d*n div nnz) and at 1 Mi rows per minibatch
  C4b    Criteo-shaped CSR, HOFM degree 3 rank 32, AdaGrad synchronous minibatch + batched decisionFunction
  C5     libffm-shaped field CSR (39 fields, one feature per field), FFM rank 8: predict+grad samples/s

One JSON object per config on stdout.  Sizes are reduced with --rows (default 2M) to keep the run
short; every line states its size.  Not the driver's bench contract -- that is bench.py.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def gen_ml100k(n=100_000, n_users=943, n_items=1682, seed=5):
    rng = np.random.default_rng(seed)
    u = rng.integers(0, n_users, n)
    v = rng.integers(0, n_items, n) + n_users
    idx = np.stack([u, v], axis=1).astype(np.int64)
    y = rng.integers(1, 6, n).astype(np.float64)
    return np.ones(2 * n), idx.reshape(-1), np.arange(n + 1, dtype=np.int64) * 2, y, n_users + n_items


def gen_ffm_rows(n, seed, n_fields=39, d=1_000_000):
    rng = np.random.default_rng(seed)
    R = d // n_fields
    s = 1.05
    u = rng.random((n, n_fields))
    rank = np.floor(((R ** (1.0 - s) - 1.0) * u + 1.0) ** (1.0 / (1.0 - s))).astype(np.int64) - 1
    np.clip(rank, 0, R - 1, out=rank)
    idx = np.arange(n_fields)[None, :] * R + rank
    data = np.where(rng.random((n, n_fields)) < 0.66, 1.0, rng.random((n, n_fields)))
    fields = np.tile(np.arange(n_fields, dtype=np.int64), n)
    y = np.where(rng.random(n) < 0.5, -1.0, 1.0)
    return data.reshape(-1), idx.reshape(-1), np.arange(n + 1, dtype=np.int64) * n_fields, fields, y, d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=2_000_000)
    ap.add_argument("--ffm-rows", type=int, default=200_000)
    ap.add_argument("--only", default="")
    ap.add_argument("--cpu", action="store_true", help="also time the oracle port on the host (1 thread)")
    args = ap.parse_args()
    import nimfm_b200 as nf
    from nimfm_b200 import _lib
    from oracle import oracle as orc
    from oracle.oracle import CSR
    lib, ctx = _lib.load(), _lib.ctx()
    peak, _ = bench.measured_peak()
    want = set(args.only.split(",")) if args.only else None

    def on(tag):
        return want is None or tag in want

    # ---------------------------------------------------------------- C1 / C2: CD
    for tag, degree in (("C1", 2), ("C2", 3)):
        if not on(tag):
            continue
        data, idx, ptr, y, d = gen_ml100k()
        n = len(y)
        csr = CSR(data, idx, ptr, n, d)
        csc = orc.csr_to_csc(csr)
        ds = nf.newCSCDataset(csc.data, csc.indices, csc.indptr, n, d)
        rng = np.random.default_rng(1)
        P = rng.standard_normal((degree - 1, 30, d)) * 0.01
        fm = nf.newFactorizationMachine(nf.regression, degree=degree, nComponents=30, warmStart=True)
        fm.P, fm.w, fm.intercept, fm.isInitialized = P.copy(), np.zeros(d), 0.0, True
        kw = dict(alpha0=1e-10, alpha=1e-10, beta=1e-3)
        opt = nf.newCD(maxIter=1, verbose=0, tol=0.0, **kw)
        opt.fit(ds, y, fm)                      # warm-up (builds batches, JIT-free but first-touch)
        fm.P, fm.w, fm.intercept = P.copy(), np.zeros(d), 0.0
        epochs = 5
        opt = nf.newCD(maxIter=epochs, verbose=0, tol=0.0, **kw)
        l0 = _lib.launch_count()
        t0 = time.perf_counter()
        opt.fit(ds, y, fm)
        dt = time.perf_counter() - t0
        launches = _lib.launch_count() - l0
        ep = float(np.median(opt.epoch_seconds))
        line = {"config": tag, "what": f"CD epoch, ML-100K shape n={n} d={d} nnz={2 * n}, degree {degree} rank 30",
                "gpu_s_per_epoch": ep, "fit_wall_s_per_epoch": dt / epochs,
                "note": "gpu_s_per_epoch = median wall time of the blocking nimfm_fm_cd_epoch call; fit_wall also "
                f"has upload/cd_begin/cd_end amortised over {epochs} epochs",
                "coordinate_steps_per_epoch": (degree - 1) * 30 * d + d + 1,
                "us_per_coordinate_step": ep / ((degree - 1) * 30 * d + d + 1) * 1e6,
                "kernel_launches_per_epoch": launches / epochs, "objective": opt.history[-1][1] + opt.history[-1][2]}
        if args.cpu:
            t0 = time.perf_counter()
            ref = orc.cd_fit(csc, y, P, np.zeros(d), 0.0, degree, "squared", max_iter=epochs, **kw)
            cdt = time.perf_counter() - t0
            line.update(cpu_s_per_epoch=cdt / epochs, cpu_kind="oracle port of cd.nim, 1 thread",
                        cpu_objective=ref["loss"][-1] + ref["reg"][-1],
                        objective_rel_err=abs(line["objective"] - (ref["loss"][-1] + ref["reg"][-1]))
                        / abs(ref["loss"][-1] + ref["reg"][-1]))
        print(json.dumps(line), flush=True)

    # ---------------------------------------------------------------- C3: MBPSGD epoch
    if on("C3"):
        n = args.rows
        data, idx, ptr, y = bench.gen_criteo_rows(n, 2000)
        ds = nf.newCSRDataset(data, idx, ptr, n, bench.D_FEATURES)
        rng = np.random.default_rng(2)
        P = rng.standard_normal((1, 16, bench.D_FEATURES)) * 0.01
        for mb_tag, mb in (("reference default d*n div nnz", -1), ("1Mi", 1 << 20)):
            fm = nf.newFactorizationMachine(nf.classification, degree=2, nComponents=16, warmStart=True)
            fm.P, fm.w, fm.intercept, fm.isInitialized = P.copy(), np.zeros(bench.D_FEATURES), 0.0, True
            opt = nf.newMBPSGD(maxIter=2, eta0=0.1, alpha0=1e-6, alpha=1e-3, beta=1e-4, gamma=0.0, loss=nf.Logistic(),
                               miniBatchSize=mb, verbose=0, tol=0.0, shuffle=False)
            rmb, inner = opt.resolve_sizes(ds)
            t0 = time.perf_counter()
            opt.fit(ds, y, fm)
            dt = time.perf_counter() - t0
            B = 39 * (8 + 4 + 8 + 8 * 16) + 16 + 8 + 39 * (8 * 16 + 8)
            print(json.dumps({"config": "C3", "what": f"MBPSGD fit (2 epochs), FM degree 2 rank 16, logistic, n={n}",
                              "miniBatchSize": rmb, "inner_iterations_per_epoch": inner, "mb_choice": mb_tag,
                              "samples_per_s": rmb * inner / float(np.min(opt.epoch_seconds)),
                              "seconds_per_epoch": float(np.min(opt.epoch_seconds)), "fit_wall_s": dt,
                              "note": "epoch = one blocking nimfm_fm_mbpsgd_epoch call (K2 + reduction + dense "
                                      "step per minibatch); fit_wall also has dataset/parameter upload+download",
                              "epoch_losses": opt.history, "algorithmic_bytes_per_row_sparse": B}), flush=True)
        ds.free()

    # ---------------------------------------------------------------- C4b: AdaGrad minibatch + decisionFunction
    if on("C4b"):
        n = args.rows
        data, idx, ptr, y = bench.gen_criteo_rows(n, 3000)
        ds = nf.newCSRDataset(data, idx, ptr, n, bench.D_FEATURES)
        P, w, b = bench.model_params(7)
        fm = nf.newFactorizationMachine(nf.classification, degree=3, nComponents=32, warmStart=True)
        fm.P, fm.w, fm.intercept, fm.isInitialized = P, w, b, True
        ds.handle()
        yhat = fm.decisionFunction(ds)          # warm-up
        t0 = time.perf_counter()
        yhat = fm.decisionFunction(ds)
        dt_pred = time.perf_counter() - t0
        # eta0 small: the synchronous variant's first step scales like eta0*sqrt(minibatch) (sums, not means)
        opt = nf.newAdaGrad(maxIter=1, eta0=1e-4, eps=1e-10, loss=nf.Logistic(), verbose=0, tol=0.0, shuffle=False,
                            miniBatchSize=1 << 20)
        t0 = time.perf_counter()
        opt.fit(ds, y, fm)
        dt = time.perf_counter() - t0
        cpu = {}
        if args.cpu:
            # the reference's multithreaded AdaGrad: Hogwild epochs (adagrad_multi.nim:15-101), T = 2*cores
            # (sgd_multi.nim:13-18) and T = cores, on a row sample of the same data (racy by design: timed only)
            csr = CSR(data, idx, ptr, n, bench.D_FEATURES)
            nn = min(n, 200_000)
            cores = os.cpu_count() or 1
            for T in sorted({cores, min(2 * cores, 256)}):
                Pf = orc.to_feature_major(P)
                t0 = time.perf_counter()
                orc.hogwild_adagrad_epoch(csr, y, Pf, w.copy(), b, 3, T, loss_kind="logistic", n_rows=nn)
                cpu[f"hogwild_adagrad_samples_per_s_T{T}"] = nn / (time.perf_counter() - t0)
            cpu["cpu_kind"] = (f"oracle port of adagrad_multi.nim (Hogwild, racy), first {nn} rows, host cores={cores}; "
                               "Nim toolchain unavailable")
        print(json.dumps({"config": "C4b", **cpu, "what": f"HOFM degree 3 rank 32, n={n}: decisionFunction (host result) "
                          "and one AdaGrad epoch with 1Mi-row synchronous minibatches",
                          "decision_function_samples_per_s": n / dt_pred,
                          "adagrad_samples_per_s": n / float(np.min(opt.epoch_seconds)), "adagrad_fit_wall_s": dt,
                          "note": "decisionFunction: wall clock incl. 512 MB parameter upload + result download; "
                                  "adagrad: the blocking nimfm_fm_adagrad_epoch call",
                          "adagrad_epoch_loss": opt.history[-1][1]}), flush=True)
        # SGD: the sequential solver (the reference's SGD.fit) on a row sample, and fit(..., maxThreads=T) -- the
        # reference's Hogwild entry point -- as synchronous minibatches of T samples (eta0 tiny: a minibatch moves
        # the parameters by T * eta * mean gradient, like T sequential steps)
        sgd = {}
        for T in (4096, 65536):
            f2 = nf.newFactorizationMachine(nf.classification, degree=3, nComponents=32, warmStart=True)
            f2.P, f2.w, f2.intercept, f2.isInitialized = P.copy(), w.copy(), b, True
            o2 = nf.newSGD(maxIter=2, eta0=1e-6, loss=nf.Logistic(), verbose=0, tol=0.0, shuffle=False)
            o2.fit(ds, y, f2, maxThreads=T)
            sgd[f"minibatch_{T}_samples_per_s"] = n / float(np.min(o2.epoch_seconds))
        ns = min(n, 20000)
        f2 = nf.newFactorizationMachine(nf.classification, degree=3, nComponents=32, warmStart=True)
        f2.P, f2.w, f2.intercept, f2.isInitialized = P.copy(), w.copy(), b, True
        o2 = nf.newSGD(maxIter=1, eta0=1e-6, loss=nf.Logistic(), verbose=0, tol=0.0, shuffle=False)
        o2.fit(ds[0:ns], y[:ns], f2)
        sgd["sequential_samples_per_s"] = ns / float(np.min(o2.epoch_seconds))
        print(json.dumps({"config": "C4c", "what": f"HOFM degree 3 rank 32, n={n}: SGD sequential (first {ns} rows) and "
                          "synchronous-minibatch SGD, the device analogue of Hogwild fit(..., maxThreads=T)", **sgd}),
              flush=True)
        ds.free()

    # ---------------------------------------------------------------- C5: FFM predict+grad
    if on("C5"):
        n = args.ffm_rows
        data, idx, ptr, fields, y, d = gen_ffm_rows(n, 4000)
        ds = nf.newCSRFieldDataset(data, idx, ptr, fields, n, d, 39)
        ds.set_targets(y)
        rng = np.random.default_rng(3)
        m = nf.newFieldAwareFactorizationMachine(nf.classification, nComponents=8, warmStart=True)
        m.P = rng.standard_normal((39, d, 8)) * 0.01
        m.w, m.intercept, m.isInitialized = np.zeros(d), 0.0, True
        h = m._to_device(ds)
        out = {}
        for grad in (1, 0):
            ms = C.c_float()
            _lib.check(lib.nimfm_ffm_time_loss_grad(ctx, h, ds.handle(), 2, n, n, 3, grad, C.byref(ms)))
            bts = 190_968 if grad else 95_800
            out["grad" if grad else "fwd"] = {"kernel_ms": ms.value, "samples_per_s": n / (ms.value / 1e3),
                                              "algorithmic_bytes_per_row": bts,
                                              "achieved_GBs": bts * n / (ms.value / 1e3) / 1e9,
                                              "frac_of_measured_hbm": bts * n / (ms.value / 1e3) / 1e9 / peak}
        lib.nimfm_ffm_free(ctx, h)
        # the config's solver: AdaGrad with synchronous minibatches (2 epochs; the second has the refresh pass)
        mbs = min(1 << 16, n)
        opt = nf.newAdaGrad(maxIter=2, eta0=1e-3, eps=1e-10, loss=nf.Logistic(), verbose=0, tol=0.0, shuffle=False,
                            miniBatchSize=mbs)
        t0 = time.perf_counter()
        opt.fit(ds, y, m)
        out["adagrad"] = {"miniBatchSize": mbs, "samples_per_s": n / float(np.min(opt.epoch_seconds)),
                          "epoch_seconds": opt.epoch_seconds, "fit_wall_s": time.perf_counter() - t0,
                          "epoch_loss": opt.history[-1][1],
                          "note": "one blocking nimfm_ffm_adagrad_epoch call per epoch (count -> refresh -> pair "
                                  "kernel -> apply per minibatch)"}
        line = {"config": "C5", "what": f"FFM 39 fields rank 8, one feature per field, d={d}, n={n}, logistic", **out}
        if args.cpu:
            csr = CSR(data, idx, ptr, n, d, fields=fields, n_fields=39)
            nn = min(n, 2000)
            t0 = time.perf_counter()
            orc.ffm_loss_grad(csr, y, m.P, m.w, 0.0, "logistic", row_end=nn, mini_batch_size=nn)
            line["cpu_samples_per_s"] = nn / (time.perf_counter() - t0)
            line["cpu_kind"] = "oracle port of sgd_ffm.predictWithGrad + scatter, 1 thread"
            # the config's solver on the CPU: Hogwild FFM AdaGrad (adagrad_ffm_multi.nim:16-104), T = 2*cores
            cores = os.cpu_count() or 1
            T = min(2 * cores, 256)
            nh = min(n, 20_000)
            Pf = np.ascontiguousarray(m.P)
            t0 = time.perf_counter()
            orc.hogwild_adagrad_epoch(csr, y, Pf, m.w.copy(), 0.0, 2, T, is_ffm=True, loss_kind="logistic", n_rows=nh)
            line[f"cpu_hogwild_adagrad_samples_per_s_T{T}"] = nh / (time.perf_counter() - t0)
            line["cpu_hogwild_kind"] = f"oracle port of adagrad_ffm_multi.nim (Hogwild, racy), first {nh} rows, host cores={cores}"
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
