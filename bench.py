#!/usr/bin/env python
"""bench.py -- the headline measurement: FP64 HOFM degree-3 (rank 32, explicit lower orders)
predict + gradient over a synthetic Criteo-shaped CSR (39 nnz/row, 1M hashed features), BASELINE.json
config 4 / SURVEY 8d "C4", in samples/s.

  python bench.py --gpus N --steps K --warmup W          (N>1: launched by torch.distributed.run)
  python bench.py --impl reference ...                   (the reference's CPU path, oracle port)

A "step" is one predict+grad pass (zero grads -> fused forward + dloss + gradient scatter ->
loss/gb reduction [-> NCCL all-reduce of grad P/w when N>1]) over this rank's resident shard of
`--rows` rows (weak scaling: every rank holds its own 10M-row shard).  One JSON line is printed by
rank 0.  See DESIGN.md "Measurement" for the byte accounting behind `roofline`.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D_FEATURES = 1_000_000
Z = 39            # nnz per row: 13 numeric + 26 categorical slots
N_NUM, N_CAT = 13, 26
DEGREE, K, N_ORDERS = 3, 32, 2
# algorithmic bytes per row (indices counted at the device width 4, every gathered or
# scattered parameter element once, no cache credit)
B_FWD = Z * (8 + 4 + 8 + 8 * K * N_ORDERS) + 8 + 8
B_GRAD = B_FWD + 8 + Z * (8 * K * N_ORDERS + 8)


def gen_criteo_rows(n, seed, d=D_FEATURES, dist="criteo"):
    """Criteo-shaped synthetic CSR in the reference dtypes: 13 numeric slots (features 0..12, values
    U(0,1]) and 26 categorical slots, each with its own hashed id range and a Zipf(1.05)-like rank
    distribution (bounded inverse-CDF), value 1.0; indices sorted and unique within a row; y = +-1."""
    rng = np.random.default_rng(seed)
    R = (d - N_NUM) // N_CAT
    indices = np.empty((n, Z), dtype=np.int64)
    data = np.empty((n, Z), dtype=np.float64)
    indices[:, :N_NUM] = np.arange(N_NUM)[None, :]
    s = 1.05
    step = 1 << 20
    for a in range(0, n, step):
        b = min(n, a + step)
        data[a:b, :N_NUM] = 1.0 - rng.random((b - a, N_NUM))
        u = rng.random((b - a, N_CAT))
        rank = np.floor(((R ** (1.0 - s) - 1.0) * u + 1.0) ** (1.0 / (1.0 - s))).astype(np.int64) - 1
        np.clip(rank, 0, R - 1, out=rank)
        indices[a:b, N_NUM:] = N_NUM + np.arange(N_CAT)[None, :] * R + rank
    data[:, N_NUM:] = 1.0
    if dist != "criteo":   # experiments only (never the reported workload)
        R2 = d // Z
        for a in range(0, n, step):
            b = min(n, a + step)
            if dist == "uniform":      # every slot uniform over its own id range: no hot features at all
                rank = rng.integers(0, R2, size=(b - a, Z))
            else:                      # "zipfall": every slot Zipf-like, no always-present columns
                u = rng.random((b - a, Z))
                rank = np.floor(((R2 ** (1.0 - s) - 1.0) * u + 1.0) ** (1.0 / (1.0 - s))).astype(np.int64) - 1
                np.clip(rank, 0, R2 - 1, out=rank)
            indices[a:b] = np.arange(Z)[None, :] * R2 + rank
    indptr = np.arange(n + 1, dtype=np.int64) * Z
    y = np.where(rng.random(n) < 0.5, -1.0, 1.0)
    return data.reshape(-1), indices.reshape(-1), indptr, y


def model_params(seed, d=D_FEATURES):
    rng = np.random.default_rng(seed)
    P = rng.standard_normal((N_ORDERS, K, d)) * 0.01      # newFactorizationMachine scale=0.01
    w = rng.standard_normal(d) * 0.01
    return P, w, 0.0


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, or None"""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("fm_rows_grad_dram_bytes_per_row")
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_rate(orc, rows, seed, min_seconds, threads_note=True):
    """Time the oracle port of updateGradient (minibatch_psgd.nim:67-88 + sgd.nim:191-202), 1 thread,
    on `rows` rows of the same workload; returns (samples/s, rows used, seconds)."""
    from oracle.oracle import CSR
    data, indices, indptr, y = gen_criteo_rows(rows, seed)
    P, w, b = model_params(7)
    csr = CSR(data, indices, indptr, rows, D_FEATURES)
    Pf = orc.to_feature_major(P)
    del P
    gP, gw, gb = np.zeros_like(Pf), np.zeros(D_FEATURES), C.c_double(0.0)
    dA = np.zeros_like(Pf)
    lib = orc.lib()
    done, t0 = 0, time.perf_counter()
    chunk = max(1000, rows // 20)
    while done < rows:
        e = min(rows, done + chunk)
        lib.ref_fm_loss_grad(C.c_int64(D_FEATURES), orc._d(csr.data), orc._i(csr.indices), orc._i(csr.indptr),
                             orc._d(y), C.c_int64(done), C.c_int64(e), C.c_int(DEGREE), C.c_int(K), C.c_int(N_ORDERS),
                             C.c_int(0), C.c_int(1), C.c_int(1), orc._d(Pf), orc._d(w), C.c_double(b),
                             C.c_int(orc.LOSS["logistic"]), C.c_double(1.0), C.c_int64(rows), orc._d(gP), orc._d(gw),
                             C.byref(gb), None, orc._d(dA))
        done = e
        if time.perf_counter() - t0 > min_seconds and done >= chunk * 2:
            break
    dt = time.perf_counter() - t0
    return done / dt, done, dt


def cpu_hogwild_rate(orc, rows, seed, threads, reps=2):
    """Time the oracle port of the reference's multithreaded path for this workload: Hogwild AdaGrad
    epochs (adagrad_multi.nim:15-101: per sample predictWithGrad + lazy update + updateG, lock-free),
    `threads` threads over `rows` rows; optimizer state is allocated outside the timed call.
    Returns (samples/s of the best repetition, rows)."""
    from oracle.oracle import CSR
    data, indices, indptr, y = gen_criteo_rows(rows, seed)
    P, w, b = model_params(7)
    csr = CSR(data, indices, indptr, rows, D_FEATURES)
    Pf = orc.to_feature_major(P)
    del P
    lib = orc.lib()
    best = 0.0
    for _ in range(reps):
        gsP, gnP = np.zeros_like(Pf), np.full_like(Pf, 1e-10)
        gsw, gnw = np.zeros(D_FEATURES), np.full(D_FEATURES, 1e-10)
        gsb, gnb, bb = C.c_double(0.0), C.c_double(1e-10), C.c_double(b)
        itc, viol = C.c_int64(1), C.c_double(0.0)
        Pw, ww = Pf.copy(), w.copy()
        t0 = time.perf_counter()
        lib.ref_hogwild_adagrad_epoch(
            C.c_int(0), C.c_int64(rows), C.c_int64(D_FEATURES), orc._d(csr.data), orc._i(csr.indices),
            orc._i(csr.indptr), None, orc._d(y), C.c_int(DEGREE), C.c_int(K), C.c_int(N_ORDERS), C.c_int(0), C.c_int(0),
            C.c_int(1), C.c_int(1), orc._d(Pw), orc._d(ww), C.byref(bb), C.c_int(orc.LOSS["logistic"]), C.c_double(1.0),
            C.c_double(0.1), C.c_double(1e-6), C.c_double(1e-3), C.c_double(1e-3), C.byref(itc), orc._d(gsP), orc._d(gnP),
            orc._d(gsw), orc._d(gnw), C.byref(gsb), C.byref(gnb), None, C.c_int(threads), C.byref(viol))
        best = max(best, rows / (time.perf_counter() - t0))
    return best, rows


def hogwild_threads():
    """nThreads of sgd_multi.nim:13-18 for maxThreads < 0: 2 x countProcessors, capped by MaxThreadPoolSize (256)"""
    return min(2 * (os.cpu_count() or 1), 256)


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path.  The Nim toolchain is not
    in this image, so this is the oracle PORT of MBPSGD.updateGradient, which the reference runs on ONE
    thread (it has no multithreaded MBPSGD; its only threading is Hogwild SGD/AdaGrad)."""
    if rank != 0:
        return
    from oracle import oracle as orc
    orc.build()
    sample_rows = args.ref_rows
    # each step is a bounded sample; size it so the whole run ends within minutes
    rate0, used0, dt0 = cpu_port_rate(orc, min(sample_rows, 20000), 123, 1.0)
    per_step = int(max(2000, min(sample_rows, rate0 * 2.0)))    # ~2 s per step
    times = []
    for s in range(args.warmup + args.steps):
        r, used, dt = cpu_port_rate(orc, per_step, 123, 1e9)
        if s >= args.warmup:
            times.append(dt / used)
    per_row = float(np.mean(times))
    single = 1.0 / per_row
    # all the host threads the reference can use on this workload: its Hogwild AdaGrad (the C4 solver)
    T = hogwild_threads()
    hog, hog_rows = cpu_hogwild_rate(orc, min(args.ref_rows * 4, 400_000), 123, T)
    value = max(single, hog)
    line = {
        "impl": "reference", "metric": "samples/sec FM/HOFM predict+grad", "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_row * per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, world, per_step),
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": T if hog >= single else 1, "kind": "port",
                         "sample": f"the faster of (a) {per_step} rows/step through the oracle port of "
                                   "minibatch_psgd.updateGradient on 1 thread (the reference's MBPSGD is "
                                   f"single-threaded): {single:.0f} samples/s, and (b) {hog_rows} rows through the "
                                   f"oracle port of its Hogwild AdaGrad (adagrad_multi.nim) on {T} threads: "
                                   f"{hog:.0f} samples/s; Nim toolchain unavailable",
                         "single_thread_update_gradient": single, "hogwild_adagrad": hog, "hogwild_threads": T,
                         "host_cores": os.cpu_count()},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world, rows_per_step):
    return {"workload": "C4: synthetic Criteo-shaped CSR (39 nnz/row, 1M hashed features), HOFM degree 3 "
                        "rank 32 explicit lower orders, logistic loss, predict+grad "
                        "(sgd.predictWithGrad + minibatch_psgd.updateGradient)",
            "rows_per_gpu": rows_per_step, "nnz_per_row": Z, "n_features": D_FEATURES, "degree": DEGREE,
            "rank": K, "n_orders": N_ORDERS, "parallelism": f"row-sharded x{world}, NCCL allreduce of grad P/w",
            "l2": "inputs larger than L2 (no flush needed)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=10_000_000, help="rows per GPU (weak scaling)")
    ap.add_argument("--e2e-rows", type=int, default=0,
                    help="rows per end-to-end step (host buffers); default: the whole 10 M-row batch of the "
                         "device-resident step on one GPU, 2 M rows per rank under torchrun (pinned host memory "
                         "of 8 ranks on one box)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="bounded CPU-baseline sample")
    ap.add_argument("--ref-rows", type=int, default=100_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dist", default="criteo", choices=["criteo", "uniform", "zipfall"],
                    help="index distribution; anything but 'criteo' is an experiment, not the reported workload")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import nimfm_b200 as nf
    from nimfm_b200 import _lib

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _lib.load()
    ctx = _lib.ctx(local_rank)
    if world > 1:   # library-owned NCCL communicator; the 128-byte id travels over torch.distributed
        uid = (C.c_char * 128)()
        if rank == 0:
            _lib.check(lib.nimfm_comm_unique_id(uid))
        box = [bytes(uid)]
        dist.broadcast_object_list(box, src=0)
        uid = (C.c_char * 128).from_buffer_copy(box[0])
        _lib.check(lib.nimfm_comm_init(ctx, rank, world, uid))

    n = args.rows
    t_gen = time.perf_counter()
    data, indices, indptr, y = gen_criteo_rows(n, 1000 + rank, dist=args.dist)
    ds = nf.newCSRDataset(data, indices, indptr, n, D_FEATURES)
    ds.set_targets(y)
    P, w, b = model_params(7)
    fm = nf.newFactorizationMachine(nf.classification, degree=DEGREE, nComponents=K, fitLower=nf.explicit)
    fm.P, fm.w, fm.intercept, fm.isInitialized = P, w, b, True
    h = fm._to_device(D_FEATURES)
    t_gen = time.perf_counter() - t_gen
    loss_kind = nf.Logistic().kind
    mb_global = n * world

    def step():
        ls = C.c_double()
        _lib.check(lib.nimfm_fm_loss_grad(ctx, h, ds.handle(), loss_kind, 1.0, 0, n, None, mb_global, 1,
                                          int(world > 1), C.byref(ls)))
        return ls.value

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    l0 = _lib.launch_count()
    _lib.check(lib.nimfm_timer_start(ctx))
    for _ in range(args.steps):
        loss_sum = step()
    ms = C.c_float()
    _lib.check(lib.nimfm_timer_stop(ctx, C.byref(ms)))
    launches = _lib.launch_count() - l0
    barrier()
    clocks = sampler.stop()
    t = torch.tensor([ms.value], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = n * world / (ms_per_step / 1e3)

    # ---- dominant kernel alone (CUDA events on the launching stream inside the library)
    kms = C.c_float()
    _lib.check(lib.nimfm_fm_time_loss_grad(ctx, h, ds.handle(), loss_kind, n, mb_global, max(args.steps, 3), 1,
                                           C.byref(kms)))
    fms = C.c_float()
    _lib.check(lib.nimfm_fm_time_loss_grad(ctx, h, ds.handle(), loss_kind, n, mb_global, max(args.steps, 3), 0,
                                           C.byref(fms)))
    peak, peak_src = measured_peak()
    achieved = B_GRAD * n / (kms.value / 1e3) / 1e9
    traffic_row = ncu_traffic()
    # NOTE on frac > 1: `achieved` counts ALGORITHMIC bytes (every gathered / scattered parameter element
    # once, no cache credit).  On the Zipf-shaped workload ~2/3 of those sectors hit L2
    # (ncu: 11.9 KB/row of DRAM traffic vs 41 KB algorithmic), so the figure can exceed the DRAM copy peak.
    roofline = {"bound": "hbm", "kernel": "fm_rows_stream_kernel<3,true,MODE_GRAD,32>", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                "traffic": None if traffic_row is None else traffic_row * n,
                "algorithmic_bytes_per_row": B_GRAD, "kernel_ms": kms.value,
                "kernel_share_of_step": kms.value / ms_per_step,
                "frac_of_nominal_8TBs": achieved / 8000.0,
                "forward_only": {"kernel": "fm_rows_stream_kernel<3,true,MODE_PREDICT,32>", "kernel_ms": fms.value,
                                 "samples_per_s": n / (fms.value / 1e3), "algorithmic_bytes_per_row": B_FWD,
                                 "achieved": B_FWD * n / (fms.value / 1e3) / 1e9,
                                 "frac": B_FWD * n / (fms.value / 1e3) / 1e9 / peak}}

    # ---- end to end through the C ABI with HOST buffers (pinned), H2D inside the timed region
    ne = min(args.e2e_rows if args.e2e_rows > 0 else (10_000_000 if world == 1 else 2_000_000), n)
    hb = [torch.from_numpy(a).pin_memory() for a in (data[:ne * Z], indices[:ne * Z], indptr[:ne + 1], y[:ne])]
    hp = [C.c_void_p(t_.data_ptr()) for t_ in hb]
    h2d = sum(t_.numel() * t_.element_size() for t_ in hb)

    def e2e_step():
        ls = C.c_double()
        _lib.check(lib.nimfm_fm_loss_grad_host(ctx, h, ne, D_FEATURES, hp[0], hp[1], hp[2], hp[3], loss_kind, 1.0,
                                               ne * world, 0, 1, int(world > 1), C.byref(ls)))
        return ls.value

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_loss = e2e_step()
    torch.cuda.synchronize()
    te = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = ne * world * args.e2e_steps / float(te.item())

    def link_stats():
        a, b, t_ = C.c_int64(), C.c_int64(), C.c_int32()
        _lib.check(lib.nimfm_stream_stats(ctx, C.byref(a), C.byref(b), C.byref(t_)))
        return a.value, b.value, t_.value
    e2e_h2d, e2e_d2h, e2e_threads = link_stats()

    # end-to-end batched decisionFunction from the same host buffers (predictions read back)
    pred_host = torch.empty(ne, dtype=torch.float64).pin_memory()

    def e2e_predict():
        _lib.check(lib.nimfm_fm_decision_function_host(ctx, h, ne, D_FEATURES, hp[0], hp[1], hp[2], 0,
                                                       C.c_void_p(pred_host.data_ptr())))
    e2e_predict()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_predict()
    torch.cuda.synchronize()
    tp = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tp, op=dist.ReduceOp.MAX)
    e2e_predict_value = ne * world * args.e2e_steps / float(tp.item())
    pred_h2d, pred_d2h, _ = link_stats()

    line = {
        "metric": "samples/sec FM/HOFM predict+grad", "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, world, n),
        "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": int(e2e_h2d), "d2h_bytes_per_step": int(e2e_d2h),
                "rows_per_step": ne, "host_bytes_per_step": int(h2d), "host_staging_threads": int(e2e_threads),
                "note": "nimfm_fm_loss_grad_host: pinned host CSR (f64 data, i64 indices/indptr, f64 y) -> "
                "[ids narrowed to int32 by the library's host staging threads when host_staging_threads > 0] -> chunked "
                "H2D overlapped with the kernel -> loss read back; h2d/d2h bytes are what the library put on the "
                "link (nimfm_stream_stats), host_bytes_per_step the size of the caller's arrays"},
        "e2e_predict": {"value": e2e_predict_value, "unit": "samples/s", "what": "nimfm_fm_decision_function_host: the "
                        "same pinned host CSR -> chunked H2D overlapped with the forward kernel -> predictions copied "
                        "back", "h2d_bytes_per_step": int(pred_h2d), "d2h_bytes_per_step": int(pred_d2h)},
        "roofline": roofline,
        "loss_sum": loss_sum, "setup_seconds": t_gen,
    }
    if args.dist != "criteo":
        line["config"]["experiment_dist"] = args.dist
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as orc
        orc.build()
        rate, used, dt = cpu_port_rate(orc, 1_500_000, 1000, args.cpu_seconds)
        T = hogwild_threads()
        hog, hog_rows = cpu_hogwild_rate(orc, 300_000, 1000, T, reps=1)
        line["cpu_baseline"] = {"value": rate, "unit": "samples/s", "cores": 1, "kind": "port",
                                "hogwild_adagrad": {"value": hog, "threads": T, "rows": hog_rows,
                                                    "what": "oracle port of adagrad_multi.nim (the reference's "
                                                            "multithreaded path), racy by design, timed only"},
                                "sample": f"first {used} rows of the same workload in {dt:.1f} s, oracle port of "
                                          "minibatch_psgd.updateGradient + sgd.predictWithGrad (the reference "
                                          "runs MBPSGD on one thread)", "host_cores": os.cpu_count()}
    if rank == 0:
        print(json.dumps(line), flush=True)
    lib.nimfm_fm_free(ctx, h)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
