"""torchrun: end-to-end host-fed predict+grad on every rank at once, host staging threads per rank varied"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
import nimfm_b200 as nf
from nimfm_b200 import _lib
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = int(os.environ.get("ROWS", 2_000_000))
data, indices, indptr, y = bench.gen_criteo_rows(n, 100 + rank)
P, w, b = bench.model_params(2)
fm = nf.newFactorizationMachine(nf.classification, degree=3, nComponents=32, warmStart=True)
fm.P, fm.w, fm.intercept, fm.isInitialized = P, w, 0.0, True
lib, ctx = _lib.load(), _lib.ctx(lr)
h = fm._to_device(bench.D_FEATURES)
hb = [torch.from_numpy(a).pin_memory() for a in (data, indices, indptr, y)]
hp = [C.c_void_p(t.data_ptr()) for t in hb]
out = torch.empty(n, dtype=torch.float64).pin_memory()
for thr in os.environ.get("THREADS", "0,2,1,4").split(","):
    os.environ["NIMFM_HOST_THREADS"] = thr
    ls = C.c_double()
    def grad():
        _lib.check(lib.nimfm_fm_loss_grad_host(ctx, h, n, bench.D_FEATURES, hp[0], hp[1], hp[2], hp[3], 2, 1.0, n * world, 0, 1, 0, C.byref(ls)))
    def pred():
        _lib.check(lib.nimfm_fm_decision_function_host(ctx, h, n, bench.D_FEATURES, hp[0], hp[1], hp[2], 0, C.c_void_p(out.data_ptr())))
    res = {}
    for name, f in (("grad", grad), ("pred", pred)):
        f()
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3): f()
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[name] = 3 * n * world / float(t.item()) / 1e6
    if rank == 0:
        print(f"world {world} host threads/rank {thr}: grad {res['grad']:.1f} M/s  pred {res['pred']:.1f} M/s (whole job)", flush=True)
dist.destroy_process_group()
