"""all-reduce timing: library-owned NCCL communicator vs torch.distributed on the same payload"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import nimfm_b200 as nf
from nimfm_b200 import _lib, distributed as nd
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
lib, ctx = _lib.load(), _lib.ctx(lr)
nd.init_comm(rank, world)
d, k = 1_000_000, 32
fm = nf.newFactorizationMachine(nf.classification, degree=3, nComponents=k)
fm.P, fm.w, fm.intercept, fm.isInitialized = np.zeros((2, k, d)), np.zeros(d), 0.0, True
h = fm._to_device(d)
ds = nf.newCSRDataset([1.0], [0], [0, 1], 1, d); ds.set_targets([1.0])
def lib_ar():
    ls = C.c_double()
    _lib.check(lib.nimfm_fm_loss_grad(ctx, h, ds.handle(), 2, 1.0, 0, 0, None, 1, 0, 1, C.byref(ls)))
for _ in range(3): lib_ar()
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(10): lib_ar()
t_lib = (time.perf_counter() - t0) / 10
x = torch.zeros(2 * k * d + d + 2, dtype=torch.float64, device="cuda")
for _ in range(3): dist.all_reduce(x)
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(10): dist.all_reduce(x)
torch.cuda.synchronize()
t_torch = (time.perf_counter() - t0) / 10
nbytes = x.numel() * 8
if rank == 0:
    bus = lambda t: 2 * (world - 1) / world * nbytes / t / 1e9
    print(f"world={world} payload={nbytes/1e6:.0f} MB  lib allreduce {t_lib*1e3:.2f} ms ({bus(t_lib):.0f} GB/s bus)  torch {t_torch*1e3:.2f} ms ({bus(t_torch):.0f} GB/s bus)")
dist.destroy_process_group()
