"""Row-sharded (data-parallel) paths on 2 GPUs of one box, one process per GPU over NCCL: batched
decisionFunction, MBPSGD (reduce-scatter -> sharded step -> all-gather, and the all-reduce route), synchronous-
minibatch AdaGrad / SGD (FM and FFM) -- each compared with the single-process oracle on the FULL data, with
uneven shards and shares (n % world != 0, miniBatchSize % world != 0).  The checks themselves live in
tests/sharded_parity.py; bench.py runs the same checks under torchrun at every N > 1, because the driver's
GPU-test box has one GPU and this test is skipped there (run it with `gpurun --gpus 2`)."""
import json
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from nimfm_b200 import distributed as nd
    import sharded_parity
    nd.init_comm(rank, world)
    res = sharded_parity.run_checks(rank, world)
    json.dump(res, open(os.path.join(out_dir, f"res{rank}.json"), "w"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_row_sharded_paths(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2); bench.py runs the same checks under torchrun (parity_check)")
    import torch.multiprocessing as mp
    from oracle import oracle as orc
    orc.build()
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        res = json.load(open(tmp_path / f"res{r}.json"))
        assert res["ok"], res["failed"]
        assert res["shard_rows"] == [49, 48] and res["shares"] == [4, 3]     # uneven on purpose


def test_one_gpu_sharded_parity_harness():
    """The same harness with a single rank (world = 1 communicator): keeps the checker itself exercised on the
    driver's one-GPU box."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from nimfm_b200 import distributed as nd
    import sharded_parity
    nd.init_comm(0, 1)
    res = sharded_parity.run_checks(0, 1)
    assert res["ok"], res["failed"]
