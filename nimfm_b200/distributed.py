"""One process per GPU: host-side plumbing for the row-sharded (data-parallel) paths.

The reference has no distributed backend at all (its only parallelism is Hogwild threads,
optimizer/sgd_multi.nim:86-95); this module is the synchronous replacement: every rank owns a
contiguous row shard of the CSR (== X[slice], tensor/sparse.nim:263-290), parameters are replicated,
and the library all-reduces grad P / w (MBPSGD) or the AdaGrad deltas over its own NCCL communicator.
torch.distributed is used only to carry the 128-byte NCCL unique id and for barriers.
"""
import ctypes as C

from . import _lib


def shard_rows(n, rank, world):
    """Contiguous row range [begin, end) of rank `rank` (rows split as evenly as possible)."""
    base, rem = divmod(int(n), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def exchange_unique_id(rank, make_id):
    """Rank 0 creates the id with `make_id()` (bytes); everybody returns the same bytes.
    Needs an initialised torch.distributed process group (any backend: gloo works on CPU)."""
    import torch.distributed as dist
    box = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return box[0]


def nccl_unique_id():
    uid = (C.c_char * 128)()
    _lib.check(_lib.load().nimfm_comm_unique_id(uid))
    return bytes(uid)


_state = {"rank": 0, "world": 1}


def rank():
    return _state["rank"]


def world():
    return _state["world"]


def init_comm(rank, world):
    """Create the library-owned NCCL communicator for this process' context.  After this call the
    solvers' fit() treat X as THIS rank's row shard, miniBatchSize as the GLOBAL size (each rank
    contributes local_batch(miniBatchSize, rank, world) rows per step) and all-reduce inside the library."""
    _state["rank"], _state["world"] = int(rank), int(world)
    if world == 1:
        _lib.check(_lib.load().nimfm_comm_init(_lib.ctx(), 0, 1, None))
        return
    raw = exchange_unique_id(rank, nccl_unique_id)
    uid = (C.c_char * 128).from_buffer_copy(raw)
    _lib.check(_lib.load().nimfm_comm_init(_lib.ctx(), rank, world, uid))


def allgather_i64(values):
    """[world, len(values)] int64 array of every rank's `values` (the library's communicator carries it)."""
    import numpy as np
    mine = np.ascontiguousarray(values, dtype=np.int64)
    out = np.empty((world(), mine.size), dtype=np.int64)
    _lib.check(_lib.load().nimfm_comm_allgather_i64(_lib.ctx(), _lib.ptr(mine), int(mine.size), _lib.ptr(out)))
    return out


def global_shape(X):
    """(nSamples, nnz) summed over the ranks' shards -- what the reference's size rules
    (minibatch_psgd.nim:157-165) see when X is the whole dataset."""
    if world() == 1:
        return int(X.nSamples), int(X.nnz)
    g = allgather_i64([X.nSamples, X.nnz]).sum(axis=0)
    return int(g[0]), int(g[1])


def local_batch(mini_batch_size, rank, world):
    """Rows of a global minibatch processed by this rank (the global size is what coef divides by,
    minibatch_psgd.nim:73)."""
    b, e = shard_rows(mini_batch_size, rank, world)
    return e - b
