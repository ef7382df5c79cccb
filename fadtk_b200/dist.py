"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink / NVSwitch).

The reference has no distributed code at all.  Clips are independent, so
ranks embed disjoint shards and the only exchange on the path is ONE all-reduce(sum) of the
packed fp64 statistics ``[n, sum, outer]`` (d^2 + d + 1 values, 132 KB at d = 128) per dataset.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as td


def is_distributed() -> bool:
    return td.is_available() and td.is_initialized() and td.get_world_size() > 1


def rank() -> int:
    return td.get_rank() if (td.is_available() and td.is_initialized()) else 0


def world_size() -> int:
    return td.get_world_size() if (td.is_available() and td.is_initialized()) else 1


def init_from_env(backend: str | None = None) -> bool:
    """Initialise from torchrun's env (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*)."""
    if td.is_initialized():
        return True
    if int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        return False
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        td.init_process_group(backend=backend, device_id=torch.device("cuda", local))
    else:
        td.init_process_group(backend=backend)
    return True


def shutdown():
    if td.is_available() and td.is_initialized():
        td.destroy_process_group()


def shard(items, r: int | None = None, w: int | None = None):
    """Contiguous split like np.array_split (the reference's file sharding, fad_batch.py:43)."""
    r = rank() if r is None else r
    w = world_size() if w is None else w
    n = len(items)
    base, extra = divmod(n, w)
    start = r * base + min(r, extra)
    return items[start:start + base + (1 if r < extra else 0)]


_native_engine = None


def enable_native_allreduce(engine) -> bool:
    """Create the C ABI's own NCCL communicator (include/fadtk_b200.h: fad_comm_unique_id / fad_comm_init) for this
    process group and route the statistics all-reduce through it, so that the one collective of the path runs behind
    the boundary a non-Python host binds - torch.distributed only ships the 128-byte rendezvous id.  No-op (False)
    single-process, on the gloo backend (CPU tests), or with FADTK_NATIVE_ALLREDUCE=0."""
    global _native_engine
    if not is_distributed() or td.get_backend() != "nccl" or os.environ.get("FADTK_NATIVE_ALLREDUCE", "1") == "0":
        return False
    if _native_engine is engine and engine.has_comm:
        return True
    uid = broadcast_object(engine.comm_unique_id() if rank() == 0 else None)
    engine.comm_init(uid, rank(), world_size())
    _native_engine = engine
    return True


def allreduce_sum_(t: torch.Tensor) -> torch.Tensor:
    if is_distributed():
        eng = _native_engine
        if (eng is not None and eng.has_comm and t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()
                and t.device.index == eng.device):
            eng.allreduce_sum_(t)                             # ncclAllReduce on the current stream, through the C ABI
        else:
            td.all_reduce(t, op=td.ReduceOp.SUM)
    return t


def allgather_objects(obj):
    if not is_distributed():
        return [obj]
    out = [None] * world_size()
    td.all_gather_object(out, obj)
    return out


def broadcast_object(obj, src: int = 0):
    """a picklable object from rank ``src`` to every rank (the list of files still to embed)"""
    if not is_distributed():
        return obj
    box = [obj if rank() == src else None]
    td.broadcast_object_list(box, src=src)
    return box[0]


def broadcast_int64(arr, src: int = 0):
    """numpy int64 array from rank ``src`` to every rank (FAD-inf bootstrap indices: one RNG stream)."""
    if not is_distributed():
        return arr
    dev = torch.device("cuda", torch.cuda.current_device()) if td.get_backend() == "nccl" else torch.device("cpu")
    t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.int64)).to(dev)
    td.broadcast(t, src=src)
    return t.cpu().numpy()


def barrier():
    if is_distributed():
        td.barrier()


def max_over_ranks(x: float) -> float:
    if not is_distributed():
        return x
    dev = torch.device("cuda", torch.cuda.current_device()) if td.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    td.all_reduce(t, op=td.ReduceOp.MAX)
    return float(t.item())
