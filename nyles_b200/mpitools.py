"""Rank bookkeeping and scalar reductions, API of core/mpi/mpitools.py (mpi4py -> torch.distributed).

One process per GPU, launched with torchrun; without a process group everything is rank 0 of 1.
"""
import numpy as np
import torch
import torch.distributed as dist


def _active():
    return dist.is_available() and dist.is_initialized()


def get_size():
    return dist.get_world_size() if _active() else 1


def get_myrank(procs=None):
    if procs is not None:
        assert get_size() == int(np.prod(procs)), \
            "launch with torchrun --nproc-per-node %i (found world size %i)" % (np.prod(procs), get_size())
    return dist.get_rank() if _active() else 0


def barrier():
    if _active():
        dist.barrier()


def abort():
    raise SystemExit(1)


def _reduce(value, op, device):
    if not _active():
        return value
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=op)
    return t.item()


def _dev():
    if _active() and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def global_sum(localsum):
    return _reduce(localsum, dist.ReduceOp.SUM, _dev()) if _active() else localsum


def global_max(localmax):
    return _reduce(localmax, dist.ReduceOp.MAX, _dev()) if _active() else localmax
