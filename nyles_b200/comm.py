"""The slab communicator of libnyles_b200.so (ny_comm: NCCL over NVLink, one rank per GPU).

Plays the role of MPI.COMM_WORLD in the reference (core/mpi/mpitools.py, mgfor's mpi_f08 calls).
The process group of torch.distributed is only used to bootstrap it: rank 0 asks the library
for a NCCL unique id, the 128 bytes are broadcast, every rank joins.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import lib

_comm = None


def active():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def get():
    """ny_comm* of this process, or a null pointer when there is a single rank."""
    global _comm
    if not active():
        return C.c_void_p(0)
    if _comm is None:
        L = lib.load()
        buf = C.create_string_buffer(128)
        if dist.get_rank() == 0:
            lib.check(L.ny_comm_unique_id(buf))
        box = [buf.raw]
        dist.broadcast_object_list(box, src=0)
        h = C.c_void_p()
        lib.check(L.ny_comm_init(lib.context(), dist.get_world_size(), dist.get_rank(), box[0], C.byref(h)))
        torch.cuda.synchronize()
        _comm = h
    return _comm


def free():
    global _comm
    if _comm is not None:
        lib.load().ny_comm_free(_comm)
        _comm = None
