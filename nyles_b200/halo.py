"""Halo filling, API of core/mpi/halo.py (``set_halo(param, state)`` -> ``.fill(Scalar|Vector|tensor)``).

The reference exchanges 26 boxes per array with persistent MPI requests.  Here the domain is cut
into slabs along z only (one GPU per slab), so a fill is:
  1. the two z faces (nh contiguous planes each) exchanged with the slab neighbours through
     torch.distributed (NCCL over NVLink; gloo in the CPU tests) -- or wrapped locally when the
     z neighbour is this very rank (one slab, z-periodic);
  2. the periodic x / y directions wrapped locally over ALL planes, halo planes included, which
     reproduces the edge and corner boxes of the 26-neighbour exchange.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import lib
from . import mpitools


class Halo(object):
    def __init__(self, grid):
        self.neighbours = grid["neighbours"]
        self.nh = grid["nh"]
        self.size = tuple(grid["size"])
        self.domainindices = grid["domainindices"]
        self.myrank = mpitools.get_myrank()
        ng = self.neighbours
        self.below = ng.get((-1, 0, 0))
        self.above = ng.get((1, 0, 0))
        self.yper = (0, -1, 0) in ng
        self.xper = (0, 0, -1) in ng
        for key, rank in ng.items():
            if key[0] == 0 and rank != self.myrank:
                raise NotImplementedError("nyles_b200 decomposes along z only (npx = npy = 1)")
        self.z_self = self.below is not None and self.below == self.myrank and self.above == self.myrank
        self.z_remote = (self.below is not None and self.below != self.myrank) or \
                        (self.above is not None and self.above != self.myrank)

    # ------------------------------------------------------------------ public
    def fill(self, thing, local_only=False):
        """local_only: wrap the periodic directions of this rank only, no exchange with the slab neighbours
        (for fields whose halo planes were computed locally from exchanged inputs)."""
        nature = type(thing).__name__
        if nature == "Scalar":
            self.fillarrays([thing.tensor], local_only)
        elif nature == "Vector":
            self.fillarrays([thing[d].tensor for d in "ijk"], local_only)
        elif nature == "FieldView":
            self.fillarrays([thing.tensor], local_only)
        elif isinstance(thing, torch.Tensor):
            self.fillarrays([thing], local_only)
        else:
            raise ValueError("try to fill halo with unidentified object")

    def fillvector(self, vector):
        self.fillarrays([vector[d].tensor for d in "ijk"])

    def fillarray(self, x):
        self.fillarrays([x])

    def fillarrays(self, xs, local_only=False):
        if local_only and self.z_remote:
            per = (0, 1 if self.yper else 0, 1 if self.xper else 0)
            if any(per):
                for x in xs:
                    self._wrap(x, per)
            return
        if self.z_remote and xs[0].is_cuda:
            self._exchange_lib(xs)              # z faces through NCCL + local x/y wrap, one library call
            return
        if self.z_remote:
            self._exchange_z(xs)                # host tensors (gloo): exercised by the CPU tests only
        per = (1 if self.z_self else 0, 1 if self.yper else 0, 1 if self.xper else 0)
        if any(per):
            for x in xs:
                self._wrap(x, per)

    # ------------------------------------------------------------------ pieces
    def _wrap(self, x, per):
        if not x.is_cuda:
            return _wrap_host(x, self.nh, per)
        arr = (C.c_int * 3)(*per)
        lib.check(lib.load().ny_halo_fill_self(lib.context(x.device), lib.ptr(x), lib.ext(x), self.nh,
                                               C.byref(arr), lib.stream()))

    def _exchange_lib(self, xs):
        from . import comm
        x0 = xs[0]
        ptrs = (C.c_void_p * len(xs))(*[lib.ptr(x).value for x in xs])
        below = -1 if self.below is None else self.below
        above = -1 if self.above is None else self.above
        lib.check(lib.load().ny_halo_exchange(lib.context(x0.device), comm.get(), ptrs, len(xs), lib.ext(x0), self.nh,
                                              below, above, 1 if self.yper else 0, 1 if self.xper else 0,
                                              lib.stream()))

    def _exchange_z(self, xs):
        nh = self.nh
        ops = []
        # sends first, then receives in the opposite neighbour order: with two ranks and periodic z
        # both neighbours are the same peer and messages are matched in posting order
        for x in xs:
            n = x.shape[0]
            lo = nh if self.below is not None else 0
            hi = n - nh if self.above is not None else n
            if self.below is not None:
                ops.append(dist.P2POp(dist.isend, x[lo:lo + nh].contiguous(), self.below))
            if self.above is not None:
                ops.append(dist.P2POp(dist.isend, x[hi - nh:hi].contiguous(), self.above))
        recvs = []
        for x in xs:
            n = x.shape[0]
            if self.above is not None:
                buf = torch.empty_like(x[n - nh:n]); recvs.append((x, slice(n - nh, n), buf))
                ops.append(dist.P2POp(dist.irecv, buf, self.above))
            if self.below is not None:
                buf = torch.empty_like(x[0:nh]); recvs.append((x, slice(0, nh), buf))
                ops.append(dist.P2POp(dist.irecv, buf, self.below))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        for x, sl, buf in recvs:
            x[sl] = buf


def _wrap_host(x, nh, per):
    """Periodic wrap of a host tensor (CPU tests of the slab logic): z, then y, then x, over all cells."""
    for ax in range(3):
        if per[ax]:
            n = x.shape[ax] - 2 * nh
            lo = [slice(None)] * 3; hi = [slice(None)] * 3; slo = [slice(None)] * 3; shi = [slice(None)] * 3
            lo[ax] = slice(0, nh); slo[ax] = slice(n, n + nh)
            hi[ax] = slice(nh + n, 2 * nh + n); shi[ax] = slice(nh, 2 * nh)
            x[tuple(lo)] = x[tuple(slo)].clone()
            x[tuple(hi)] = x[tuple(shi)].clone()


def check_halo_width(procs, shape, nh):
    """A sub-domain must be at least as wide as the halo it feeds (core/mpi/halo.py)."""
    for p, s in zip(procs, shape):
        if p > 1 and s < nh:
            raise ValueError("subdomain narrower than the halo")


def set_halo(param, state):
    b = state.b
    localgrid = {"shape": b.shape, "size": tuple(b.tensor.shape), "nh": param["nh"],
                 "neighbours": param["neighbours"], "domainindices": b.domainindices, "extension": 6}
    check_halo_width(param["procs"], b.shape, param["nh"])
    return Halo(localgrid)
