"""Multigrid pressure solver, API of core/mgfordriver.py:5-129.

The reference wraps the Fortran library libmgmod64.so (ctypes) and moves the right-hand side
and the solution through two full host copies per solve (mgfordriver.py:72-78).  Here the
hierarchy lives in GPU memory inside libnyles_b200.so; embedding div into the padded multigrid
array and extracting p*dx^2 are device kernels.
"""
import ctypes as C

import numpy as np
import torch

from . import lib


def _tensor(a):
    return a.tensor if hasattr(a, "tensor") else a


class MG(object):
    IVAR = dict(x=1, b=2, r=3, y=4, diag=5, idiag=6, msk=7, Rcoef=8, Pcoef=9)

    def __init__(self, npx, npy, nx, ny, nz, nh, topology=1, device=None, npz=1):
        """nx, ny, nz: LOCAL interior extents of this rank (as in the reference); npz slabs along z."""
        if npx != 1 or npy != 1:
            raise NotImplementedError("nyles_b200 decomposes along z only (npx = npy = 1)")
        if nh != 3:
            raise ValueError("the multigrid works with nh = 3 (MG_Param, core/mgfor/mg_types.f90:19)")
        self.nh = nh
        self.L = lib.load()
        self.ctx = lib.context(device)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        h = C.c_void_p()
        if npz > 1:
            from . import comm
            if not comm.active() or self.L.ny_comm_size(comm.get()) != npz:
                raise lib.NylesB200Error("npz = %d needs %d ranks (torchrun --nproc-per-node %d)" % (npz, npz, npz))
            lib.check(self.L.ny_mg_create_slab(self.ctx, comm.get(), nx, ny, nz * npz, topology, C.byref(h)))
        else:
            lib.check(self.L.ny_mg_create(self.ctx, nx, ny, nz, topology, C.byref(h)))
        self.mg = h
        self.nlevels = self.L.ny_mg_nlevels(self.mg)
        self.shape = self.get_arrayshape()
        self.stats = {"normb": 0, "res": [0], "blowup": False, "nite": 0}
        self.nvcycles = 0                         # running total, for the bytes model of bench.py
        self._stats = lib.ny_mg_stats()

    def __del__(self):
        try:
            if getattr(self, "mg", None):
                self.L.ny_mg_destroy(self.mg)
                self.mg = None
        except Exception:
            pass

    # ---- shapes / indices
    def get_arrayshape(self, lev=1):
        s = (C.c_int * 3)()
        lib.check(self.L.ny_mg_shape(self.mg, lev, C.byref(s)))
        return tuple(s)

    def get_idx_from_neighbours(self, neighbours):
        nh = self.nh
        i0 = 0 if (0, 0, -1) in neighbours else nh
        i1 = None if (0, 0, 1) in neighbours else -nh
        j0 = 0 if (0, -1, 0) in neighbours else nh
        j1 = None if (0, 1, 0) in neighbours else -nh
        k0 = 0 if (-1, 0, 0) in neighbours else nh
        k1 = None if (1, 0, 0) in neighbours else -nh
        return (slice(k0, k1), slice(j0, j1), slice(i0, i1))

    def preallocate_for_nyles(self, dx, neighbours, halo):
        self.dx = dx
        self.idx = self.get_idx_from_neighbours(neighbours)
        self.lo = (C.c_int * 3)(*[s.start for s in self.idx])
        self.halo = halo

    def set_param(self, maxite=20, tol=1e-6, omega=0.9):
        lib.check(self.L.ny_mg_set_param(self.mg, maxite, tol, omega))

    # ---- solves
    def _record(self):
        s = self._stats
        self.stats = {"normb": s.normb, "res": [s.reshist[i] for i in range(s.nres)], "blowup": False,
                      "nite": s.nite, "final_res": s.res}
        self.nvcycles += s.nite

    def solve_directly(self, p, div):
        """b_mg[idx] = div ; solve ; p = x_mg[idx]*dx**2  (mgfordriver.py:66-78)."""
        p, div = _tensor(p), _tensor(div)
        self.halo.fill(div)
        lib.check(self.L.ny_mg_solve_directly(self.mg, lib.ptr(p), lib.ptr(div), lib.ext(div), C.byref(self.lo),
                                              self.dx ** 2, C.byref(self._stats), lib.stream()))
        self._record()

    def project(self, state, grid):
        """projection.compute_p fused around the in-place solve (ny_mg_project): div from u, solve,
        p = x*dx**2, u -= delta p."""
        u = state.u
        t = state.p.tensor
        lib.check(self.L.ny_mg_project(self.mg, lib.ptr(u["i"].tensor), lib.ptr(u["j"].tensor), lib.ptr(u["k"].tensor),
                                       lib.ptr(state.div.tensor), lib.ptr(t), grid.idx2, grid.idy2, grid.idz2,
                                       lib.ext(t), C.byref(self.lo), self.dx ** 2, C.byref(self._stats), lib.stream()))
        self._record()

    def solve(self, x, b, fill_halo=False):
        """Full-array interface: b and x have the padded multigrid shape (mgfordriver.py:101-106).
        fill_halo: fill the periodic / slab halos of b (and of the warm-start x) first, as
        solve_directly's caller does; the solver then knows they are consistent."""
        self.set_array(b, ivar=2)
        if fill_halo:
            self.op("fill", 1)
        lib.check(self.L.ny_mg_solve(self.mg, C.byref(self._stats), lib.stream()))
        self._record()
        self.get_array(x, ivar=1)

    # ---- whole-level array access (set_pyarray / get_pyarray)
    def _as_device(self, array, lev):
        if isinstance(array, np.ndarray):
            array = torch.as_tensor(np.ascontiguousarray(array), dtype=torch.float64)
        t = _tensor(array).to(self.device)
        assert tuple(t.shape) == self.get_arrayshape(lev), "array must have the padded multigrid shape"
        return t.contiguous()

    def set_array(self, array, ivar=1, lev=1):
        t = self._as_device(array, lev)
        lib.check(self.L.ny_mg_set_array(self.mg, lev, ivar, lib.ptr(t), lib.stream()))
        torch.cuda.current_stream().synchronize()      # t may be a temporary

    def get_array(self, array=None, ivar=1, lev=1):
        out = torch.empty(self.get_arrayshape(lev), dtype=torch.float64, device=self.device)
        lib.check(self.L.ny_mg_get_array(self.mg, lev, ivar, lib.ptr(out), lib.stream()))
        if array is None:
            return out
        if isinstance(array, np.ndarray):
            array[...] = out.cpu().numpy()
        else:
            _tensor(array).copy_(out)
        return array

    def set_mask(self, msk):
        """Obstacles: msk (padded multigrid shape of level 1, 1 = fluid, 0 = solid) replaces the default mask;
        the coarse masks and the coefficients of every level are rebuilt by the reference's own setup
        (mgfor/tests.f90:207-212: write oper(1)%msk, setup_fine_msk, setup_operators)."""
        self.set_array(msk, ivar=7)
        lib.check(self.L.ny_mg_setup_operators(self.mg, lib.stream()))

    def is_box(self):
        return bool(self.L.ny_mg_is_box(self.mg))

    def set_fast_path(self, on):
        """on=False: generic kernels reading the coefficient arrays; on=True: fused box kernels."""
        lib.check(self.L.ny_mg_set_fast_path(self.mg, 1 if on else 0))

    def set_fused_legs(self, on):
        """on=False: V-cycles use one box kernel per operator instead of the fused legs; on=2: fused legs
        with every tile through the general kernel instance (no wall-free specialisation)."""
        lib.check(self.L.ny_mg_set_fused_legs(self.mg, int(on)))

    def op(self, name, lev=1):
        code = dict(smooth=1, residual=2, restriction=3, prolongation=4, vcycle=5, fill=6)[name]
        lib.check(self.L.ny_mg_op(self.mg, code, lev, lib.stream()))
