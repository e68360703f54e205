"""-div(tracer * U) with WENO upwinding, API of core/tracer.py:8-77 (fortran_upwind.f90:3-87)."""
from . import lib
from .timing import timing


class Tracer_numerics(object):
    def __init__(self, param, grid, traclist, order, diff_coef=[], linear=None):
        import nyles_b200
        # the linear branch of fortran_upwind.f90:33-64 instead of WENO (dormant in the shipped reference)
        self.linear = nyles_b200.LINEAR_UPWIND if linear is None else linear
        self.traclist = traclist
        assert order in {1, 2, 3, 4, 5}
        self.order = order
        self.diffusion = len(diff_coef) > 0
        if self.diffusion:
            self.ids2 = grid.ids2
            self.diff_coef = diff_coef
        ngbs = param["neighbours"]
        keys = {"i": ((0, 0, -1), (0, 0, 1)), "j": ((0, -1, 0), (0, 1, 0)), "k": ((-1, 0, 0), (1, 0, 0))}
        self.i0 = {d: 0 if keys[d][0] in ngbs else 1 for d in "ijk"}
        self.i1 = {d: 0 if keys[d][1] in ngbs else 1 for d in "ijk"}

    @timing
    def rhstrac(self, state, rhs, last=False):
        """rhs.<tracer> = -div(U tracer) (+ diffusion when `last` and a coefficient is set)."""
        L = lib.load()
        U = state.U
        for tracname in self.traclist:
            trac, dtrac = state.get(tracname).tensor, rhs.get(tracname).tensor
            args = (lib.context(trac.device), lib.ptr(trac), lib.ptr(U["i"].tensor), lib.ptr(U["j"].tensor),
                    lib.ptr(U["k"].tensor), lib.ptr(dtrac))
            if self.linear:
                lib.check(L.ny_upwind_linear(*args, int(self.order), lib.ext(trac), lib.stream()))
                if self.diffusion and last and tracname in self.diff_coef:
                    c = [self.diff_coef[tracname] * self.ids2[d] for d in "ijk"]
                    raise NotImplementedError("tracer diffusion with the linear upwind branch (the reference interleaves "
                                              "the Laplacian between the three sweeps, tracer.py:72-77; coefficients %r)" % c)
            elif self.diffusion and last and tracname in self.diff_coef:
                c = [self.diff_coef[tracname] * self.ids2[d] for d in "ijk"]
                lib.check(L.ny_upwind_diff(*args, c[0], c[1], c[2], lib.ext(trac), lib.stream()))
            else:
                lib.check(L.ny_upwind(*args, lib.ext(trac), lib.stream()))
