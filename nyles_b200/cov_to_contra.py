"""Covariant -> contravariant velocity, API of core/cov_to_contra.py:4-20 (Cartesian metric)."""
from . import lib
from .timing import timing


@timing
def U_from_u(state, grid):
    u, U = state.u, state.U
    t = u["i"].tensor
    lib.u_epoch += 1
    lib.check(lib.load().ny_U_from_u(
        lib.context(t.device), lib.ptr(u["i"].tensor), lib.ptr(u["j"].tensor), lib.ptr(u["k"].tensor),
        lib.ptr(U["i"].tensor), lib.ptr(U["j"].tensor), lib.ptr(U["k"].tensor),
        grid.idx2, grid.idy2, grid.idz2, lib.ext(t), lib.stream()))
