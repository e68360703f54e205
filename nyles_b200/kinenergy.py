"""Kinetic energy at cell centres, API of core/kinenergy.py:7-24 (fortran_kinenergy.f90:3-56)."""
from . import lib
from .timing import timing


@timing
def kinenergy(state, grid, order=2):
    u = state.u
    t = u["i"].tensor
    lib.check(lib.load().ny_kin(
        lib.context(t.device), lib.ptr(u["i"].tensor), lib.ptr(u["j"].tensor), lib.ptr(u["k"].tensor),
        lib.ptr(state.ke.tensor), grid.ids2["i"], grid.ids2["j"], grid.ids2["k"], lib.ext(t), lib.stream()))
