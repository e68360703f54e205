"""Covariant vorticity, API of core/vorticity.py:7-34 (fortran_vorticity.f90:2-28).

omega_k = delta_i[u_j] - delta_j[u_i] for the three components in one kernel, with the
reference's closures at the array ends and f*dx*dy added to omega_z when fparameter > 0.
"""
from . import lib
from .timing import timing


@timing
def vorticity(state, fparameter):
    u, w = state.u, state.vor
    t = u["i"].tensor
    lib.check(lib.load().ny_vorticity(
        lib.context(t.device), lib.ptr(u["i"].tensor), lib.ptr(u["j"].tensor), lib.ptr(u["k"].tensor),
        lib.ptr(w["i"].tensor), lib.ptr(w["j"].tensor), lib.ptr(w["k"].tensor),
        lib.ext(t), float(fparameter), lib.stream()))


vorticity_all_comp = vorticity
