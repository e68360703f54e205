"""Vortex-force term, API of core/vortex_force.py:13-81.

The reference makes six Fortran line sweeps (vortex_force_direc / vortex_force_flip for the
index triplets "ikj", "jik", "kji", fortran_vortex_force.f90:10-165) on transposed copies of
the fields; here one kernel evaluates the six WENO fluxes per cell in the canonical layout and
adds them to rhs.u in the same order.
"""
from . import lib
from .timing import timing


@timing
def vortex_force(state, rhs, order, linear=None):
    """rhs.u += vortex force.  `order` is ignored unless the linear branch is on, like the shipped Fortran
    (linear = .false., fortran_vortex_force.f90:28); linear=None follows nyles_b200.LINEAR_UPWIND."""
    assert order in {1, 2, 3, 4, 5}
    U, w, du = state.U, state.vor, rhs.u
    t = U["i"].tensor
    if linear is None:
        import nyles_b200
        linear = nyles_b200.LINEAR_UPWIND
    if linear:
        lib.check(lib.load().ny_vortex_force_linear(
            lib.context(t.device), lib.ptr(U["i"].tensor), lib.ptr(U["j"].tensor), lib.ptr(U["k"].tensor),
            lib.ptr(w["i"].tensor), lib.ptr(w["j"].tensor), lib.ptr(w["k"].tensor),
            lib.ptr(du["i"].tensor), lib.ptr(du["j"].tensor), lib.ptr(du["k"].tensor),
            int(order), lib.ext(t), lib.stream()))
        return
    lib.check(lib.load().ny_vortex_force(
        lib.context(t.device), lib.ptr(U["i"].tensor), lib.ptr(U["j"].tensor), lib.ptr(U["k"].tensor),
        lib.ptr(w["i"].tensor), lib.ptr(w["j"].tensor), lib.ptr(w["k"].tensor),
        lib.ptr(du["i"].tensor), lib.ptr(du["j"].tensor), lib.ptr(du["k"].tensor),
        lib.ext(t), lib.stream()))
