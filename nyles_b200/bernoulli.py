"""Bernoulli term without pressure, API of core/bernoulli.py:14-32 (fortran_bernoulli.f90:2-58)."""
from . import lib
from .timing import timing


def _bern(state, rhs, grid, euler):
    du = rhs.u
    t = state.ke.tensor
    lib.check(lib.load().ny_bernoulli(
        lib.context(t.device), lib.ptr(state.ke.tensor), lib.ptr(None if euler else state.b.tensor),
        lib.ptr(du["i"].tensor), lib.ptr(du["j"].tensor), lib.ptr(du["k"].tensor),
        grid.dz, 1 if euler else 0, lib.ext(t), lib.stream()))


@timing
def bernoulli(state, rhs, grid):
    """Add b*grad(z) - grad(ke) to rhs.u."""
    _bern(state, rhs, grid, False)


@timing
def bernoulli_euler(state, rhs, grid):
    """Add -grad(ke) to rhs.u (Euler3d model, no buoyancy)."""
    _bern(state, rhs, grid, True)
