"""Projection enforcing div U = 0, API of core/projection.py:16-87."""
from . import lib
from .timing import timing


@timing
def compute_div(state, **kwargs):
    """div = delta_i[U] + delta_j[V] + delta_k[W] (left differences, fortran_bernoulli.f90:61-97)."""
    U = state.U
    t = state.div.tensor
    lib.check(lib.load().ny_div(lib.context(t.device), lib.ptr(U["i"].tensor), lib.ptr(U["j"].tensor),
                                lib.ptr(U["k"].tensor), lib.ptr(t), lib.ext(t), lib.stream()))


@timing
def compute_p(mg, state, grid, ngbs):
    """Solve the Poisson equation for p from div(U) and correct u with -delta[p]."""
    compute_div(state)
    mg.solve_directly(state.p.view("i"), state.div.view("i"))
    u = state.u
    t = state.p.tensor
    lib.check(lib.load().ny_gradp(lib.context(t.device), lib.ptr(t), lib.ptr(u["i"].tensor),
                                  lib.ptr(u["j"].tensor), lib.ptr(u["k"].tensor), lib.ext(t), lib.stream()))
