"""Laplacian viscosity on the covariant velocity, API of core/viscosity.py:3-9
(fortran_dissipation.f90:2-35 along i, j, k for each component)."""
from . import lib


def add_viscosity(grid, state, dstate, viscosity):
    cx, cy, cz = (viscosity * grid.ids2[d] for d in "ijk")
    for comp in "ijk":
        t = state.u[comp].tensor
        lib.check(lib.load().ny_add_laplacian(lib.context(t.device), lib.ptr(t), lib.ptr(dstate.u[comp].tensor),
                                              cx, cy, cz, lib.ext(t), lib.stream()))
