"""Sub-domain partition helpers, API of core/mpi/topology.py.

`topology` is the module-level geometry string exactly as in the reference
(topology.py:24: set once with ``topo.topology = 'closed'``).
"""
import itertools

import numpy as np

topology = "undefined"

POSSIBLE = ["closed", "perio_x", "perio_xy", "perio_y", "perio_xyz"]


def rank2loc(rank, procs):
    """(k, j, i) location of `rank` in the process grid procs = [npz, npy, npx] (topology.py:27-40)."""
    return [rank // (procs[2] * procs[1]), (rank // procs[2]) % procs[1], rank % procs[2]]


def loc2rank(loc, procs):
    return (loc[0] * procs[1] + loc[1]) * procs[2] + loc[2]


def get_neighbours(location, procs, incr=(1, 1, 1), extension=26, topo=None):
    """Dictionary {(dk,dj,di): rank} of existing neighbours (topology.py:72-157)."""
    geom = topology if topo is None else topo
    assert geom in POSSIBLE, "you forgot to set the topology"
    if isinstance(location, (int, np.integer)):
        k, j, i = rank2loc(int(location), procs)
    elif isinstance(location, (list, tuple)):
        k, j, i = location
    else:
        raise ValueError
    nz, ny, nx = procs
    incz, incy, incx = incr
    wanted = {6: (1,), 18: (1, 2), 26: (1, 2, 3)}
    if extension not in wanted:
        raise ValueError("neighbours should be 6, 18 or 26")
    ngs = {}
    for d in itertools.product([-1, 0, 1], repeat=3):
        if sum(abs(c) for c in d) not in wanted[extension]:
            continue
        dk, dj, di = d
        kk, jj, ii = k + dk * incz, j + dj * incy, i + di * incx
        if "x" not in geom and not 0 <= ii < nx:
            continue
        if "y" not in geom and not 0 <= jj < ny:
            continue
        if "z" not in geom and not 0 <= kk < nz:
            continue
        ngs[d] = loc2rank([kk % nz, jj % ny, ii % nx], procs)
    return ngs


def get_variable_shape(innersize, ngbs, nh):
    """Array extents with a halo only on sides that have a neighbour (topology.py:253-304).

    Returns (size, (k0, k1, j0, j1, i0, i1))."""
    size = list(innersize)
    lo, hi = [0, 0, 0], [0, 0, 0]
    minus = [(-1, 0, 0), (0, -1, 0), (0, 0, -1)]
    plus = [(1, 0, 0), (0, 1, 0), (0, 0, 1)]
    for ax in range(3):
        if minus[ax] in ngbs:
            size[ax] += nh
            lo[ax] = nh
        hi[ax] = size[ax]
        if plus[ax] in ngbs:
            size[ax] += nh
    return size, (lo[0], hi[0], lo[1], hi[1], lo[2], hi[2])


def noneighbours():
    return {}
