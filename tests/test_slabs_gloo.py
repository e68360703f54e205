"""CPU, two processes over gloo: the host side of the z-slab decomposition.

What runs here is the same Python that runs on the GPUs -- `topology`, `variables` (with
param['device']='cpu'), `grid`, `halo.Halo` (its torch.distributed face exchange) and
`mpitools` -- with host tensors in place of device tensors.  Each slab is compared with the
matching cut of ONE undecomposed domain filled by the oracle's single-process halo
(oracle/model.py: core/mpi/halo.py:93-178), so a fill over two ranks must reproduce the
26-neighbour exchange of the reference including edge and corner boxes.
The device legs of the same path (ny_halo_exchange, the slab multigrid) are covered on two
B200s by tests/test_gpu_slabs.py.
"""
import os
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import model as M

WORLD = 2
NH = 3
GLOBAL = dict(nx=8, ny=6, nz=16)


def _global_reference(geometry, seed):
    """Undecomposed domain, four random fields, halos filled by the oracle."""
    p = M.make_param(geometry=geometry, **GLOBAL)
    st = M.get_state(p)
    halo = M.Halo(p, st.b)
    rng = np.random.default_rng(seed)
    fields = []
    for s in (st.b, st.u["i"], st.u["j"], st.u["k"]):
        a = s.view("i")
        a[...] = rng.standard_normal(a.shape)
        fields.append(a)
    ref_unfilled = [a.copy() for a in fields]
    for a in fields:
        halo.fillarray(a)
    return p, st.b.domainindices, ref_unfilled, fields


def _worker(rank, initfile, geometry, q):
    try:
        dist.init_process_group("gloo", init_method="file://" + initfile, rank=rank, world_size=WORLD)
        from nyles_b200 import grid as G, halo as H, mpitools, topology as topo, variables as V

        procs = [WORLD, 1, 1]
        topo.topology = geometry
        assert mpitools.get_size() == WORLD and mpitools.get_myrank(procs) == rank
        loc = topo.rank2loc(rank, procs)
        ngbs = topo.get_neighbours(loc, procs)
        nzl = GLOBAL["nz"] // WORLD
        param = dict(nx=GLOBAL["nx"], ny=GLOBAL["ny"], nz=nzl, nh=NH, neighbours=ngbs, procs=procs, loc=loc,
                     npx=1, npy=1, npz=WORLD, Lx=1.0, Ly=1.0, Lz=2.0, device="cpu")
        state = V.get_state(param)
        halo = H.set_halo(param, state)
        grid = G.Grid(param)

        gp, gdomi, unfilled, filled = _global_reference(geometry, seed=7)
        gk0 = gdomi[0]                       # first interior plane of the global array
        k0, k1 = state.b.domainindices[:2]
        zlo = gk0 + rank * nzl - k0          # global plane stored in local plane 0
        nloc = state.b.tensor.shape[0]

        def global_planes():
            idx = np.arange(zlo, zlo + nloc)
            if "z" in geometry:              # periodic: the wrap planes of the edge slabs
                n = GLOBAL["nz"]
                idx = (idx - gk0) % n + gk0
            return idx

        planes = global_planes()
        mine = [state.b, state.u["i"], state.u["j"], state.u["k"]]
        for s, a in zip(mine, unfilled):
            cut = a[planes].copy()
            # start from garbage halos: interior from the global field, halo planes poisoned
            t = torch.full(tuple(s.tensor.shape), float("nan"), dtype=torch.float64)
            j0, j1, i0, i1 = s.domainindices[2:]
            t[k0:k1, j0:j1, i0:i1] = torch.from_numpy(cut[k0:k1, j0:j1, i0:i1])
            s.tensor.copy_(t)

        halo.fill(state.b)                   # Scalar
        halo.fill(state.u)                   # Vector: three arrays in one exchange
        for name, s, ref in zip("b u_i u_j u_k".split(), mine, filled):
            got = s.tensor.numpy()
            want = ref[planes]
            assert np.array_equal(got, want), "rank %d %s: slab fill differs from the undecomposed fill" % (rank, name)

        # coordinates continue across the slab interface (grid.py:84-135)
        zg = (np.arange(GLOBAL["nz"] + 2 * gk0) + 0.5 - gk0) * grid.dz
        assert np.allclose(grid.z_b_1D, zg[zlo:zlo + nloc], rtol=0, atol=1e-14)
        assert grid.dz == 2.0 / GLOBAL["nz"]

        # scalar reductions of nyles.compute_dt / the blow-up test (mpitools.py:40-53)
        assert mpitools.global_max(float(rank + 1)) == float(WORLD)
        assert mpitools.global_sum(float(rank + 1)) == float(sum(range(1, WORLD + 1)))
        loc_max = float(np.max(np.abs(unfilled[1][planes][k0:k1])))
        assert mpitools.global_max(loc_max) == float(np.max(np.abs(unfilled[1][gk0:gk0 + GLOBAL["nz"]])))
        mpitools.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except BaseException as e:               # noqa: BLE001 -- report to the parent, then die
        import traceback
        q.put((rank, traceback.format_exc()))
        raise e


@pytest.mark.parametrize("geometry", ["closed", "perio_xy", "perio_xyz"])
def test_two_slab_halo_fill_matches_single_domain(geometry):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    with tempfile.TemporaryDirectory() as d:
        initfile = os.path.join(d, "rdzv")
        procs = [ctx.Process(target=_worker, args=(r, initfile, geometry, q)) for r in range(WORLD)]
        for p in procs:
            p.start()
        results = {}
        for _ in range(WORLD):
            r, msg = q.get(timeout=240)
            results[r] = msg
        for p in procs:
            p.join(timeout=60)
    for r in range(WORLD):
        assert results.get(r) == "ok", "rank %d:\n%s" % (r, results.get(r))


def test_slab_shapes_and_neighbours():
    """procs = [npz,1,1]: who is above/below, which sides carry halos (topology.py:72-157,253-304)."""
    from nyles_b200 import topology as topo
    procs = [4, 1, 1]
    n = topo.get_neighbours(topo.rank2loc(0, procs), procs, topo="closed")
    assert n == {(1, 0, 0): 1}
    n = topo.get_neighbours(topo.rank2loc(2, procs), procs, topo="closed")
    assert n[(-1, 0, 0)] == 1 and n[(1, 0, 0)] == 3 and len(n) == 2
    n = topo.get_neighbours(topo.rank2loc(3, procs), procs, topo="perio_xyz")
    assert n[(1, 0, 0)] == 0 and n[(-1, 0, 0)] == 2 and n[(0, 0, 1)] == 3 and len(n) == 26
    n = topo.get_neighbours(topo.rank2loc(0, procs), procs, topo="perio_xy")
    assert (-1, 0, 0) not in n and n[(1, 1, 1)] == 1 and n[(0, -1, 0)] == 0
    size, domi = topo.get_variable_shape([8, 6, 4], n, NH)
    assert size == [8 + NH, 6 + 2 * NH, 4 + 2 * NH] and domi == (0, 8, NH, NH + 6, NH, NH + 4)
