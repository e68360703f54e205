"""Decomposition self-check: z slabs on N GPUs against the same problem on ONE GPU.

Every stencil of the path evaluates a cell from exact halo values, so the slab run must reproduce the
single-GPU run bit for bit -- every field of every slab, every dt, every V-cycle count (the two multigrid
norms are summed in another order; they only feed the stop test).  bench.py runs this before its timed
region whenever it is launched on more than one GPU and reports the outcome as "parity_check", so that
the scaling record carries N-rank correctness (core/mpi/halo.py:123-188, core/mgfor/mod_halo.f90:200-262,
mod_gluesplit.f90:142-207 are the reference routines whose replacement this exercises).

Usage (one process per GPU, torchrun):
    ref = selfcheck.single_gpu_runs(world)        # BEFORE dist.init_process_group: whole problems, this GPU
    dist.init_process_group("nccl", ...)
    ok, report = selfcheck.slab_runs(ref)         # the same problems on `world` slabs
"""
import numpy as np
import torch
import torch.distributed as dist

FIELDS = ["b", "u_i", "u_j", "u_k", "p", "ke", "vor_i", "vor_j", "vor_k", "div"]


def cases(world, nzl=64):
    n = (64, 64, nzl * world)
    L = 2 * np.pi
    return [
        dict(name="closed LES 64x64x%d" % n[2], modelname="LES", geometry="closed", n=n,
             L=(4.0, 4.0, 4.0 * world * nzl / 64), dt_max=0.05, steps=3),
        dict(name="perio_xyz Euler3d 64x64x%d" % n[2], modelname="Euler3d", geometry="perio_xyz", n=n,
             L=(L, L, L * world * nzl / 64), dt_max=0.02, steps=3),
        dict(name="perio_xy rotating LES 64x64x%d" % n[2], modelname="LES", geometry="perio_xy", n=n,
             L=(4.0, 4.0, 4.0 * world * nzl / 64), dt_max=0.05, steps=3, rotating=True),
    ]


def _make(kw, npz):
    from . import nyles, parameters
    parameters.InextensibleDict.unfreeze()
    up = parameters.UserParameters()
    up.model["modelname"] = kw["modelname"]
    up.model["geometry"] = kw["geometry"]
    up.model["Lx"], up.model["Ly"], up.model["Lz"] = kw["L"]
    up.discretization["global_nx"], up.discretization["global_ny"], up.discretization["global_nz"] = kw["n"]
    up.MPI["npz"] = npz
    up.time["cfl"], up.time["dt_max"] = 0.8, kw["dt_max"]
    up.physics["rotating"] = kw.get("rotating", False)
    up.IO["datadir"] = ""
    return nyles.Nyles(up)


def _ic(kw):
    """Global seeded initial state (nz, ny, nx), without halos: a front in b plus velocity noise."""
    nx, ny_, nz = kw["n"]
    rng = np.random.default_rng(42)
    x = (np.arange(nx) + 0.5) * kw["L"][0] / nx
    z = (np.arange(nz) + 0.5) * kw["L"][2] / nz
    full = {}
    if kw["modelname"] == "LES":
        full["b"] = np.tanh((x[None, None, :] - 0.4 * kw["L"][0] + 0.3 * rng.standard_normal((nz, ny_, nx))) * 2.0) \
            + 0.2 * np.sin(2 * np.pi * z / kw["L"][2])[:, None, None]
    for d in "ijk":
        full["u_" + d] = 0.05 * (kw["L"][0] / nx) * rng.standard_normal((nz, ny_, nx))
    return full


def _interior(ny, name):
    st = ny.model.state
    k0, k1, j0, j1, i0, i1 = st.b.domainindices
    return st.get(name).tensor[k0:k1, j0:j1, i0:i1]


def _run(kw, npz):
    ny = _make(kw, npz)
    st = ny.model.state
    k0, k1, j0, j1, i0, i1 = st.b.domainindices
    nzl = k1 - k0
    z0 = ny.param["loc"][0] * nzl
    for name, arr in _ic(kw).items():
        t = st.get(name).tensor
        t[k0:k1, j0:j1, i0:i1] = torch.as_tensor(arr[z0:z0 + nzl], device=t.device)
    ny.model.halo.fill(st.b)
    ny.model.halo.fill(st.u)
    ny.model.diagnose_var(st)
    t, log = 0.0, [(0.0, ny.model.mg.stats["nite"])]
    for _ in range(kw["steps"] + 1):                      # Euler start-up step + LFAM3 steps
        dt = ny.compute_dt()
        ny.model.forward(t, dt)
        t += dt
        log.append((dt, ny.model.mg.stats["nite"]))
    torch.cuda.synchronize()
    return ny, log


def single_gpu_runs(world):
    """Phase A: no process group yet; this rank solves every whole problem on its own GPU."""
    assert not (dist.is_available() and dist.is_initialized()), "call before dist.init_process_group"
    out = []
    for kw in cases(world):
        ny, log = _run(kw, 1)
        out.append((kw, {f: _interior(ny, f).clone() for f in FIELDS}, log))
        del ny
    return out


def slab_runs(ref):
    """Phase B: the same problems on world slabs.  Returns (ok, report) -- identical on every rank."""
    world, rank = dist.get_world_size(), dist.get_rank()
    failures = []
    for kw, fields, rlog in ref:
        ny, log = _run(kw, world)
        nzl = kw["n"][2] // world
        for n, ((dt0, n0), (dt1, n1)) in enumerate(zip(rlog, log)):
            if n0 != n1 or dt0 != dt1:
                failures.append("%s: step %d dt / V-cycles %r on one GPU, %r on slabs" % (kw["name"], n, (dt0, n0), (dt1, n1)))
        for f in FIELDS:
            got, want = _interior(ny, f), fields[f][rank * nzl:(rank + 1) * nzl]
            if not torch.equal(got, want):
                failures.append("%s rank %d: %s differs from the single-GPU run (max abs %.3e)"
                                % (kw["name"], rank, f, (got - want).abs().max().item()))
        del ny
    flag = torch.tensor([len(failures)], device="cuda")
    dist.all_reduce(flag)
    report = {"ranks": world, "cases": [kw["name"] for kw, _, _ in ref], "fields": FIELDS,
              "steps": ref[0][0]["steps"] + 1, "criterion": "bit-equal fields, equal dt and V-cycle counts",
              "failures_all_ranks": int(flag.item()), "failures_this_rank": failures[:8]}
    return flag.item() == 0, report
