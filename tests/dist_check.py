"""Slab-decomposition parity, run under torchrun on N GPUs (not collected by pytest directly;
tests/test_gpu_slabs.py launches it when the box has >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_check.py

Phase A (no process group): every rank solves the WHOLE problem on its own GPU.
Phase B (NCCL): the same problem on z slabs.  Stencil arithmetic is independent of the
decomposition, so every field of every slab must equal the single-GPU result bit for bit, and the
V-cycle counts must be identical (the norms are summed in another order; they only feed the stop test).
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from nyles_b200 import lib, nyles, parameters          # noqa: E402
from nyles_b200.mgfordriver import MG                  # noqa: E402


def make_nyles(kw, npz):
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


def set_ic(ny, kw, full):
    """full[name]: global arrays (nz, ny, nx) without halos; write this rank's interior, then fill halos."""
    st = ny.model.state
    k0, k1, j0, j1, i0, i1 = st.b.domainindices
    nzl = k1 - k0
    z0 = ny.param["loc"][0] * nzl
    for name, arr in full.items():
        t = st.get(name).tensor
        t[k0:k1, j0:j1, i0:i1] = torch.as_tensor(arr[z0:z0 + nzl], device=t.device)
    ny.model.halo.fill(st.b)
    ny.model.halo.fill(st.u)


def interior(ny, name):
    st = ny.model.state
    k0, k1, j0, j1, i0, i1 = st.b.domainindices
    return st.get(name).tensor[k0:k1, j0:j1, i0:i1]


MODEL_CASES = [
    dict(modelname="LES", geometry="closed", n=(64, 32, 128), L=(4.0, 2.0, 8.0), dt_max=0.05, steps=6),
    dict(modelname="LES", geometry="perio_xy", n=(32, 64, 128), L=(2.0, 4.0, 8.0), dt_max=0.05, steps=5, rotating=True),
    dict(modelname="Euler3d", geometry="perio_xyz", n=(64, 64, 128), L=(2 * np.pi,) * 2 + (4 * np.pi,), dt_max=0.02, steps=5),
]
MG_CASES = [(64, 64, 128, 1), (128, 64, 256, 1), (64, 64, 128, 6), (64, 128, 256, 5)]
FIELDS = ["b", "u_i", "u_j", "u_k", "p", "ke", "vor_i", "vor_j", "vor_k", "div"]


def model_ic(kw):
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


def make_b(nx, ny_, nz, topo):
    """Padded global right-hand side, zero mean, halos filled as the model's halo fill of div does."""
    gen = torch.Generator(device="cuda").manual_seed(7)
    b = torch.zeros((nz + 6, ny_ + 6, nx + 6), dtype=torch.float64, device="cuda")
    inner = torch.randn((nz, ny_, nx), dtype=torch.float64, device="cuda", generator=gen)
    b[3:-3, 3:-3, 3:-3] = inner - inner.mean()
    if topo in (5, 6):
        b[:, :3, :] = b[:, -6:-3, :].clone(); b[:, -3:, :] = b[:, 3:6, :].clone()
        b[:, :, :3] = b[:, :, -6:-3].clone(); b[:, :, -3:] = b[:, :, 3:6].clone()
    if topo == 6:
        b[:3] = b[-6:-3].clone(); b[-3:] = b[3:6].clone()
    return b


def run_model(kw, npz):
    ny = make_nyles(kw, npz)
    set_ic(ny, kw, model_ic(kw))
    ny.model.diagnose_var(ny.model.state)
    t, log = 0.0, []
    for _ in range(kw["steps"]):
        dt = ny.compute_dt()
        ny.model.forward(t, dt)
        t += dt
        log.append((dt, ny.model.mg.stats["nite"]))
    torch.cuda.synchronize()
    return ny, log


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    L = lib.load()
    L.ny_mg_set_gather_cells(40000)                 # keep two levels distributed on these small grids
    L.ny_mg_set_overlap_cells(1 << 18)              # ... and let them take the overlapped-exchange path

    # ---------------- phase A: whole problems on one GPU
    ref_models = []
    for kw in MODEL_CASES:
        ny, log = run_model(kw, 1)
        ref_models.append(({f: interior(ny, f).clone() for f in FIELDS}, log))
        del ny
    ref_mg = []
    for (nx, ny_, nz, topo) in MG_CASES:
        g = MG(1, 1, nx, ny_, nz, 3, topo)
        bglob = make_b(nx, ny_, nz, topo)
        xs, its = [], []
        for rep in range(2):                        # the second solve warm-starts
            x = torch.zeros_like(bglob)
            g.solve(x, bglob * (1.0 + rep))
            xs.append(x.clone()); its.append((g.stats["nite"], list(g.stats["res"])))
        ref_mg.append((bglob, xs, its))
        del g

    # ---------------- phase B: slabs
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    failures = []
    for kw, (ref, rlog) in zip(MODEL_CASES, ref_models):
        ny, log = run_model(kw, world)
        nzl = kw["n"][2] // world
        for (dt0, n0), (dt1, n1) in zip(rlog, log):
            if n0 != n1 or abs(dt0 - dt1) > 1e-15 * abs(dt0):
                failures.append("%s/%s: dt or V-cycle count differs (%r vs %r)" % (kw["modelname"], kw["geometry"], (dt0, n0), (dt1, n1)))
        for f in FIELDS:
            got = interior(ny, f)
            want = ref[f][rank * nzl:(rank + 1) * nzl]
            if not torch.equal(got, want):
                err = (got - want).abs().max().item()
                failures.append("%s/%s rank %d: field %s differs from the single-GPU run (max abs %.3e)"
                                % (kw["modelname"], kw["geometry"], rank, f, err))
        del ny
    for (nx, ny_, nz, topo), (bglob, xs, its) in zip(MG_CASES, ref_mg):
        nzl = nz // world
        g = MG(1, 1, nx, ny_, nzl, 3, topo, npz=world)
        assert g.L.ny_mg_first_gathered_level(g.mg) >= 2
        win = slice(rank * nzl, rank * nzl + nzl + 6)       # this slab, halo planes included, in the padded global array
        for rep in range(2):
            g.set_array((bglob[win] * (1.0 + rep)).contiguous(), ivar=2)
            lib.check(g.L.ny_mg_solve(g.mg, lib.C.byref(g._stats), lib.stream()))
            g._record()
            x = g.get_array(ivar=1)
            nite, res = its[rep]
            want = xs[rep][win]
            if g.stats["nite"] != nite:
                failures.append("MG %r solve %d: %d V-cycles on slabs, %d on one GPU" % ((nx, ny_, nz, topo), rep, g.stats["nite"], nite))
            elif not np.allclose(g.stats["res"], res, rtol=1e-10, atol=0):
                failures.append("MG %r solve %d: residual history differs" % ((nx, ny_, nz, topo), rep))
            if not torch.equal(x, want):
                failures.append("MG %r solve %d rank %d: solution differs (max abs %.3e)"
                                % ((nx, ny_, nz, topo), rep, rank, (x - want).abs().max().item()))
        del g
    flag = torch.tensor([len(failures)], device="cuda")
    dist.all_reduce(flag)
    for f in failures:
        print("[rank %d] FAIL %s" % (rank, f), flush=True)
    if rank == 0:
        print("dist_check: %s (%d ranks, %d model cases, %d multigrid cases)"
              % ("OK" if flag.item() == 0 else "%d FAILURES" % flag.item(), world, len(MODEL_CASES), len(MG_CASES)), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 0 else 1)


if __name__ == "__main__":
    main()
