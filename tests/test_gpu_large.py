"""GPU parity AT THE BENCHMARKED SIZES: the CUDA path against the oracle on BASELINE configs[1]
(Taylor-Green vortex 256^3, perio_xyz, Euler3d; experiments/taylorgreen/tgv.py:68-84) and on a 256^3
Rayleigh-Taylor box (closed, LES; experiments/rayleightaylor/RT.py:69-71 -- configs[2] at half the edge).

The oracle here is the `strictomp` build (oracle/Makefile): the strict, source-order, no-FMA C restatement
with its loops spread over the host threads.  Per-cell arithmetic does not depend on the thread count, so its
fields are bit-identical to the serial strict build (tests/test_oracle_kat.py pins that on the CPU); only the two
multigrid norms are summed in another order.

Checked, in the order of core/model_les_euler.py:96-125 / core/model_les.py:98-144 / core/mgfor/solvers.f90:8-55:
  diag0   every field after the first diagnose_var (projection included)           <= 1e-12, V-cycle count equal
  rhs0    one bare right-hand side                                                 <= 1e-12 (bit-equal when the
          state it starts from is bit-equal, which it is unless a stop test ties)
  steps   1 Euler start-up + 3 LFAM3 steps: dt <= 1e-12, fields <= 1e-9, V-cycle counts equal

NY_LARGE_N=512 runs the same comparison at 512^3 (the benchmarked grid; ~40 GB of host memory and a few minutes of
CPU time -- recorded once per round under profiles/, not part of the routine suite)."""
import json
import os
import time

import numpy as np
import pytest
import torch

from oracle import model as M

pytestmark = pytest.mark.gpu

N = int(os.environ.get("NY_LARGE_N", "256"))
RECORD = os.environ.get("NY_LARGE_RECORD")          # path of a JSON file that receives the measured errors


def make_nyles(kw):
    from nyles_b200 import parameters, nyles
    parameters.InextensibleDict.unfreeze()
    up = parameters.UserParameters()
    up.model["modelname"] = kw.get("modelname", "LES")
    up.model["geometry"] = kw["geometry"]
    up.model["Lx"], up.model["Ly"], up.model["Lz"] = kw["Lx"], kw["Ly"], kw["Lz"]
    up.discretization["global_nx"], up.discretization["global_ny"], up.discretization["global_nz"] = \
        kw["nx"], kw["ny"], kw["nz"]
    up.time["cfl"], up.time["dt_max"] = kw["cfl"], kw["dt_max"]
    for k in ("rotating", "coriolis", "forced"):
        if k in kw:
            up.physics[k] = kw[k]
    up.IO["datadir"] = ""
    return nyles.Nyles(up)


def relerr(a, b):
    d = float(np.max(np.abs(a - b)))
    s = float(np.max(np.abs(b)))
    return d / s if s > 0 else d


def case_tgv(n):
    L = 2 * np.pi
    kw = dict(nx=n, ny=n, nz=n, geometry="perio_xyz", Lx=L, Ly=L, Lz=L, modelname="Euler3d", cfl=0.8,
              dt_max=0.4 * 32 / n)

    def ic(grid):                                   # tgv.py:68-84 (covariant components: times dx)
        x, y, z = grid.x_b_1D[None, None, :], grid.y_b_1D[None, :, None], grid.z_b_1D[:, None, None]
        dx = L / n
        u = np.sin(x + 1.2) * np.cos(y + 1.8) * np.cos(z + 0.5) * dx
        v = -np.cos(x + 1.2) * np.sin(y + 1.8) * np.cos(z + 0.5) * dx
        return {"u_i": u, "u_j": v}
    return kw, ic


def case_rt(n):
    dx = 0.25
    kw = dict(nx=n, ny=n, nz=n, geometry="closed", Lx=n * dx, Ly=n * dx, Lz=n * dx, modelname="LES", cfl=0.8,
              dt_max=0.1)

    def ic(grid):                                   # RT.py:69-71, plus a seeded velocity field so that the vortex
        rng = np.random.default_rng(1234)           # force and the WENO closures see non-trivial data from step 0
        shape = (n, n, n)
        z = grid.z_b_1D[:, None, None]
        x, y = grid.x_b_1D[None, None, :] / kw["Lx"], grid.y_b_1D[None, :, None] / kw["Ly"]
        zz = z / kw["Lz"]
        b = np.tanh((0.5 * kw["Lz"] - z + 0.01 * rng.standard_normal(shape)) / dx)
        amp = 0.3 * dx
        out = {"b": b}
        out["u_i"] = amp * (np.sin(2 * np.pi * x) * np.cos(2 * np.pi * y) * np.cos(np.pi * zz) + 0.1 * rng.standard_normal(shape))
        out["u_j"] = amp * (-np.cos(2 * np.pi * x) * np.sin(2 * np.pi * y) * np.cos(np.pi * zz) + 0.1 * rng.standard_normal(shape))
        out["u_k"] = amp * (0.3 * np.sin(2 * np.pi * zz) * np.sin(2 * np.pi * x) + 0.1 * rng.standard_normal(shape))
        return out
    return kw, ic


def case_plume(n):
    """configs[3] scaled down: experiments/forced_convection/forced_plume.py (closed, LES, rotating f = 1, Gaussian
    column heat source, linear stratification) on an n x n x n/2 box; the forcing object offers device_tendencies, so
    the CUDA side runs its fused RHS + time-scheme launches with the source added in-kernel."""
    nx, nz = n, n // 2
    kw = dict(nx=nx, ny=nx, nz=nz, geometry="closed", Lx=16.0, Ly=16.0, Lz=8.0, modelname="LES", cfl=0.8, dt_max=0.8,
              rotating=True, coriolis=1.0, forced=True)

    def ic(grid):
        rng = np.random.default_rng(4321)
        shape = (nz, nx, nx)
        z = grid.z_b_1D[:, None, None] / kw["Lz"]
        out = {"b": np.broadcast_to(0.1 * (z - 0.5), shape).copy()}
        dx = kw["Lx"] / nx
        for d in "ijk":                             # the plume starts from rest; a little noise wakes every term up
            out["u_" + d] = 0.05 * dx * rng.standard_normal(shape)
        return out
    return kw, ic


SCALARS = ("b", "p", "ke", "div")
VECTORS = ("u", "U", "vor")


@pytest.mark.parametrize("case", ["tgv", "rt", "plume"])
def test_benchmark_size_against_oracle(case):
    kw, ic = {"tgv": case_tgv, "rt": case_rt, "plume": case_plume}[case](N)
    euler = kw["modelname"] == "Euler3d"
    o = M.LES(M.make_param(**kw), flavour="strictomp")
    ny = make_nyles(kw)
    if kw.get("forced"):
        from golden_cases import DevicePlumeForcing, PlumeForcing
        o.forcing = PlumeForcing(o.param, o.grid)
        ny.model.forcing = DevicePlumeForcing(ny.param, ny.grid)
    assert np.array_equal(np.asarray(ny.grid.x_b_1D), o.grid.x_b_1D)
    fields = ic(o.grid)
    for name, a in fields.items():
        o.state.get(name).view("i")[:] = a
        ny.model.state.get(name).view("i")[:] = a
    del fields, a
    rec = {"case": case, "n": N, "param": {k: v for k, v in kw.items()}, "oracle": "strictomp (OpenMP, %d threads)"
           % (os.cpu_count() or 1)}

    def compare(tag, tol):
        worst, bitequal = 0.0, True
        st = ny.model.state
        for s in SCALARS:
            if euler and s == "b":
                continue
            a, r = getattr(st, s).tensor.cpu().numpy(), getattr(o.state, s).data
            e = relerr(a, r)
            bitequal &= bool(np.array_equal(a, r))
            assert e <= tol, "%s %s: %s differs from the oracle by %.3e" % (case, tag, s, e)
            worst = max(worst, e)
        for v in VECTORS:
            for d in "ijk":
                a, r = getattr(st, v)[d].tensor.cpu().numpy(), getattr(o.state, v)[d].data
                e = relerr(a, r)
                bitequal &= bool(np.array_equal(a, r))
                assert e <= tol, "%s %s: %s_%s differs from the oracle by %.3e" % (case, tag, v, d, e)
                worst = max(worst, e)
        rec[tag] = {"max_rel_err": worst, "bit_equal": bitequal}
        return bitequal

    t0 = time.time()
    o.diagnose_var(o.state)
    rec["oracle_diag0_s"] = time.time() - t0
    ny.model.diagnose_var(ny.model.state)
    assert ny.model.mg.stats["nite"] == o.mg_log[-1][0], "V-cycles of the first projection: %d vs %d" % (
        ny.model.mg.stats["nite"], o.mg_log[-1][0])
    same_state = compare("diag0", 1e-12)

    # one bare right-hand side
    ds_o = o.state.duplicate_prognostic_variables()
    o.rhs(o.state, 0.0, ds_o, last=True)
    ds = ny.model.state.duplicate_prognostic_variables()
    ny.model.rhs(ny.model.state, 0.0, ds, last=True)
    worst, bitequal = 0.0, True
    for name in ([] if euler else ["b"]) + ["u_i", "u_j", "u_k"]:
        a, r = ds.get(name).tensor.cpu().numpy(), ds_o.get(name).data
        e = relerr(a, r)
        assert e <= 1e-12, "%s rhs0: d%s differs from the oracle by %.3e" % (case, name, e)
        bitequal &= bool(np.array_equal(a, r))
        worst = max(worst, e)
    if same_state:
        assert bitequal, "%s: identical state, but the right-hand sides are not bit-equal" % case
    rec["rhs0"] = {"max_rel_err": worst, "bit_equal": bitequal}
    del ds, ds_o

    t, cyc_g, cyc_o, oracle_s = 0.0, [], [], []
    for n in range(4):                              # Euler start-up step + 3 LFAM3 steps
        dt_o, dt_g = o.compute_dt(), ny.compute_dt()
        assert abs(dt_o - dt_g) <= 1e-12 * dt_o, "dt differs at step %d: %r vs %r" % (n, dt_g, dt_o)
        before, nlog = ny.model.mg.nvcycles, len(o.mg_log)
        t0 = time.time()
        o.forward(t, dt_o)
        oracle_s.append(time.time() - t0)
        ny.model.forward(t, dt_g)
        cyc_g.append(ny.model.mg.nvcycles - before)
        cyc_o.append(sum(m[0] for m in o.mg_log[nlog:]))
        t += dt_o
    torch.cuda.synchronize()
    assert cyc_g == cyc_o, "V-cycles per step differ: %r vs %r" % (cyc_g, cyc_o)
    compare("final", 1e-9)
    rec.update(vcycles_per_step=cyc_g, oracle_seconds_per_step=oracle_s, t_end=t)
    if RECORD:
        try:
            with open(RECORD) as f:
                allrec = json.load(f)
        except (OSError, ValueError):
            allrec = []
        allrec.append(rec)
        with open(RECORD, "w") as f:
            json.dump(allrec, f, indent=1)
    print(json.dumps(rec))
