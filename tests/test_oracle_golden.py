"""CPU: the oracle's NumPy drivers (oracle/model.py) reproduce, bit for bit, what the
reference's own Python drivers produced on the same kernels (tests/golden/*.npz)."""
import os

import numpy as np
import pytest

from oracle import model as M
from golden_cases import CASES, SCALARS, VECTORS, flat_param, forcing_of, tracers


def _run_oracle(name, g, flavour="strict"):
    kw, nsteps = flat_param(name)
    p = M.make_param(**kw)
    m = M.LES(p, flavour=flavour)
    m.forcing = forcing_of(name, p, m.grid)
    st = m.state
    if kw["modelname"] != "Euler3d":
        st.b.view("i")[:] = g["ic_b"]
    for d in "ijk":
        st.u[d].view("i")[:] = g["ic_u_" + d]
    for nick in tracers(name):
        st.get(nick).view("i")[:] = g["ic_" + nick]
    m.diagnose_var(st)
    out = {}

    def snap(tag):
        for s in tuple(SCALARS) + tuple(tracers(name)):
            out["%s_%s" % (tag, s)] = getattr(st, s).view("i").copy()
        for v in VECTORS:
            for d in "ijk":
                out["%s_%s_%s" % (tag, v, d)] = getattr(st, v)[d].view("i").copy()

    snap("diag0")
    ds = st.duplicate_prognostic_variables()
    m.rhs(st, 0.0, ds, last=True)
    out["rhs0_b"] = ds.b.view("i").copy()
    for nick in tracers(name):
        out["rhs0_" + nick] = ds.get(nick).view("i").copy()
    for d in "ijk":
        out["rhs0_u_" + d] = ds.u[d].view("i").copy()
    t, dts = 0.0, []
    for n in range(nsteps):
        dt = m.compute_dt()
        m.forward(t, dt)
        t += dt
        dts.append(dt)
    out["dts"] = np.array(dts)
    snap("final")
    return out


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_python(name, golden_dir):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    out = _run_oracle(name, g)
    for key, val in out.items():
        assert np.array_equal(val, g[key]), "%s: %s differs from the reference-driver fixture" % (name, key)


@pytest.mark.parametrize("name", ["les_closed", "euler_perio_xyz", "les_perio_xy_rot"])
def test_openmp_strict_build_equals_the_serial_one(name, golden_dir, monkeypatch):
    """The `strictomp` build (the checker of the 256^3 / 512^3 GPU parity tests) spreads the loops of the strict
    build over threads; every cell is still computed by the same instruction sequence, so its fields equal the
    fixtures bit for bit.  Only the multigrid norms are reduced in another order -- they decide when a solve
    stops, and on these cases they stop at the same V-cycle."""
    monkeypatch.setenv("OMP_NUM_THREADS", "4")
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    out = _run_oracle(name, g, flavour="strictomp")
    for key, val in out.items():
        assert np.array_equal(val, g[key]), "%s: %s differs between the OpenMP and the serial strict build" % (name, key)
