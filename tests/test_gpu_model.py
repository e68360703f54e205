"""GPU parity, model level: the product's Nyles(param) / model_les / timescheme path against
(a) the committed fixtures produced by the reference's own Python drivers (tests/golden) and
(b) the oracle model run side by side (lock-exchange, BASELINE config 0: 100 steps).

north_star tolerances: per-field RHS <= 1e-12 relative, fields after 100 steps <= 1e-9 relative with
identical V-cycle counts.  The kernels are bit-exact, so (a) is checked with array_equal where the
multigrid did not hit a stopping-test tie, and with the stated tolerances otherwise."""
import os

import numpy as np
import pytest
import torch

from oracle import model as M
from golden_cases import CASES, SCALARS, VECTORS, flat_param, forcing_of, tracers

pytestmark = pytest.mark.gpu


def make_nyles(kw):
    from nyles_b200 import parameters, nyles
    parameters.InextensibleDict.unfreeze()
    up = parameters.UserParameters()
    up.model["modelname"] = kw.get("modelname", "LES")
    up.model["geometry"] = kw["geometry"]
    up.model["Lx"], up.model["Ly"], up.model["Lz"] = kw["Lx"], kw["Ly"], kw["Lz"]
    up.discretization["global_nx"], up.discretization["global_ny"], up.discretization["global_nz"] = \
        kw["nx"], kw["ny"], kw["nz"]
    up.time["cfl"], up.time["dt_max"] = kw.get("cfl", 0.8), kw.get("dt_max", 0.05)
    up.IO["datadir"] = ""
    for k in ("rotating", "coriolis", "diff_coef", "forced"):
        if k in kw:
            up.physics[k] = kw[k]
    if "timestepping" in kw:
        up.time["timestepping"] = kw["timestepping"]
    if "n_tracers" in kw:
        up.model["n_tracers"] = kw["n_tracers"]
    ny = nyles.Nyles(up)
    return ny


def relerr(a, b):
    d = np.max(np.abs(a - b))
    s = np.max(np.abs(b))
    return d / s if s > 0 else d


def set_ic(ny, g, euler, extra=()):
    st = ny.model.state
    if not euler:
        st.b.view("i")[:] = g["ic_b"]
    for d in "ijk":
        st.u[d].view("i")[:] = g["ic_u_" + d]
    for nick in extra:
        st.get(nick).view("i")[:] = g["ic_" + nick]


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("name", sorted(CASES))
def test_against_reference_driver_fixtures(name, fused, golden_dir):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    kw, nsteps = flat_param(name)
    ny = make_nyles(kw)
    ny.model.fused = fused
    # fused: the forcing object offers device_tendencies and the model stays on ny_rhs_step; unfused: the
    # reference's plain `add(state, dstate, t)` object
    ny.model.forcing = forcing_of(name, ny.param, ny.grid, device=fused)
    euler = kw["modelname"] == "Euler3d"
    set_ic(ny, g, euler, tracers(name))
    st = ny.model.state
    ny.model.diagnose_var(st)

    def compare(tag, tol):
        for s in tuple(SCALARS) + tuple(tracers(name)):
            a = getattr(st, s).tensor.cpu().numpy()
            assert relerr(a, g["%s_%s" % (tag, s)]) <= tol, "%s %s_%s" % (name, tag, s)
        for v in VECTORS:
            for d in "ijk":
                a = getattr(st, v)[d].tensor.cpu().numpy()
                assert relerr(a, g["%s_%s_%s" % (tag, v, d)]) <= tol, "%s %s_%s_%s" % (name, tag, v, d)

    compare("diag0", 1e-12)
    ds = st.duplicate_prognostic_variables()
    ny.model.rhs(st, 0.0, ds, last=True)
    assert relerr(ds.b.tensor.cpu().numpy(), g["rhs0_b"]) <= 1e-12
    for nick in tracers(name):
        assert relerr(ds.get(nick).tensor.cpu().numpy(), g["rhs0_" + nick]) <= 1e-12
    for d in "ijk":
        assert relerr(ds.u[d].tensor.cpu().numpy(), g["rhs0_u_" + d]) <= 1e-12
    t = 0.0
    for n in range(nsteps):
        dt = ny.compute_dt()
        assert abs(dt - g["dts"][n]) <= 1e-12 * g["dts"][n]
        ny.model.forward(t, dt)
        t += dt
    compare("final", 1e-10)


def lock_exchange_ic(shape, x, dx):
    rng = np.random.default_rng(1234)
    noise = 0.1 * rng.standard_normal(shape)
    return np.tanh((x + noise - 8) / (2 * dx))


@pytest.mark.parametrize("fast_arith", [False, True])
def test_lock_exchange_100_steps_vs_oracle(fast_arith, monkeypatch):
    """BASELINE config 0: experiments/lockechange/lockexchange.py at its default grid (128x32x32,
    closed, LES, LFAM3, cfl 0.8, dt_max 0.1), 100 steps, GPU against the oracle.
    Strict arithmetic (the default): dt agrees to 1e-12 at every step.  Fast arithmetic (opt-in): the
    few-ulp differences of the RHS accumulate, dt is held to the same 1e-9 as the fields."""
    import nyles_b200
    monkeypatch.setattr(nyles_b200, "FAST_ARITH", fast_arith)
    # fast mode: the lock exchange is unstable (Kelvin-Helmholtz billows), a few-ulp change of the RHS
    # grows by ~1e3 every 28 steps; it is compared over the first 20 steps only
    dt_tol = 1e-10 if fast_arith else 1e-12
    nsteps = 20 if fast_arith else 100
    kw = dict(nx=128, ny=32, nz=32, geometry="closed", Lx=32.0, Ly=8.0, Lz=8.0, cfl=0.8, dt_max=0.1)
    o = M.LES(M.make_param(**kw))
    ny = make_nyles(kw)
    ic = lock_exchange_ic(o.grid.x_b.shape, o.grid.x_b, o.grid.dx)
    o.state.b.view("i")[:] = ic
    # the product's own grid must give the same coordinates the experiment script would use
    assert np.array_equal(np.asarray(ny.grid.x_b.view("i")), o.grid.x_b)
    ny.model.state.b.view("i")[:] = ic
    o.diagnose_var(o.state)
    ny.model.diagnose_var(ny.model.state)
    t = 0.0
    gpu_cycles = []
    for n in range(nsteps):
        dt_o = o.compute_dt()
        dt_g = ny.compute_dt()
        assert abs(dt_o - dt_g) <= dt_tol * dt_o, "dt differs at step %d" % n
        before = ny.model.mg.nvcycles
        o.forward(t, dt_o)
        ny.model.forward(t, dt_g)
        gpu_cycles.append(ny.model.mg.nvcycles - before)
        t += dt_o
    ref_cycles = [c[0] for c in o.mg_log[1:]]
    assert sum(gpu_cycles) == sum(ref_cycles), "total V-cycles differ: %d vs %d" % (sum(gpu_cycles), sum(ref_cycles))
    st = ny.model.state
    assert relerr(st.b.tensor.cpu().numpy(), o.state.b.data) <= 1e-9
    for d in "ijk":
        assert relerr(st.u[d].tensor.cpu().numpy(), o.state.u[d].data) <= 1e-9
    assert float(st.u["i"].tensor.abs().max()) > (0.05 if nsteps == 100 else 0.005)      # the current actually developed


def test_tracer_conservation_and_projection_properties_large():
    """Size-independent properties at a size the oracle is not run at (256x128x128 closed box):
    flux-form advection conserves the tracer (sum db = 0 up to round-off) and the projection
    reduces the divergence norm by the solver tolerance."""
    kw = dict(nx=256, ny=128, nz=128, geometry="closed", Lx=2.0, Ly=1.0, Lz=1.0, cfl=0.8, dt_max=0.05)
    ny = make_nyles(kw)
    st = ny.model.state
    gen = torch.Generator(device="cuda").manual_seed(7)
    shape = st.b.tensor.shape
    st.b.tensor.copy_(torch.randn(shape, dtype=torch.float64, device="cuda", generator=gen))
    for d in "ijk":
        st.u[d].tensor.copy_(1e-3 * torch.randn(shape, dtype=torch.float64, device="cuda", generator=gen))
    # closed box: no flow through the walls (last face of each direction)
    st.u["i"].tensor[:, :, -1] = 0
    st.u["j"].tensor[:, -1, :] = 0
    st.u["k"].tensor[-1, :, :] = 0
    from nyles_b200 import cov_to_contra, projection
    cov_to_contra.U_from_u(st, ny.grid)
    projection.compute_div(st)
    div0 = float((st.div.tensor ** 2).sum())
    ny.model.diagnose_var(st)
    cov_to_contra.U_from_u(st, ny.grid)
    projection.compute_div(st)
    div1 = float((st.div.tensor ** 2).sum())
    assert div1 < 1e-5 * div0
    ds = st.duplicate_prognostic_variables()
    ny.model.rhs(st, 0.0, ds)
    total = float(ds.b.tensor.sum())
    scale = float(ds.b.tensor.abs().sum())
    assert abs(total) <= 1e-12 * scale


def test_views_survive_the_buffer_rotation_and_passive_tracers_follow():
    """The fused step rotates the buffers of b and u (LES.rhs_step).  A view taken before the run -- the
    way experiment scripts hold `b = state.b.view('i')` -- must keep showing the field, host round trips
    through step_host must see the rotated buffers, and a passive tracer (updated by the generic
    time-scheme kernels next to the fused fields) must match the oracle."""
    kw = dict(nx=32, ny=16, nz=16, geometry="closed", Lx=4.0, Ly=2.0, Lz=2.0, cfl=0.8, dt_max=0.05, n_tracers=1)
    o = M.LES(M.make_param(**kw))
    ny = make_nyles(kw)
    rng = np.random.default_rng(21)
    ic = np.tanh((o.grid.x_b - 2.0 + 0.1 * rng.standard_normal(o.grid.x_b.shape)) / (2 * o.grid.dx))
    t0 = np.cos(3.0 * o.grid.x_b) * np.sin(2.0 * o.grid.z_b)
    bview = ny.model.state.b.view("i")
    uview = ny.model.state.u["k"].view("j")
    for m in (o, ny.model):
        m.state.b.view("i")[:] = ic
        m.state.get("t0").view("i")[:] = t0
    o.diagnose_var(o.state)
    ny.model.diagnose_var(ny.model.state)
    t = 0.0
    host = ny.allocate_host_state()
    for n in range(5):
        dt = o.compute_dt()
        o.forward(t, dt)
        if n < 3:
            assert abs(ny.compute_dt() - dt) <= 1e-12 * dt
            ny.model.forward(t, dt)
        else:                                   # the same step through the host-buffer entry point
            for h, d in zip(host, ny.prognostic_tensors()):
                h.copy_(d)
            assert abs(ny.step_host(t, host) - dt) <= 1e-12 * dt
            for h, d in zip(host, ny.prognostic_tensors()):
                assert torch.equal(h, d.cpu())
        t += dt
        assert np.array_equal(np.asarray(bview), ny.model.state.b.tensor.cpu().numpy())
        assert np.array_equal(np.asarray(uview), ny.model.state.u["k"].tensor.permute(2, 0, 1).cpu().numpy())
    st = ny.model.state
    assert relerr(np.asarray(bview), o.state.b.data) <= 1e-11
    assert relerr(st.get("t0").tensor.cpu().numpy(), o.state.get("t0").data) <= 1e-11
    for d in "ijk":
        assert relerr(st.u[d].tensor.cpu().numpy(), o.state.u[d].data) <= 1e-11
    assert float(st.u["i"].tensor.abs().max()) > 0


@pytest.mark.parametrize("orders", [(5, 5), (3, 3), (1, 1)])
def test_linear_upwind_model_vs_oracle(orders, monkeypatch):
    """nyles_b200.LINEAR_UPWIND: the whole LES step with the reference's linear upwind branch (what Nyles does once
    `linear` is set in fortran_upwind.f90:31 / fortran_vortex_force.f90:28,108) at orderA / orderVF, against the oracle
    with the same switch: dt to 1e-12, fields to 1e-10 over a few steps, identical V-cycle counts."""
    import nyles_b200
    monkeypatch.setattr(nyles_b200, "LINEAR_UPWIND", True)
    oa, ovf = orders
    kw = dict(nx=32, ny=16, nz=16, geometry="closed", Lx=4.0, Ly=2.0, Lz=2.0, cfl=0.8, dt_max=0.05)
    o = M.LES(M.make_param(orderA=oa, orderVF=ovf, **kw), linear_upwind=True)
    from nyles_b200 import parameters, nyles
    parameters.InextensibleDict.unfreeze()
    up = parameters.UserParameters()
    up.model["geometry"] = "closed"
    up.model["Lx"], up.model["Ly"], up.model["Lz"] = kw["Lx"], kw["Ly"], kw["Lz"]
    up.discretization["global_nx"], up.discretization["global_ny"], up.discretization["global_nz"] = 32, 16, 16
    up.discretization["orderA"], up.discretization["orderVF"] = oa, ovf
    up.time["cfl"], up.time["dt_max"] = kw["cfl"], kw["dt_max"]
    up.IO["datadir"] = ""
    ny = nyles.Nyles(up)
    assert ny.model.tracer.linear and not ny.model.fused
    rng = np.random.default_rng(8)
    ic = np.tanh((o.grid.x_b - 2.0 + 0.1 * rng.standard_normal(o.grid.x_b.shape)) / (2 * o.grid.dx))
    o.state.b.view("i")[:] = ic
    ny.model.state.b.view("i")[:] = ic
    o.diagnose_var(o.state)
    ny.model.diagnose_var(ny.model.state)
    t = 0.0
    for n in range(5):
        dt = o.compute_dt()
        assert abs(ny.compute_dt() - dt) <= 1e-12 * dt
        nlog, before = len(o.mg_log), ny.model.mg.nvcycles
        o.forward(t, dt)
        ny.model.forward(t, dt)
        assert ny.model.mg.nvcycles - before == sum(m[0] for m in o.mg_log[nlog:])
        t += dt
    st = ny.model.state
    assert relerr(st.b.tensor.cpu().numpy(), o.state.b.data) <= 1e-10
    for d in "ijk":
        assert relerr(st.u[d].tensor.cpu().numpy(), o.state.u[d].data) <= 1e-10
    assert float(st.u["i"].tensor.abs().max()) > 0


def test_step_host_honours_edited_host_buffers():
    """step_host(modified=True): the caller has edited the host buffers, so the upload must be followed by
    diagnose_var and an Euler start-up step, as a run that begins from that state does (core/nyles.py:125,
    core/timescheme.py:131-139).  Checked against the oracle restarted from the same edited state."""
    kw = dict(nx=32, ny=16, nz=16, geometry="closed", Lx=4.0, Ly=2.0, Lz=2.0, cfl=0.8, dt_max=0.05)
    ny = make_nyles(kw)
    o = M.LES(M.make_param(**kw))
    rng = np.random.default_rng(33)
    ic = np.tanh((o.grid.x_b - 2.0 + 0.1 * rng.standard_normal(o.grid.x_b.shape)) / (2 * o.grid.dx))
    ny.model.state.b.view("i")[:] = ic
    ny.model.diagnose_var(ny.model.state)
    t = 0.0
    for n in range(3):
        dt = ny.compute_dt()
        ny.model.forward(t, dt)
        t += dt
    host = ny.allocate_host_state()
    for h, d in zip(host, ny.prognostic_tensors()):
        h.copy_(d)
    # the edit: a warm blob dropped into b, a kick added to u_i (names in the order of get_prognostic_scalars)
    names = ny.model.state.get_prognostic_scalars()
    hb, hu = host[names.index("b")], host[names.index("u_i")]
    hb[4:10, 4:10, 8:20] += 0.5
    hu[:, :, 10:14] += 0.01
    # the oracle starts a fresh run from the edited state
    for name, h in zip(names, host):
        o.state.get(name).view("i")[:] = h.numpy()
    o.diagnose_var(o.state)
    for n in range(3):
        dt_o = o.compute_dt()
        dt = ny.step_host(t, host, modified=(n == 0))
        assert abs(dt - dt_o) <= 1e-12 * dt_o, "dt differs at step %d after the edit" % n
        o.forward(t, dt_o)
        t += dt_o
    for name, h in zip(names, host):
        assert relerr(h.numpy(), o.state.get(name).data) <= 1e-10, name


def test_sub_views_follow_the_field_through_buffer_rotation():
    """A sub-view kept across steps (bint = b.view('i')[k0:k1]) stays bound to the Scalar, not to the buffer the
    field lived in when the view was taken (the fused step rotates buffers; NumPy views of the reference stay valid)."""
    kw = dict(nx=32, ny=16, nz=16, geometry="closed", Lx=4.0, Ly=2.0, Lz=2.0, cfl=0.8, dt_max=0.05)
    ny = make_nyles(kw)
    rng = np.random.default_rng(34)
    ny.model.state.b.view("i")[:] = np.tanh(rng.standard_normal((16, 16, 32)))
    ny.model.diagnose_var(ny.model.state)
    sub = ny.model.state.b.view("i")[2:5, :, 3:9]
    subj = ny.model.state.b.view("j")[3:9][:, 2:5]
    t = 0.0
    for n in range(3):
        dt = ny.compute_dt()
        ny.model.forward(t, dt)
        t += dt
        full = ny.model.state.b.tensor.cpu().numpy()
        assert np.array_equal(np.asarray(sub), full[2:5, :, 3:9])
        assert np.array_equal(np.asarray(subj), full.transpose(2, 0, 1)[3:9][:, 2:5])
    sub[:] = 7.0                                   # writes land in the live field
    assert float(ny.model.state.b.tensor[2:5, :, 3:9].min()) == 7.0


def test_vortex_force_work_diagnostic():
    """core/online_diag.py VFwork after a few steps of a developing flow: the work field equals the oracle's
    bit for bit (same elementwise order), its interior sum to round-off."""
    kw = dict(nx=32, ny=16, nz=16, geometry="closed", Lx=4.0, Ly=2.0, Lz=2.0, cfl=0.8, dt_max=0.05)
    o = M.LES(M.make_param(**kw))
    ny = make_nyles(kw)
    rng = np.random.default_rng(5)
    ic = np.tanh((o.grid.x_b - 2.0 + 0.1 * rng.standard_normal(o.grid.x_b.shape)) / (2 * o.grid.dx))
    o.state.b.view("i")[:] = ic
    ny.model.state.b.view("i")[:] = ic
    o.diagnose_var(o.state)
    ny.model.diagnose_var(ny.model.state)
    t = 0.0
    for n in range(4):
        dt = o.compute_dt()
        o.forward(t, dt)
        ny.model.forward(t, dt)
        t += dt
    ds = o.state.duplicate_prognostic_variables()
    want = M.vf_work(o.K, o.state, ds, o.state.work, o.state.b.domainindices)
    ny.diag.compute()
    got = ny.model.state.work.tensor.cpu().numpy()
    assert np.array_equal(got, o.state.work.data)
    assert np.abs(o.state.work.data).max() > 0
    assert abs(ny.diag.worksum - want) <= 1e-12 * np.abs(o.state.work.data).sum()


def test_run_writes_the_history_file(tmp_path, capsys):
    """Nyles.run() end to end (core/nyles.py:119-225): time loop, history cadence, the VFwork print, the final
    snapshot, the netCDF file in the reference's layout with the model's fields in it."""
    from scipy.io import netcdf_file
    from nyles_b200 import parameters, nyles
    parameters.InextensibleDict.unfreeze()
    up = parameters.UserParameters()
    up.model["geometry"] = "closed"
    up.model["Lx"], up.model["Ly"], up.model["Lz"] = 4.0, 2.0, 2.0
    up.discretization["global_nx"], up.discretization["global_ny"], up.discretization["global_nz"] = 32, 16, 16
    up.time["cfl"], up.time["dt_max"], up.time["tend"] = 0.8, 0.05, 0.22
    up.IO["datadir"], up.IO["expname"], up.IO["timestep_history"] = str(tmp_path), "run_test", 0.1
    up.IO["variables_in_history"] = "p+p"
    ny = nyles.Nyles(up)
    rng = np.random.default_rng(1)
    x = np.asarray(ny.grid.x_b.view("i"))
    ny.model.state.b.view("i")[:] = np.tanh((x - 2.0 + 0.1 * rng.standard_normal(x.shape)) / (2 * ny.grid.dx))
    ny.run()
    out = capsys.readouterr().out
    assert "Kdiss = " in out and "Job completed as expected" in out
    assert ny.n == 5 and abs(ny.t - 0.25) < 1e-12          # dt = dt_max while the flow is slow
    f = netcdf_file(ny.IO.hist_path, "r", mmap=False)
    f.__dict__["mode"] = "r"
    assert list(f.variables["n"][:]) == [0, 2, 4, 5]        # t = 0, 0.1, 0.2 and the final state
    assert np.allclose(f.variables["t"][:], [0.0, 0.1, 0.2, 0.25])
    assert np.array_equal(f.variables["b"][3], ny.model.state.b.tensor.cpu().numpy())
    assert np.array_equal(f.variables["u"][3], ny.model.state.u["i"].tensor.cpu().numpy())
    assert np.array_equal(f.variables["p"][3], ny.model.state.p.tensor.cpu().numpy())
    assert f.global_nx == 32 and f.geometry == b"closed"
    f.close()
    assert os.path.isfile(os.path.join(ny.IO.output_directory, "param.pkl"))
