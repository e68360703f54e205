"""CPU: known-answer tests of the oracle.

The reference has no golden vectors for this path; what it has are `__main__` self-checks with a
stated expected value.  They are transcribed here against the C restatement, next to the
arithmetic facts the restatement must reproduce (REAL(4) literals of weno.f90, the closures of
flux1d, the skipped first cell of the kinetic energy) and the invariants SURVEY.md 8c lists.
"""
import numpy as np
import pytest

from oracle import model as M
from oracle.kernels import Kernels, OracleMG

K = Kernels("strict")


# ------------------------------------------------------------------ reference self-checks
def test_vorticity_of_solid_body_rotation():
    """core/vorticity.py:95-106: u = -Omega*y, v = Omega*x  =>  omega_k = 2*Omega (per unit cell area)."""
    n, Omega = 16, 2.0
    p = M.make_param(nx=n, ny=n, nz=n)
    st = M.get_state(p)
    x = np.linspace(0, n, n)                       # the reference's own coordinates: spacing n/(n-1)
    h = n / (n - 1.0)
    st.u["i"].view("i")[...] = -Omega * x[None, :, None]
    st.u["j"].view("i")[...] = Omega * x[None, None, :]
    M.vorticity(K, st, 0.0)
    wz = st.vor["k"].view("i")
    assert np.allclose(wz[:, :-1, :-1], 2 * Omega * h, rtol=0, atol=1e-12)
    assert np.all(wz[:, :-1, -1] == 0.0)           # fortran_vorticity.f90:24-26: last point of the sweep zeroed
    assert np.all(wz[:, -1, :] == 0.0)             # last row untouched (was zero)
    assert np.all(st.vor["i"].view("i") == 0.0) and np.all(st.vor["j"].view("i") == 0.0)
    # rotating frame: f*dx*dy added on all but the last row/column (vorticity.py:33-34)
    M.vorticity(K, st, 0.25)
    wz2 = st.vor["k"].view("i")
    assert np.allclose(wz2[:, :-1, :-1], 2 * Omega * h + 0.25, rtol=0, atol=1e-12)
    assert np.all(wz2[:, -1, :] == 0.0)


def test_kinetic_energy_of_uniform_flow():
    """core/kinenergy.py:60-77: max ke = 1/2 (u0^2/dx^2 + v0^2/dy^2 + w0^2/dz^2)."""
    n = 8
    p = M.make_param(nx=n, ny=n, nz=n, Lx=1.0, Ly=2.0, Lz=0.5)
    st, g = M.get_state(p), M.Grid(p)
    u0, v0, w0 = 0.8, -4.0, 0.6
    st.u["i"].view("i")[...] = u0
    st.u["j"].view("i")[...] = v0
    st.u["k"].view("i")[...] = w0
    M.kinenergy(K, st, g)
    ke = st.ke.view("i")
    want = 0.5 * (u0 ** 2 * g.idx2 + v0 ** 2 * g.idy2 + w0 ** 2 * g.idz2)
    assert abs(ke.max() - want) <= 4e-16 * want
    assert np.allclose(ke[1:, 1:, 1:], want, rtol=1e-15)
    # fortran_kinenergy.f90:20-24: the first cell of every sweep is skipped
    assert ke[0, 0, 0] == 0.0
    assert abs(ke[0, 1, 1] - 0.5 * (u0 ** 2 * g.idx2 + v0 ** 2 * g.idy2)) <= 1e-13


def test_compute_dt_cases():
    """core/nyles.py:306-321, the four cases printed by the reference."""
    p = M.make_param(nx=8, ny=8, nz=8, dt_max=1.5, cfl=1.0)
    m = M.LES(p)
    assert m.compute_dt() == 1.5
    for d, norm2 in zip("ijk", (1.0, 2.0, 3.0)):
        m.state.U[d].view("i")[...] += 1.0
        assert m.compute_dt() == min(p["cfl"] / np.sqrt(norm2), 1.5)
    p2 = M.make_param(nx=8, ny=8, nz=8, auto_dt=False, dt=0.125)
    assert M.LES(p2).compute_dt() == 0.125


# ------------------------------------------------------------------ weno.f90 arithmetic
def test_weno_literals_are_single_precision():
    """the literals of weno.f90:25-54 are REAL(4), so the three candidate stencils
    do not sum to one and constant data is reproduced only to ~1e-7."""
    c13, c76, c116 = np.float32(1. / 3.), np.float32(7. / 6.), np.float32(11. / 6.)
    c16, c56 = np.float32(1. / 6.), np.float32(5. / 6.)
    assert float(c13).hex() == "0x1.5555560000000p-2" and float(c76).hex() == "0x1.2aaaaa0000000p+0"
    assert float(c116).hex() == "0x1.d555560000000p+0" and float(c16).hex() == "0x1.5555560000000p-3"
    assert float(c56).hex() == "0x1.aaaaaa0000000p-1"
    q = 3.0
    # beta = 0 => w = (1, 6, 3)/10, the value is the weighted mean of the three biased candidates
    qi1 = float(c13) * q - float(c76) * q + float(c116) * q
    qi2 = -float(c16) * q + float(c56) * q + float(c13) * q
    qi3 = float(c13) * q + float(c56) * q - float(c16) * q
    want = ((1.0 * qi1 + 6.0 * qi2) + 3.0 * qi3) / ((1.0 + 6.0) + 3.0)
    got = K.weno5(q, q, q, q, q)
    assert got == want
    assert got != q and abs(got / q - 1.0) < 2e-7
    # weno3 has exact dyadic coefficients: constants and straight lines are exact
    assert K.weno3(q, q, q) == q
    assert K.weno3(1.0, 2.0, 3.0) == 2.5


def test_weno5_smooth_data_accuracy_and_upwinding():
    # finite-volume reconstruction: cell averages in, face value out
    h = 0.05
    xc = 0.3 + h * np.arange(5)
    avg = (np.cos(2.0 * (xc - h / 2)) - np.cos(2.0 * (xc + h / 2))) / (2.0 * h)
    xf = xc[2] + h / 2
    assert abs(K.weno5(*avg) - np.sin(2.0 * xf)) < 2e-6      # h^5 truncation + the 1e-7 literal bias
    # a step: the stencil that crosses it gets (almost) no weight
    v = K.weno5(0.0, 0.0, 0.0, 1.0, 1.0)
    assert -1e-6 < v < 0.05


def test_flux1d_closures():
    """weno.f90:106-153: which reconstruction each face uses near the two ends, both wind signs."""
    rng = np.random.default_rng(3)
    n = 12
    q = rng.standard_normal(n)
    for sign in (1.0, -1.0):
        u = sign * (0.5 + rng.random(n))
        F = K.flux1d(u, q)
        pos = sign > 0
        want = np.zeros(n)
        for i in range(1, n):                       # 1-based face i between cells i and i+1
            Q = lambda s: q[s - 1]                  # noqa: E731 (1-based cell access)
            if i == 1:
                r = Q(1) if pos else K.weno3(Q(3), Q(2), Q(1))
            elif i == 2:
                r = K.weno3(Q(1), Q(2), Q(3)) if pos else K.weno5(Q(5), Q(4), Q(3), Q(2), Q(1))
            elif i <= n - 3:
                r = K.weno5(*[Q(i + s) for s in (-2, -1, 0, 1, 2)]) if pos else \
                    K.weno5(*[Q(i + s) for s in (3, 2, 1, 0, -1)])
            elif i == n - 2:
                r = K.weno5(*[Q(i + s) for s in (-2, -1, 0, 1, 2)]) if pos else K.weno3(Q(i + 2), Q(i + 1), Q(i))
            else:
                r = K.weno3(Q(i - 1), Q(i), Q(i + 1)) if pos else Q(i + 1)
            want[i - 1] = u[i - 1] * r
        assert np.array_equal(F, want)
        assert F[-1] == 0.0                          # flux(n) = 0
    # the wind test is strictly u > 0: u == 0 takes the negative branch (and gives a zero flux)
    assert np.all(K.flux1d(np.zeros(n), q) == 0.0)


# ------------------------------------------------------------------ invariants
@pytest.mark.parametrize("geometry", ["closed", "perio_xyz"])
def test_tracer_advection_is_conservative(geometry):
    """Flux form: in a closed box (zero normal velocity) the sum of db vanishes up to round-off."""
    p = M.make_param(nx=16, ny=8, nz=8, geometry=geometry, Lx=2.0, Ly=1.0, Lz=1.0)
    m = M.LES(p)
    rng = np.random.default_rng(11)
    st = m.state
    st.b.view("i")[...] = rng.standard_normal(st.b.view("i").shape)
    for d in "ijk":
        st.u[d].view("i")[...] = 1e-2 * rng.standard_normal(st.b.view("i").shape)
    m.diagnose_var(st)                               # projection makes U divergence-free with closed walls
    ds = st.duplicate_prognostic_variables()
    m.rhs(st, 0.0, ds)
    db = ds.b.view("i")
    k0, k1, j0, j1, i0, i1 = st.b.domainindices
    total = db[k0:k1, j0:j1, i0:i1].sum()
    scale = np.abs(db).sum()
    if geometry == "closed":
        assert abs(total) <= 1e-13 * scale
    else:
        assert abs(total) <= 1e-12 * scale


def test_projection_removes_divergence():
    p = M.make_param(nx=16, ny=16, nz=16)
    m = M.LES(p)
    rng = np.random.default_rng(5)
    for d in "ijk":
        m.state.u[d].view("i")[...] = rng.standard_normal((16, 16, 16))
    # closed box: the wall-normal component on the last face is not a degree of freedom
    m.state.u["i"].view("i")[:, :, -1] = 0
    m.state.u["j"].view("i")[:, -1, :] = 0
    m.state.u["k"].view("i")[-1, :, :] = 0
    M.U_from_u(m.state, m.grid)
    M.compute_div(K, m.state)
    d0 = np.sum(m.state.div.view("i") ** 2)
    m.diagnose_var(m.state)
    nite, res, normb = m.mg_log[-1]
    assert 1 <= nite <= 20 and res < 1e-6
    M.compute_div(K, m.state)
    d1 = np.sum(m.state.div.view("i") ** 2)
    assert d1 / d0 < 1e-5


# ------------------------------------------------------------------ mgfor
def test_mg_operator_identities():
    """diag = 6 in the interior (fewer next to walls: Neumann), Rcoef = 0.5, Pcoef = 1/64."""
    mg = OracleMG(1, 1, 16, 16, 16, 3, topology=1)
    assert mg.nlevels == 4                           # 16 -> 8 -> 4 -> 2 (mg_setup.f90:225-307)
    assert mg.get_arrayshape(1) == (22, 22, 22) and mg.get_arrayshape(2) == (14, 14, 14)
    msk = mg.get_array(ivar=7)
    diag = mg.get_array(ivar=5)
    idiag = mg.get_array(ivar=6)
    assert msk[3:-3, 3:-3, 3:-3].min() == 1.0 and msk.sum() == 16 ** 3
    assert np.all(diag[4:-4, 4:-4, 4:-4] == 6.0)
    assert diag[3, 3, 3] == 3.0 and diag[3, 4, 4] == 5.0 and diag[3, 3, 4] == 4.0
    assert np.array_equal(idiag[3:-3, 3:-3, 3:-3], 1.0 / diag[3:-3, 3:-3, 3:-3])
    R = mg.get_array(ivar=8, lev=2)
    assert np.all(R[3:-3, 3:-3, 3:-3] == 0.5) and R.sum() == 0.5 * 8 ** 3
    P = mg.get_array(ivar=9, lev=1)
    assert np.all(P[5:-5, 5:-5, 5:-5] == 1.0 / 64.0)
    assert P[3, 3, 3] > 1.0 / 64.0                   # renormalised next to a wall
    # periodic box: every diagonal entry is 6 and the mask covers the z halo
    mgp = OracleMG(1, 1, 8, 8, 8, 3, topology=6)
    assert np.all(mgp.get_array(ivar=5)[3:-3, 3:-3, 3:-3] == 6.0)


def test_mg_point_source_converges():
    """core/mgfor/tests.f90:50-57: a +1/-1 pair of point sources; the reference records no expected
    residual, so the checks are the solver's own contract (solvers.f90:8-33)."""
    n = 32
    mg = OracleMG(1, 1, n, n, n, 3, topology=1)
    b = np.zeros(mg.shape)
    b[3 + n // 4, 3 + n // 4, 3 + n // 4] = 1.0
    b[3 + 3 * n // 4, 3 + 3 * n // 4, 3 + 3 * n // 4] = -1.0
    x = np.zeros(mg.shape)
    mg.solve(x, b)
    assert mg.normb == 2.0
    assert 1 <= mg.nite <= 20 and mg.res < 1e-6
    h = mg.reshist
    assert np.all(np.diff(h) < 0) and h[-1] == mg.res
    # x solves the 7-point Neumann problem: residual of the restated operator, computed independently
    X = x[3:-3, 3:-3, 3:-3]
    Xp = np.pad(X, 1, mode="edge")                   # homogeneous Neumann: mirror value => zero flux
    lap = (Xp[:-2, 1:-1, 1:-1] + Xp[2:, 1:-1, 1:-1] + Xp[1:-1, :-2, 1:-1] + Xp[1:-1, 2:, 1:-1] +
           Xp[1:-1, 1:-1, :-2] + Xp[1:-1, 1:-1, 2:] - 6.0 * X)
    r = b[3:-3, 3:-3, 3:-3] + (-lap)                 # fresidual3d: r = b + diag*x - sum6(x)
    assert np.sum(r ** 2) / 2.0 < 1e-6
    assert abs(np.sum(r ** 2) / 2.0 - mg.res) <= 1e-9 * mg.res + 1e-18


def test_mg_rejects_grids_it_cannot_coarsen():
    with pytest.raises(ValueError):
        OracleMG(1, 1, 12, 12, 8, 3, topology=1)      # 12 -> 6 -> 3 -> 1: never reaches nx == 2 or ny == 2


def test_mg_obstacle_mask():
    """A solid block inside the box (mgfor/tests.f90:207-212): no equation inside it (diag = 0), one neighbour
    fewer next to it, coarse masks follow, and the masked problem still converges."""
    n = 16
    mg = OracleMG(1, 1, n, n, n, 3, topology=1)
    msk = mg.get_array(ivar=7)
    msk[3 + 4:3 + 8, 3 + 4:3 + 8, 3 + 6:3 + 10] = 0.0
    mg.set_mask(msk)
    diag = mg.get_array(ivar=5)
    assert np.all(diag[3 + 4:3 + 8, 3 + 4:3 + 8, 3 + 6:3 + 10] == 0.0)
    assert diag[3 + 5, 3 + 5, 3 + 5] == 5.0 and diag[3 + 5, 3 + 5, 3 + 4] == 6.0      # next to the block / one further
    m2 = mg.get_array(ivar=7, lev=2)
    assert m2[3 + 2, 3 + 2, 3 + 3] == 0.0 and m2.sum() == 8 ** 3 - 2 * 2 * 2
    fluid = msk[3:-3, 3:-3, 3:-3] > 0
    rng = np.random.default_rng(2)
    inner = rng.standard_normal((n, n, n)) * fluid
    inner[fluid] -= inner[fluid].mean()
    b = np.zeros(mg.shape)
    b[3:-3, 3:-3, 3:-3] = inner
    x = np.zeros(mg.shape)
    mg.solve(x, b)
    assert 1 <= mg.nite <= 20 and mg.res < 1e-6
    assert np.all(x[3:-3, 3:-3, 3:-3][~fluid] == 0.0)
