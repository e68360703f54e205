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


# ------------------------------------------------------------------ a second, independent restatement
def _weno3_py(qm, q0, qp):
    """core/weno.f90:1-22 written again from the source, in Python floats (IEEE double, no contraction);
    un-suffixed literals are REAL(4)."""
    f = np.float32
    eps = float(f(1e-14))
    qi1 = (-qm + 3 * q0) * 0.5
    qi2 = (q0 + qp) * 0.5
    d1, d2 = q0 - qm, qp - q0
    beta1, beta2 = d1 * d1, d2 * d2
    tau = abs(beta2 - beta1)
    w1 = 1. + tau / (beta1 + eps)
    w2 = (1. + tau / (beta2 + eps)) * 2
    return (w1 * qi1 + w2 * qi2) / (w1 + w2)


def _weno5_py(qmm, qm, q0, qp, qpp):
    """core/weno.f90:25-54, same rules; tau5 is implicitly typed REAL(4)."""
    f = np.float32
    eps = float(f(1e-16))
    c13, c76, c116 = float(f(1.) / f(3.)), float(f(7.) / f(6.)), float(f(11.) / f(6.))
    c16, c56 = float(f(1.) / f(6.)), float(f(5.) / f(6.))
    qi1 = c13 * qmm - c76 * qm + c116 * q0
    qi2 = -(c16 * qm) + c56 * q0 + c13 * qp
    qi3 = c13 * q0 + c56 * qp - c16 * qpp
    k1, k2 = float(f(13.) / f(12.)), .25
    a1, a2 = qmm - 2 * qm + q0, qmm - 4 * qm + 3 * q0
    b1, b2 = qm - 2 * q0 + qp, qm - qp
    g1, g2 = q0 - 2 * qp + qpp, 3 * q0 - 4 * qp + qpp
    beta1 = k1 * (a1 * a1) + k2 * (a2 * a2)
    beta2 = k1 * (b1 * b1) + k2 * (b2 * b2)
    beta3 = k1 * (g1 * g1) + k2 * (g2 * g2)
    tau5 = float(f(abs(beta1 - beta3)))
    w1 = 1. + tau5 / (beta1 + eps)
    w2 = 6 * (1. + tau5 / (beta2 + eps))
    w3 = 3 * (1. + tau5 / (beta3 + eps))
    return (w1 * qi1 + w2 * qi2 + w3 * qi3) / (w1 + w2 + w3)


def _flux1d_py(u, q):
    """core/weno.f90:106-153 with 1-based indices kept."""
    n = len(u)
    U = lambda i: float(u[i - 1])          # noqa: E731
    Q = lambda i: float(q[i - 1])          # noqa: E731
    flux = np.zeros(n)
    flux[0] = U(1) * Q(1) if U(1) > 0 else U(1) * _weno3_py(Q(3), Q(2), Q(1))
    flux[1] = U(2) * _weno3_py(Q(1), Q(2), Q(3)) if U(2) > 0 else U(2) * _weno5_py(Q(5), Q(4), Q(3), Q(2), Q(1))
    for i in range(3, n - 2):              # do i = 3, n-3
        if U(i) > 0:
            flux[i - 1] = U(i) * _weno5_py(Q(i - 2), Q(i - 1), Q(i), Q(i + 1), Q(i + 2))
        else:
            flux[i - 1] = U(i) * _weno5_py(Q(i + 3), Q(i + 2), Q(i + 1), Q(i), Q(i - 1))
    i = n - 2
    flux[i - 1] = U(i) * _weno5_py(Q(i - 2), Q(i - 1), Q(i), Q(i + 1), Q(i + 2)) if U(i) > 0 else \
        U(i) * _weno3_py(Q(i + 2), Q(i + 1), Q(i))
    i = n - 1
    flux[i - 1] = U(i) * _weno3_py(Q(i - 1), Q(i), Q(i + 1)) if U(i) > 0 else U(i) * Q(i + 1)
    flux[n - 1] = 0.
    return flux


def test_c_restatement_equals_an_independent_python_restatement():
    """The C oracle (oracle/csrc/oracle_rhs.c) and a second transcription of weno.f90 made independently in
    Python must agree bit for bit: smooth data, steps, tiny and huge magnitudes, still regions (beta = 0),
    both wind signs, every line length down to the minimum of 5."""
    rng = np.random.default_rng(77)
    for scale in (1.0, 1e-8, 1e6, 1e-30):
        for _ in range(300):
            v = scale * rng.standard_normal(5)
            assert K.weno5(*v) == _weno5_py(*[float(x) for x in v])
            assert K.weno3(*v[:3]) == _weno3_py(*[float(x) for x in v[:3]])
    for v in ([0.0] * 5, [1.0] * 5, [0, 0, 0, 1, 1], [1, 1, 0, 0, 0], [1e-200, 0, 0, 0, -1e-200], [2.5, 2.5, 2.5, 2.5, 3.5]):
        v = [float(x) for x in v]
        assert K.weno5(*v) == _weno5_py(*v)
        assert K.weno3(*v[:3]) == _weno3_py(*v[:3])
    for n in (5, 6, 7, 8, 13, 40):
        for rep in range(5):
            q = rng.standard_normal(n) * (10.0 ** rng.integers(-3, 3))
            u = rng.standard_normal(n)
            if rep == 0:
                u[:] = np.abs(u)
            if rep == 1:
                u[:] = -np.abs(u)
            if rep == 2:
                u[::3] = 0.0
            assert np.array_equal(K.flux1d(u, q), _flux1d_py(u, q)), "n = %d" % n


def _loops(shape):
    import itertools
    return itertools.product(*[range(s) for s in shape])


def test_c_kernels_equal_independent_python_loops():
    """The remaining Fortran kernels of the right-hand side written a second time as plain Python loops over
    the source (fortran_vortex_force.f90:10-165, fortran_upwind.f90:66-82, fortran_vorticity.f90:2-28,
    fortran_kinenergy.f90:43-50, fortran_bernoulli.f90:2-97, fortran_dissipation.f90:2-35) against the C oracle,
    bit for bit, on small ragged arrays."""
    rng = np.random.default_rng(99)
    for shape in [(5, 6, 7), (6, 5, 5), (3, 7, 9)]:
        m_, n_, l_ = shape
        U, vort, res0 = (rng.standard_normal(shape) for _ in range(3))
        # vortex_force_direc: sweep along the last axis k, U averaged over (i, i+1), rows i < n-1
        want = res0.copy()
        for j in range(m_):
            for i in range(n_ - 1):
                UU_0, u1d = 0., np.zeros(l_)
                for k in range(l_):
                    UU_1 = 0.5 * (U[j, i, k] + U[j, i + 1, k])
                    u1d[k] = 0.5 * (UU_0 + UU_1)
                    UU_0 = UU_1
                q = np.concatenate(([0.], vort[j, i, :l_ - 1]))
                want[j, i, :] = want[j, i, :] - _flux1d_py(u1d, q)
        got = res0.copy()
        K.vortex_force_direc(U, vort, got)
        assert np.array_equal(got, want)
        # vortex_force_flip: sweep along the middle axis i, U averaged over (k, k+1), columns k < l-1, sign +
        want = res0.copy()
        for j in range(m_):
            for k in range(l_ - 1):
                UU_0, u1d = 0., np.zeros(n_)
                for i in range(n_):
                    UU_1 = 0.5 * (U[j, i, k] + U[j, i, k + 1])
                    u1d[i] = 0.5 * (UU_0 + UU_1)
                    UU_0 = UU_1
                q = np.concatenate(([0.], vort[j, :n_ - 1, k]))
                want[j, :, k] = want[j, :, k] + _flux1d_py(u1d, q)
        got = res0.copy()
        K.vortex_force_flip(U, vort, got)
        assert np.array_equal(got, want)
        # upwind: dtrac(1) -= flux(1); dtrac(i) += flux(i-1) - flux(i)
        trac, u, d0 = (rng.standard_normal(shape) for _ in range(3))
        want = d0.copy()
        for k, j in _loops(shape[:2]):
            flux = _flux1d_py(u[k, j], trac[k, j])
            want[k, j, 0] = want[k, j, 0] - flux[0]
            for i in range(1, shape[2]):
                want[k, j, i] = want[k, j, i] + flux[i - 1] - flux[i]
        got = d0.copy()
        K.upwind(trac, u, got)
        assert np.array_equal(got, want)
        # vorticity: rows j < m-1, last column zero, last row untouched
        ui, uj = rng.standard_normal(shape), rng.standard_normal(shape)
        want = np.full(shape, 7.0)
        for k in range(shape[0]):
            for j in range(shape[1] - 1):
                for i in range(shape[2] - 1):
                    want[k, j, i] = uj[k, j, i + 1] - uj[k, j, i] - ui[k, j + 1, i] + ui[k, j, i]
                want[k, j, -1] = 0.
        got = np.full(shape, 7.0)
        K.vorticity(ui, uj, got)
        assert np.array_equal(got, want)
        # kin: ke(i) += cff2*0.5*(u(i)^2 + u(i-1)^2), i >= 2
        ke0, ds2 = rng.standard_normal(shape), 37.3
        want, cff2 = ke0.copy(), 0.5 * ds2
        for k, j in _loops(shape[:2]):
            for i in range(1, shape[2]):
                want[k, j, i] = want[k, j, i] + cff2 * 0.5 * (u[k, j, i] * u[k, j, i] + u[k, j, i - 1] * u[k, j, i - 1])
        got = ke0.copy()
        K.kin(u, u, got, ds2)
        assert np.array_equal(got, want)
        # gradke / gradkeandb / div / add_laplacian
        ke, b, du0, dz = rng.standard_normal(shape), rng.standard_normal(shape), rng.standard_normal(shape), 0.31
        want1, want2, cff = du0.copy(), du0.copy(), 0.5 * dz
        wdiv0, wdiv1 = np.zeros(shape), du0.copy()
        wlap, coef = du0.copy(), 0.013
        for k, j in _loops(shape[:2]):
            fxm = 0.
            for i in range(shape[2] - 1):
                want1[k, j, i] = want1[k, j, i] - (ke[k, j, i + 1] - ke[k, j, i])
                want2[k, j, i] = want2[k, j, i] - (ke[k, j, i + 1] - ke[k, j, i]) + cff * (b[k, j, i + 1] + b[k, j, i])
                fx = ke[k, j, i + 1] - ke[k, j, i]
                wlap[k, j, i] = wlap[k, j, i] + coef * (fx - fxm)
                fxm = fx
            wlap[k, j, -1] = wlap[k, j, -1] + coef * (0. - fxm)
            wdiv0[k, j, 0] = u[k, j, 0]
            wdiv1[k, j, 0] = wdiv1[k, j, 0] + u[k, j, 0]
            for i in range(1, shape[2]):
                wdiv0[k, j, i] = u[k, j, i] - u[k, j, i - 1]
                wdiv1[k, j, i] = wdiv1[k, j, i] + (u[k, j, i] - u[k, j, i - 1])
        got = du0.copy(); K.gradke(ke, got); assert np.array_equal(got, want1)
        got = du0.copy(); K.gradkeandb(ke, b, got, dz); assert np.array_equal(got, want2)
        got = np.full(shape, 5.0); K.div(got, u, 0); assert np.array_equal(got, wdiv0)
        got = du0.copy(); K.div(got, u, 1); assert np.array_equal(got, wdiv1)
        got = du0.copy(); K.add_laplacian(ke, got, coef); assert np.array_equal(got, wlap)


def _wrap(a, nh, xper, yper, zper):
    """Periodic halo fill of a padded (nz, ny+2nh, nx+2nh) level array whose interior is at least nh wide in
    every wrapped direction (mod_halo.f90:200-262 then reduces to: every halo cell of a wrapped direction is
    a copy of the interior cell one period away; z planes are copied whole, after x and y)."""
    nz, ny, nx = a.shape[0] - 2 * nh, a.shape[1] - 2 * nh, a.shape[2] - 2 * nh
    if xper:
        a[nh:-nh, nh:-nh, :nh] = a[nh:-nh, nh:-nh, nx:nx + nh]
        a[nh:-nh, nh:-nh, -nh:] = a[nh:-nh, nh:-nh, nh:2 * nh]
    if yper:
        a[nh:-nh, :nh, nh:-nh] = a[nh:-nh, ny:ny + nh, nh:-nh]
        a[nh:-nh, -nh:, nh:-nh] = a[nh:-nh, nh:2 * nh, nh:-nh]
    if xper and yper:
        for js, jd in ((slice(ny, ny + nh), slice(0, nh)), (slice(nh, 2 * nh), slice(-nh, None))):
            for is_, id_ in ((slice(nx, nx + nh), slice(0, nh)), (slice(nh, 2 * nh), slice(-nh, None))):
                a[nh:-nh, jd, id_] = a[nh:-nh, js, is_]
    if zper:
        a[:nh] = a[nz:nz + nh]
        a[-nh:] = a[nh:2 * nh]


@pytest.mark.parametrize("topology", [1, 6])
def test_mg_operators_equal_independent_numpy_transcription(topology):
    """mgfor's four level operators (basicoperators.f90:32-60, 173-231, 300-323, 363-400) written a second time
    with NumPy slices in the source's summation order, against the C oracle, bit for bit -- closed box and
    triply periodic box, random x / b including their halos."""
    nh, n = 3, 8
    mg = OracleMG(1, 1, n, n, n, nh, topology=topology)
    per = topology == 6
    rng = np.random.default_rng(12)
    shape1 = mg.get_arrayshape(1)
    x, b = rng.standard_normal(shape1), rng.standard_normal(shape1)
    if per:
        _wrap(x, nh, 1, 1, 1)
        _wrap(b, nh, 1, 1, 1)
    omega, cff1 = 0.9, 1.0 - 0.9
    idiag, diag, msk = mg.get_array(ivar=6), mg.get_array(ivar=5), mg.get_array(ivar=7)

    def S(a, ks, js, is_):
        sh = lambda s, d: slice(s.start + d, s.stop + d)      # noqa: E731
        return a[ks, js, sh(is_, -1)] + a[ks, js, sh(is_, 1)] + a[ks, sh(js, -1), is_] + a[ks, sh(js, 1), is_] + \
            a[sh(ks, -1), js, is_] + a[sh(ks, 1), js, is_]

    # ---- smooth: sweep 1 on the interior + 1 ring, sweep 2 on the interior, fill(x)
    mg.set_array(x, ivar=1)
    mg.set_array(b, ivar=2)
    mg.set_array(np.zeros(shape1), ivar=4)
    mg.op("smooth", 1)
    y = np.zeros(shape1)
    r1 = (slice(nh - 1, shape1[0] - nh + 1), slice(nh - 1, nh + n + 1), slice(nh - 1, nh + n + 1))
    y[r1] = cff1 * x[r1] + omega * (S(x, *r1) - b[r1]) * idiag[r1]
    it = (slice(nh, shape1[0] - nh), slice(nh, nh + n), slice(nh, nh + n))
    xs = x.copy()
    xs[it] = cff1 * y[it] + omega * (S(y, *it) - b[it]) * idiag[it]
    if per:
        _wrap(xs, nh, 1, 1, 1)
    assert np.array_equal(mg.get_array(ivar=1), xs)
    # ---- residual: r = msk*(b + diag*x - S(x)) on the interior, fill(r)
    mg.op("residual", 1)
    r = np.zeros(shape1)
    r[it] = msk[it] * (b[it] + diag[it] * xs[it] - S(xs, *it))
    if per:
        _wrap(r, nh, 1, 1, 1)
    got_r = mg.get_array(ivar=3)
    assert np.array_equal(got_r[it], r[it])
    if per:
        assert np.array_equal(got_r, r)
    # ---- restriction: b_c = Rcoef * sum of the 8 fine residuals in source order; x_c = 0; fill(b_c)
    mg.op("restriction", 1)
    shape2 = mg.get_arrayshape(2)
    R = mg.get_array(ivar=8, lev=2)
    nc = n // 2
    bc = np.zeros(shape2)
    for kc in range(nc):
        for jc in range(nc):
            for ic in range(nc):
                k, j, i = nh + 2 * kc, nh + 2 * jc, nh + 2 * ic
                s = r[k, j, i] + r[k, j, i + 1] + r[k, j + 1, i] + r[k, j + 1, i + 1] + \
                    r[k + 1, j, i] + r[k + 1, j, i + 1] + r[k + 1, j + 1, i] + r[k + 1, j + 1, i + 1]
                bc[nh + kc, nh + jc, nh + ic] = R[nh + kc, nh + jc, nh + ic] * s
    if per:
        _wrap(bc, nh, 1, 1, 1)
    got_bc = mg.get_array(ivar=2, lev=2)
    assert np.array_equal(got_bc[nh:-nh, nh:-nh, nh:-nh], bc[nh:-nh, nh:-nh, nh:-nh])
    if per:
        assert np.array_equal(got_bc, bc)
    assert not mg.get_array(ivar=1, lev=2).any()
    # ---- prolongation: x_f += Pcoef * (3 b + a | c) with the 9-3-3-1 weights of the nearest coarse cells; fill(x_f)
    xc = rng.standard_normal(shape2)
    if per:
        _wrap(xc, nh, 1, 1, 1)
    else:
        xc[:nh] = 0; xc[-nh:] = 0; xc[:, :nh] = 0; xc[:, -nh:] = 0; xc[:, :, :nh] = 0; xc[:, :, -nh:] = 0
    mg.set_array(xc, ivar=1, lev=2)
    mg.op("prolongation", 1)
    P = mg.get_array(ivar=9, lev=1)
    xf = xs.copy()
    for k in range(n):
        for j in range(n):
            for i in range(n):
                kc, jc, ic = nh + k // 2, nh + j // 2, nh + i // 2
                dk, dj, di = (1 if k % 2 else -1), (1 if j % 2 else -1), (1 if i % 2 else -1)

                def plane(kk):
                    return 9 * xc[kk, jc, ic] + 3 * xc[kk, jc, ic + di] + 3 * xc[kk, jc + dj, ic] + xc[kk, jc + dj, ic + di]
                xf[nh + k, nh + j, nh + i] = xf[nh + k, nh + j, nh + i] + \
                    P[nh + k, nh + j, nh + i] * (3 * plane(kc) + plane(kc + dk))
    if per:
        _wrap(xf, nh, 1, 1, 1)
    assert np.array_equal(mg.get_array(ivar=1), xf)


# ---------------------------------------------------------------- the linear (non-WENO) upwind branch
# Second, independent transcription (plain Python loops, straight from the Fortran) of core/interpolate.f90,
# core/interpolate_tracer.f90 and of the `if (linear)` branches of fortran_upwind.f90:33-64 and
# fortran_vortex_force.f90:39-64,118-143; must agree with the C restatement bit for bit.
_f32 = np.float32
_LC = [float(-(_f32(1.) / _f32(6.))), float(_f32(5.) / _f32(6.)), float(_f32(2.) / _f32(6.))]
_LE = [float(-(_f32(1.) / _f32(12.))), float(_f32(7.) / _f32(12.))]
_LB = [float(_f32(2.) / _f32(60.)), float(-(_f32(13.) / _f32(60.))), float(_f32(47.) / _f32(60.)),
       float(_f32(27.) / _f32(60.)), float(-(_f32(3.) / _f32(60.)))]


def _py_interpolate(v, order, tracer):
    """Returns dicts qp, qm keyed by the Fortran index."""
    n = len(v)
    q = lambda i: v[i - 1]                         # noqa: E731
    c1, c2, c3 = _LC
    e1, e2 = _LE
    b1, b2, b3, b4, b5 = _LB
    qp, qm = {}, {}

    def third(i):
        qp[i] = c1 * q(i - 1) + c2 * q(i) + c3 * q(i + 1)
        qm[i] = c3 * q(i - 1) + c2 * q(i) + c1 * q(i + 1)

    def fifth(i):
        qp[i] = b1 * q(i - 2) + b2 * q(i - 1) + b3 * q(i) + b4 * q(i + 1) + b5 * q(i + 2)
        qm[i] = b5 * q(i - 2) + b4 * q(i - 1) + b3 * q(i) + b2 * q(i + 1) + b1 * q(i + 2)

    def copy(i):
        qp[i] = q(i); qm[i] = q(i)
    if tracer:                                     # interpolate_tracer.f90
        if order == 5:
            copy(1); third(2)
            for i in range(3, n - 1):
                fifth(i)
            third(n - 1); copy(n)
        elif order == 3:
            copy(1)
            for i in range(2, n):
                third(i)
            copy(n)
        elif order == 1:
            for i in range(1, n + 1):
                copy(i)
        elif order == 2:
            for i in range(1, n):
                qp[i] = 0.5 * (q(i) + q(i + 1))
        elif order == 4:
            qp[1] = e2 * (q(1) + q(2)) + e1 * (q(3))
            for i in range(2, n - 1):
                qp[i] = e2 * (q(i) + q(i + 1)) + e1 * (q(i - 1) + q(i + 2))
            i = n - 1
            qp[i] = e2 * (q(i) + q(i + 1)) + e1 * (q(i - 1))
    else:                                          # interpolate.f90
        if order == 5:
            qp[0] = 0.0
            copy(1); third(2)
            for i in range(3, n - 2):
                fifth(i)
            third(n - 2); copy(n - 1)
            qm[n] = q(n)
        elif order == 3:
            qp[0] = 0.0
            copy(1)
            for i in range(2, n - 1):
                third(i)
            copy(n - 1)
            qm[n] = q(n)
        elif order == 1:
            qp[0] = 0.0
            for i in range(1, n):
                copy(i)
            qm[n] = q(n)
        elif order == 2:
            qm[1] = 0.5 * q(1)
            for i in range(2, n + 1):
                qm[i] = 0.5 * (q(i - 1) + q(i))
        elif order == 4:
            qm[1] = e2 * (q(1)) + e1 * (q(2))
            qm[2] = e2 * (q(1) + q(2)) + e1 * (q(3))
            for i in range(3, n):
                qm[i] = e2 * (q(i - 1) + q(i)) + e1 * (q(i - 2) + q(i + 1))
            qm[n] = e2 * (q(n - 1) + q(n)) + e1 * (q(n - 2))
    return qp, qm


@pytest.mark.parametrize("order", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("n", [5, 6, 7, 12])
def test_linear_interpolations_second_transcription(order, n):
    rng = np.random.default_rng(100 + order + n)
    v = rng.standard_normal(n) * 10.0 ** rng.integers(-3, 4)
    for tracer in (False, True):
        qp, qm = _py_interpolate(v, order, tracer)
        cp, cm = (K.interpolate_tr if tracer else K.interpolate_vf)(v, order)
        for i, val in qp.items():
            got = cp[i - 1] if tracer else cp[i]          # the vortex-force flavour stores qp(0:n-1)
            assert got == val, (tracer, "qp", i)
        for i, val in qm.items():
            assert cm[i - 1] == val, (tracer, "qm", i)
        # nothing else is assigned
        assert np.count_nonzero(~np.isnan(cp)) == len(qp) and np.count_nonzero(~np.isnan(cm)) == len(qm)


@pytest.mark.parametrize("order", [1, 2, 3, 4, 5])
def test_linear_upwind_and_vortex_force_second_transcription(order):
    rng = np.random.default_rng(200 + order)
    l, m, n = 3, 4, 9
    trac, u = rng.standard_normal((l, m, n)), rng.standard_normal((l, m, n))
    d0 = rng.standard_normal((l, m, n))
    ref = d0.copy()
    for k in range(l):
        for j in range(m):
            up = [0.5 * (x + abs(x)) for x in u[k, j]]
            um = [0.5 * (x - abs(x)) for x in u[k, j]]
            qp, qm = _py_interpolate(trac[k, j], order, True)
            fxm = 0.0
            for i in range(1, n):
                fx = u[k, j, i - 1] * qp[i] if order % 2 == 0 else up[i - 1] * qp[i] + um[i - 1] * qm[i + 1]
                ref[k, j, i - 1] = ref[k, j, i - 1] + fxm - fx
                fxm = fx
            ref[k, j, n - 1] = ref[k, j, n - 1] + fxm - 0.0
    got = d0.copy()
    K.upwind_linear(trac, u, got, order)
    assert np.array_equal(got, ref)
    # vortex_force_direc / _flip, arrays (m, n, l) = [j, i, k]
    m, n, l = 3, 7, 8
    U, w, r0 = (rng.standard_normal((m, n, l)) for _ in range(3))
    ref = r0.copy()
    for j in range(m):
        for i in range(n - 1):
            UU_0, vU, up, um = 0.0, [], [], []
            for k in range(l):
                UU_1 = 0.5 * (U[j, i, k] + U[j, i + 1, k])
                vU.append(w[j, i, k])
                Ui = 0.5 * (UU_0 + UU_1)
                up.append(0.5 * (Ui + abs(Ui))); um.append(0.5 * (Ui - abs(Ui)))
                UU_0 = UU_1
            qp, qm = _py_interpolate(np.array(vU), order, False)
            for k in range(1, l + 1):
                if order % 2 == 0:
                    ref[j, i, k - 1] = ref[j, i, k - 1] - qm[k]
                else:
                    ref[j, i, k - 1] = ref[j, i, k - 1] - qp[k - 1] * up[k - 1] - qm[k] * um[k - 1]
    got = r0.copy()
    K.vortex_force_direc(U, w, got, order, linear=True)
    assert np.array_equal(got, ref)
    ref = r0.copy()
    for j in range(m):
        for k in range(l - 1):
            UU_0, vU, up, um = 0.0, [], [], []
            for i in range(n):
                UU_1 = 0.5 * (U[j, i, k] + U[j, i, k + 1])
                vU.append(w[j, i, k])
                Ui = 0.5 * (UU_0 + UU_1)
                up.append(0.5 * (Ui + abs(Ui))); um.append(0.5 * (Ui - abs(Ui)))
                UU_0 = UU_1
            qp, qm = _py_interpolate(np.array(vU), order, False)
            for i in range(1, n + 1):
                if order % 2 == 0:
                    ref[j, i - 1, k] = ref[j, i - 1, k] + qm[i]
                else:
                    ref[j, i - 1, k] = ref[j, i - 1, k] + qp[i - 1] * up[i - 1] + qm[i] * um[i - 1]
    got = r0.copy()
    K.vortex_force_flip(U, w, got, order, linear=True)
    assert np.array_equal(got, ref)
