"""GPU parity, operator level: every entry point of libnyles_b200.so against the CPU oracle on the
same seeded inputs.  The arithmetic is fp64 in the Fortran's operation order without FMA, so the
bar is BIT-EXACT (np.array_equal; +0 == -0), far inside north_star's 1e-12."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import model as M
from oracle.kernels import Kernels

pytestmark = pytest.mark.gpu

SHAPES = [(5, 5, 5), (6, 7, 9), (8, 8, 16), (13, 10, 37), (32, 24, 40)]     # (nz, ny, nx); ragged on purpose


@pytest.fixture(scope="module")
def K():
    return Kernels("strict")


@pytest.fixture(scope="module")
def L():
    from nyles_b200 import lib
    return lib


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64).cuda()


def host(t):
    return t.cpu().numpy()


def views(a):
    """(k,j,i) array -> the reference's three orientations as strided views."""
    return {"i": a, "j": a.transpose(2, 0, 1), "k": a.transpose(1, 2, 0)}


def flip(a, d):
    return views(a)[{"i": "j", "j": "k", "k": "i"}[d]]


def rand_fields(shape, n, seed, scale=1.0):
    rng = np.random.default_rng(seed)
    return [scale * rng.standard_normal(shape) for _ in range(n)]


# --------------------------------------------------------------------------- oracle drivers
def oracle_vorticity(K, u, fparam):
    w = [np.full_like(u[0], 7.0) for _ in range(3)]       # sentinel: untouched rows must stay
    perm = {"i": ("k", "j"), "j": ("i", "k"), "k": ("j", "i")}
    comp = dict(zip("ijk", range(3)))
    for dirk in "ijk":
        dirj, diri = perm[dirk]
        wk = flip(w[comp[dirk]], dirk)
        K.vorticity(flip(u[comp[diri]], dirk), flip(u[comp[dirj]], dirk), wk)
        if fparam > 0 and dirk == "k":
            wk[:, :-1, :-1] += fparam
    return w


def oracle_vortex_force(K, U, w, du):
    comp = dict(zip("ijk", range(3)))
    for k, j, i in ["ikj", "jik", "kji"]:
        K.vortex_force_direc(flip(U[comp[k]], j), flip(w[comp[j]], j), flip(du[comp[i]], j))
        K.vortex_force_flip(flip(U[comp[i]], j), flip(w[comp[j]], j), flip(du[comp[k]], j))


def oracle_upwind(K, trac, U, coefs=None):
    d = np.full_like(trac, 3.0)
    for n, ax in enumerate("ijk"):
        dv = views(d)[ax]
        if ax == "i":
            dv[...] = 0.0
        K.upwind(views(trac)[ax], views(U[n])[ax], dv)
        if coefs is not None:
            K.add_laplacian(views(trac)[ax], dv, coefs[n])
    return d


# --------------------------------------------------------------------------- tests
@pytest.fixture(params=[1, 2], ids=["cellwise", "tma_marching"])
def mom_variant(request, L):
    """Which kernel evaluates the interior cells of the fused momentum right-hand side: k_momentum (one thread per
    cell) or k_mom3 (plane marching, TMA-staged tiles; needs an even nx and is otherwise only chosen for large
    grids).  Both must be bit-identical."""
    L.check(L.load().ny_set_momentum_variant(L.context(), request.param))
    yield request.param
    L.check(L.load().ny_set_momentum_variant(L.context(), 0))


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("fparam", [0.0, 0.37])
def test_vorticity(K, L, shape, fparam):
    u = rand_fields(shape, 3, 1)
    ref = oracle_vorticity(K, u, fparam)
    du = [dev(a) for a in u]
    w = [torch.full(shape, 7.0, dtype=torch.float64, device="cuda") for _ in range(3)]
    L.check(L.load().ny_vorticity(L.context(), *[L.ptr(t) for t in du], *[L.ptr(t) for t in w],
                                  L.ext(du[0]), fparam, L.stream()))
    for a, b in zip(ref, w):
        assert np.array_equal(a, host(b))


@pytest.mark.parametrize("shape", SHAPES)
def test_kin_div_gradp_scale(K, L, shape):
    u = rand_fields(shape, 3, 2)
    ids2 = (3.7, 0.9, 11.3)
    # kinetic energy
    ke = np.full(shape, 5.0)
    for n, ax in enumerate("ijk"):
        kv = views(ke)[ax]
        if ax == "i":
            kv[...] = 0.0
        K.kin(views(u[n])[ax], views(u[n])[ax], kv, ids2[n])
    du = [dev(a) for a in u]
    gke = torch.empty(shape, dtype=torch.float64, device="cuda")
    L.check(L.load().ny_kin(L.context(), *[L.ptr(t) for t in du], L.ptr(gke), *ids2, L.ext(gke), L.stream()))
    assert np.array_equal(ke, host(gke))
    # U = u * ids2
    gU = [torch.empty_like(t) for t in du]
    L.check(L.load().ny_U_from_u(L.context(), *[L.ptr(t) for t in du], *[L.ptr(t) for t in gU], *ids2,
                                 L.ext(gke), L.stream()))
    U = [u[n] * ids2[n] for n in range(3)]
    for a, b in zip(U, gU):
        assert np.array_equal(a, host(b))
    # divergence
    div = np.full(shape, 9.0)
    for n, ax in enumerate("ijk"):
        K.div(views(div)[ax], views(U[n])[ax], n)
    gdiv = torch.empty(shape, dtype=torch.float64, device="cuda")
    L.check(L.load().ny_div(L.context(), *[L.ptr(t) for t in gU], L.ptr(gdiv), L.ext(gdiv), L.stream()))
    assert np.array_equal(div, host(gdiv))
    # u -= grad p
    p = rand_fields(shape, 1, 3)[0]
    uu = [a.copy() for a in u]
    for n, ax in enumerate("ijk"):
        K.gradke(views(p)[ax], views(uu[n])[ax])
    gp = dev(p)
    L.check(L.load().ny_gradp(L.context(), L.ptr(gp), *[L.ptr(t) for t in du], L.ext(gp), L.stream()))
    for a, b in zip(uu, du):
        assert np.array_equal(a, host(b))


@pytest.mark.parametrize("shape", SHAPES + [(40, 37, 70), (9, 70, 34), (37, 19, 66)])
@pytest.mark.parametrize("seed", [10, 11])
def test_upwind(K, L, shape, seed):
    trac, Ux, Uy, Uz = rand_fields(shape, 4, seed)
    ref = oracle_upwind(K, trac, [Ux, Uy, Uz])
    g = [dev(a) for a in (trac, Ux, Uy, Uz)]
    out = torch.full(shape, 3.0, dtype=torch.float64, device="cuda")
    L.check(L.load().ny_upwind(L.context(), *[L.ptr(t) for t in g], L.ptr(out), L.ext(out), L.stream()))
    assert np.array_equal(ref, host(out))
    # with interleaved diffusion (tracer.py:72-77, last=True)
    coefs = (0.3, 0.11, 0.7)
    ref = oracle_upwind(K, trac, [Ux, Uy, Uz], coefs)
    L.check(L.load().ny_upwind_diff(L.context(), *[L.ptr(t) for t in g], L.ptr(out), *coefs, L.ext(out), L.stream()))
    assert np.array_equal(ref, host(out))


def test_upwind_smooth_and_zero_velocity(K, L):
    """sign test is strictly u > 0 (weno.f90:114): zero and negative-zero velocities take the else branch."""
    shape = (8, 9, 12)
    z, y, x = np.meshgrid(*[np.linspace(0, 1, n) for n in shape], indexing="ij")
    trac = np.sin(2 * np.pi * x) * np.cos(2 * np.pi * y) + z
    U = [np.zeros(shape), -np.zeros(shape), np.where(x > 0.5, 1.0, -1.0) * 0.3]
    ref = oracle_upwind(K, trac, U)
    g = [dev(a) for a in [trac] + U]
    out = torch.empty(shape, dtype=torch.float64, device="cuda")
    L.check(L.load().ny_upwind(L.context(), *[L.ptr(t) for t in g], L.ptr(out), L.ext(out), L.stream()))
    assert np.array_equal(ref, host(out))


@pytest.mark.parametrize("shape", SHAPES)
def test_vortex_force_and_bernoulli(K, L, shape):
    U = rand_fields(shape, 3, 20)
    w = rand_fields(shape, 3, 21)
    du0 = rand_fields(shape, 3, 22)
    ke, b = rand_fields(shape, 2, 23)
    dz = 0.125
    ref = [a.copy() for a in du0]
    oracle_vortex_force(K, U, w, ref)
    gU, gw, gdu = [dev(a) for a in U], [dev(a) for a in w], [dev(a) for a in du0]
    L.check(L.load().ny_vortex_force(L.context(), *[L.ptr(t) for t in gU + gw + gdu], L.ext(gU[0]), L.stream()))
    for a, t in zip(ref, gdu):
        assert np.array_equal(a, host(t))
    # bernoulli on top (LES and Euler flavours)
    for euler in (0, 1):
        r2 = [a.copy() for a in ref]
        for n, ax in enumerate("ijk"):
            if ax == "k" and not euler:
                K.gradkeandb(views(ke)[ax], views(b)[ax], views(r2[n])[ax], dz)
            else:
                K.gradke(views(ke)[ax], views(r2[n])[ax])
        g2 = [dev(a) for a in ref]
        gke, gb = dev(ke), dev(b)           # keep alive: a freed temporary's block is reused at once
        L.check(L.load().ny_bernoulli(L.context(), L.ptr(gke), L.ptr(gb), *[L.ptr(t) for t in g2],
                                      dz, euler, L.ext(g2[0]), L.stream()))
        for a, t in zip(r2, g2):
            assert np.array_equal(a, host(t))


@pytest.mark.parametrize("shape", SHAPES[1:] + [(40, 37, 70), (9, 70, 34)])
@pytest.mark.parametrize("flags", [0, 1, 2])
def test_fused_rhs_equals_operator_sequence(K, L, shape, flags, mom_variant):
    euler, linear = flags & 1, flags & 2
    b, ke = rand_fields(shape, 2, 30)
    U = rand_fields(shape, 3, 31)
    w = rand_fields(shape, 3, 32)
    dz = 0.25
    db = oracle_upwind(K, b, U)
    du = [np.zeros(shape) for _ in range(3)]
    if not linear:
        oracle_vortex_force(K, U, w, du)
    for n, ax in enumerate("ijk"):
        if ax == "k" and not euler:
            K.gradkeandb(views(ke)[ax], views(b)[ax], views(du[n])[ax], dz)
        else:
            K.gradke(views(ke)[ax], views(du[n])[ax])
    gb, gke, gU, gw = dev(b), dev(ke), [dev(a) for a in U], [dev(a) for a in w]
    gdb = torch.full(shape, 4.0, dtype=torch.float64, device="cuda")
    gdu = [torch.full(shape, 4.0, dtype=torch.float64, device="cuda") for _ in range(3)]
    L.check(L.load().ny_rhs(L.context(), L.ptr(gb), *[L.ptr(t) for t in gU + gw], L.ptr(gke), L.ptr(gdb),
                            *[L.ptr(t) for t in gdu], dz, flags, L.ext(gb), L.stream()))
    if not euler:
        assert np.array_equal(db, host(gdb))
    for a, t in zip(du, gdu):
        assert np.array_equal(a, host(t))


@pytest.mark.parametrize("shape", SHAPES[1:] + [(40, 37, 70)])
@pytest.mark.parametrize("mode", [1, 2, 3])
@pytest.mark.parametrize("flags", [0, 1, 2])
@pytest.mark.parametrize("forced", [False, True])
def test_rhs_step_equals_rhs_then_timescheme(L, shape, mode, flags, forced, mom_variant):
    """ny_rhs_step (the RHS kernels apply the LFAM3 / Euler update and write each field once, into a
    separate buffer) against ny_rhs followed by the ny_ts_* kernels: bit-identical new state, and none
    of the arrays it only reads is touched.  forced: user tendencies (what a forcing object adds to dstate after
    the RHS, core/model_les.py:143-144) for b and two of the velocity components."""
    euler = flags & 1
    if forced and shape != SHAPES[2]:
        pytest.skip("user tendencies are exercised on one shape")
    b, ke, sb_b, sn_b = rand_fields(shape, 4, 40)
    U, w, u, ub, un = (rand_fields(shape, 3, 41 + n) for n in range(5))
    dz, dt = 0.25, 0.0137
    lib = L.load()
    g = lambda a: dev(a)                                     # noqa: E731
    gU, gw, gke = [g(a) for a in U], [g(a) for a in w], g(ke)
    # reference sequence: tendencies, then the elementwise update of every field
    s = [g(b)] + [g(a) for a in u]
    sb = [g(sb_b)] + [g(a) for a in ub]
    sn = [g(sn_b)] + [g(a) for a in un]
    ds = [torch.zeros(shape, dtype=torch.float64, device="cuda") for _ in range(4)]
    L.check(lib.ny_rhs(L.context(), L.ptr(s[0]), *[L.ptr(t) for t in gU + gw], L.ptr(gke), L.ptr(ds[0]),
                       *[L.ptr(t) for t in ds[1:]], dz, flags, L.ext(s[0]), L.stream()))
    fields = range(1, 4) if euler else range(4)
    addp = None
    if forced:
        Q = [g(a) for a in rand_fields(shape, 4, 77)]
        Q[2] = None                                          # a field without user tendency
        for f in fields:
            if Q[f] is not None:
                ds[f] += Q[f]                                # forcing.add(state, dstate, t)
        addp = C.byref((C.c_void_p * 4)(*[None if (q is None or (euler and n == 0)) else L.ptr(q).value
                                          for n, q in enumerate(Q)]))
    for f in fields:
        n = s[f].numel()
        if mode == 1:
            L.check(lib.ny_ts_lfam3_first(L.context(), L.ptr(s[f]), L.ptr(ds[f]), L.ptr(sb[f]), L.ptr(sn[f]), dt, n, L.stream()))
        elif mode == 2:
            L.check(lib.ny_ts_lfam3_pred(L.context(), L.ptr(s[f]), L.ptr(ds[f]), L.ptr(sb[f]), L.ptr(sn[f]), dt, n, L.stream()))
        else:
            L.check(lib.ny_ts_lfam3_corr(L.context(), L.ptr(s[f]), L.ptr(ds[f]), L.ptr(sn[f]), dt, n, L.stream()))
    # fused, rotating form
    s2 = [g(b)] + [g(a) for a in u]
    sb2 = [g(sb_b)] + [g(a) for a in ub]
    sn2 = [g(sn_b)] + [g(a) for a in un]
    out = [torch.full(shape, 9.0, dtype=torch.float64, device="cuda") for _ in range(4)]

    def ptr4(ts):
        return C.byref((C.c_void_p * 4)(*[None if (euler and n == 0) else L.ptr(t).value for n, t in enumerate(ts)]))
    L.check(lib.ny_rhs_step(L.context(), *[L.ptr(t) for t in gU + gw], L.ptr(gke), ptr4(s2), ptr4(sb2), ptr4(sn2),
                            ptr4(out), addp, mode, dt, dz, flags, L.ext(s2[0]), L.stream()))
    for f in fields:
        assert np.array_equal(host(out[f]), host(s[f])), "field %d" % f
        # read-only inputs stay as they were
        assert np.array_equal(host(s2[f]), (b if f == 0 else u[f - 1]))
        assert np.array_equal(host(sb2[f]), (sb_b if f == 0 else ub[f - 1]))
        assert np.array_equal(host(sn2[f]), (sn_b if f == 0 else un[f - 1]))
    # what the rotation of the caller relies on: after modes 1 and 2 the old state array equals sn (and sb)
    if mode in (1, 2):
        for f in fields:
            assert np.array_equal(host(sn[f]), host(s2[f])) and np.array_equal(host(sb[f]), host(s2[f]))
    # aliasing the output with an array the launch reads is refused
    assert lib.ny_rhs_step(L.context(), *[L.ptr(t) for t in gU + gw], L.ptr(gke), ptr4(s2), ptr4(sb2), ptr4(sn2),
                           ptr4(s2 if mode != 3 else sn2), None, mode, dt, dz, flags, L.ext(s2[0]), L.stream()) != 0


@pytest.mark.parametrize("shape", SHAPES[:3])
def test_add_laplacian(K, L, shape):
    phi, dphi = rand_fields(shape, 2, 40)
    coefs = (0.2, 0.5, 0.9)
    ref = dphi.copy()
    for n, ax in enumerate("ijk"):
        K.add_laplacian(views(phi)[ax], views(ref)[ax], coefs[n])
    g = dev(dphi)
    gphi = dev(phi)
    L.check(L.load().ny_add_laplacian(L.context(), L.ptr(gphi), L.ptr(g), *coefs, L.ext(g), L.stream()))
    assert np.array_equal(ref, host(g))


def test_extent_guard(L):
    """flux1d reads out of bounds below 5 cells (weno.f90:106-153): the library refuses instead."""
    t = torch.zeros((4, 8, 8), dtype=torch.float64, device="cuda")
    rc = L.load().ny_upwind(L.context(), L.ptr(t), L.ptr(t), L.ptr(t), L.ptr(t), L.ptr(t), L.ext(t), L.stream())
    assert rc == -1 and b"extent" in L.load().ny_last_error()


def test_timescheme_kernels(L):
    n = 100003
    rng = np.random.default_rng(50)
    s, ds, sb, d1, d2 = [rng.standard_normal(n) for _ in range(5)]
    dt = 0.0371
    ctx, lib = L.context(), L.load()

    def run(fn, arrays, *scalars):
        g = [dev(a) for a in arrays]
        L.check(fn(ctx, *[L.ptr(t) for t in g], *scalars, n, L.stream()))
        return [host(t) for t in g]

    out = run(lib.ny_ts_axpy, [s, ds], dt)
    r = s.copy(); r += dt * ds
    assert np.array_equal(out[0], r)
    out = run(lib.ny_ts_lfam3_first, [s, ds, sb, d1], dt)
    r = s.copy(); r += dt * ds
    assert np.array_equal(out[0], r) and np.array_equal(out[2], s) and np.array_equal(out[3], s)
    out = run(lib.ny_ts_lfam3_pred, [s, ds, sb, d1], dt)
    r = sb + (2. * dt) * ds
    r = (1. / 12.) * (5. * r + 8. * s - sb)
    assert np.array_equal(out[0], r) and np.array_equal(out[2], s) and np.array_equal(out[3], s)
    out = run(lib.ny_ts_lfam3_corr, [s, ds, sb], dt)
    assert np.array_equal(out[0], sb + dt * ds)
    out = run(lib.ny_ts_rk3_stage2, [s, ds, d1], dt)
    r = s.copy(); r += (dt / 4.) * (d1 - 3 * ds)
    assert np.array_equal(out[0], r)
    out = run(lib.ny_ts_rk3_stage3, [s, ds, d1, d2], dt)
    r = s.copy(); r += (dt / 12.) * (8 * d2 - ds - d1)
    assert np.array_equal(out[0], r)


def test_max_speed2(L):
    n = 300007
    rng = np.random.default_rng(60)
    U, V, W = [rng.standard_normal(n) for _ in range(3)]
    out = C.c_double()
    g = [dev(U), dev(V), dev(W)]
    L.check(L.load().ny_max_speed2(L.context(), *[L.ptr(t) for t in g], n, C.byref(out), L.stream()))
    assert out.value == np.max(U ** 2 + V ** 2 + W ** 2)
    U[1234] = np.nan
    g = [dev(U), dev(V), dev(W)]
    L.check(L.load().ny_max_speed2(L.context(), *[L.ptr(t) for t in g], n, C.byref(out), L.stream()))
    assert np.isnan(out.value)


@pytest.mark.parametrize("geometry", ["closed", "perio_x", "perio_y", "perio_xy", "perio_xyz"])
def test_halo_fill_self(L, geometry):
    from nyles_b200 import halo as H, variables as V
    p = M.make_param(9, 7, 8, geometry=geometry)
    p["device"] = "cuda"
    s = V.Scalar(p, "b", "b", "")
    oracle_s = M.Scalar(p, "b")
    rng = np.random.default_rng(70)
    oracle_s.data[...] = rng.standard_normal(oracle_s.data.shape)
    s.tensor.copy_(dev(oracle_s.data))
    M.Halo(p, oracle_s).fill(oracle_s)
    st = type("S", (), {"b": s})
    H.set_halo(p, st).fill(s)
    assert np.array_equal(oracle_s.data, host(s.tensor))


@pytest.mark.parametrize("shape", SHAPES + [(40, 37, 70)])
def test_fast_arithmetic_within_north_star_tolerance(K, L, shape):
    """ny_set_arith(1): the re-associated weno5 (one reciprocal, FMAs downstream of the exact tau5)
    must stay within the north-star bar of 1e-12 relative per RHS field; strict mode is restored."""
    trac, Ux, Uy, Uz = rand_fields(shape, 4, 70)
    w = rand_fields(shape, 3, 71)
    ke, b = rand_fields(shape, 2, 72)
    dz = 0.125
    ref_db = oracle_upwind(K, trac, [Ux, Uy, Uz])
    ref_du = [np.zeros(shape) for _ in range(3)]
    oracle_vortex_force(K, [Ux, Uy, Uz], w, ref_du)
    for n, ax in enumerate("ijk"):
        if ax == "k":
            K.gradkeandb(views(ke)[ax], views(trac)[ax], views(ref_du[n])[ax], dz)
        else:
            K.gradke(views(ke)[ax], views(ref_du[n])[ax])
    g = {k: dev(v) for k, v in dict(b=trac, Ux=Ux, Uy=Uy, Uz=Uz, wx=w[0], wy=w[1], wz=w[2], ke=ke).items()}
    out = [torch.empty(shape, dtype=torch.float64, device="cuda") for _ in range(4)]
    lib = L.load()
    try:
        L.check(lib.ny_set_arith(L.context(), 1))
        L.check(lib.ny_rhs(L.context(), L.ptr(g["b"]), L.ptr(g["Ux"]), L.ptr(g["Uy"]), L.ptr(g["Uz"]), L.ptr(g["wx"]),
                           L.ptr(g["wy"]), L.ptr(g["wz"]), L.ptr(g["ke"]), *[L.ptr(t) for t in out], dz, 0,
                           L.ext(out[0]), L.stream()))
        worst = 0.0
        for ref, t in zip([ref_db] + ref_du, out):
            err = np.max(np.abs(host(t) - ref)) / np.max(np.abs(ref))
            worst = max(worst, err)
            assert err <= 1e-12, "fast arithmetic off by %.3e relative" % err
        assert worst > 0.0 or min(shape) < 6        # it really is the other code path (not bit-identical)
    finally:
        L.check(lib.ny_set_arith(L.context(), 0))
    # and strict mode is bit-exact again
    L.check(lib.ny_upwind(L.context(), L.ptr(g["b"]), L.ptr(g["Ux"]), L.ptr(g["Uy"]), L.ptr(g["Uz"]), L.ptr(out[0]),
                          L.ext(out[0]), L.stream()))
    assert np.array_equal(ref_db, host(out[0]))


def test_in_range_division_is_the_ieee_quotient():
    """ny_weno.cuh div_inrange (the compiler's fast-path sequence without its range test and slow-path
    call) against the IEEE quotient, bitwise, over the operand ranges weno5 feeds it: tau5 = 0 or a
    REAL(4)-representable value down to the float denormals, beta + eps from 1e-16 up, and signed
    numerators over many decades."""
    from nyles_b200 import lib
    L = lib.load()
    ctx = lib.context()
    gen = torch.Generator(device="cuda").manual_seed(123)
    n = 1 << 22
    total = 0
    for case in range(6):
        mant_a = 1.0 + torch.rand(n, dtype=torch.float64, device="cuda", generator=gen)
        mant_b = 1.0 + torch.rand(n, dtype=torch.float64, device="cuda", generator=gen)
        ea = torch.randint(-160, 100, (n,), device="cuda", generator=gen).double()
        eb = torch.randint(-54, 200, (n,), device="cuda", generator=gen).double()
        a = mant_a * torch.exp2(ea)
        b = mant_b * torch.exp2(eb)
        if case == 1:
            a = a.float().double()                  # REAL(4) values, like tau5
        if case == 2:
            a = torch.zeros_like(a)
        if case == 3:
            a = -a
        if case == 4:                               # tau5 next to beta: quotients near 1
            b = mant_b * torch.exp2(torch.clamp(eb, max=100.0))     # keep the REAL(4) dividend finite
            a = (b * (1.0 + 1e-3 * torch.randn(n, dtype=torch.float64, device="cuda", generator=gen))).float().double()
        if case == 5:                               # float denormals as dividend
            a = (torch.randint(1, 1 << 20, (n,), device="cuda", generator=gen).double() * 2.0 ** -149)
        out = torch.empty_like(a)
        bad = C.c_longlong(-1)
        lib.check(L.ny_debug_div(ctx, lib.ptr(a), lib.ptr(b), lib.ptr(out), n, C.byref(bad), lib.stream()))
        assert bad.value == 0, "case %d: %d quotients differ from IEEE" % (case, bad.value)
        assert torch.equal(out, a / b)
        total += n
    assert total == 6 * n


@pytest.mark.parametrize("fast", [False, True])
def test_weno5_primitive_against_oracle(fast):
    """weno5 (core/weno.f90:25-54) on smooth, sharp, still (all-zero, constant) and tiny stencils:
    strict mode bitwise, fast mode within 1e-13 of the stencil scale."""
    from nyles_b200 import lib
    o = Kernels("strict")
    rng = np.random.default_rng(77)
    n = 20000
    q = rng.standard_normal((5, n))
    q[:, :2000] = 0.0                                        # still fluid
    q[:, 2000:4000] = 1.0                                    # constant
    q[:, 4000:6000] *= 1e-12
    q[:, 6000:8000] = 1.0 + 1e-16 * rng.integers(-3, 4, (5, 2000))   # round-off noise on a constant
    q[:, 8000:10000] = np.sign(q[:, 8000:10000])            # shocks
    q[:, 10000:12000] = 1e-7 * rng.standard_normal((5, 2000)) * 1e-16   # vorticity of a potential flow
    q[:, 12000:14000] *= 1e-150
    ref = np.array([o.weno5(*q[:, t]) for t in range(n)])
    L = lib.load()
    lib.set_arith(fast)
    try:
        qd = torch.as_tensor(q, device="cuda").contiguous()
        out = torch.empty(n, dtype=torch.float64, device="cuda")
        lib.check(L.ny_debug_weno5(lib.context(), lib.ptr(qd), lib.ptr(out), n, lib.stream()))
        got = out.cpu().numpy()
    finally:
        lib.set_arith(False)
    if not fast:
        assert np.array_equal(got, ref)
    else:
        scale = np.max(np.abs(q), axis=0) + 1e-300
        assert np.max(np.abs(got - ref) / scale) <= 1e-13


def test_weno3_primitive_against_oracle():
    """weno3 (core/weno.f90:1-22; the closure of flux1d next to the line ends) on smooth, sharp, still, constant,
    round-off-noise and tiny stencils, bitwise against the oracle (there is one arithmetic mode: the edge faces
    always run the source-order code)."""
    from nyles_b200 import lib
    o = Kernels("strict")
    rng = np.random.default_rng(78)
    n = 20000
    q = rng.standard_normal((3, n))
    q[:, :2000] = 0.0
    q[:, 2000:4000] = 1.0
    q[:, 4000:6000] *= 1e-12
    q[:, 6000:8000] = 1.0 + 1e-16 * rng.integers(-3, 4, (3, 2000))
    q[:, 8000:10000] = np.sign(q[:, 8000:10000])
    q[:, 10000:12000] *= 1e-23
    q[:, 12000:14000] *= 1e-150
    q[:, 14000:16000] *= 1e+100
    ref = np.array([o.weno3(*q[:, t]) for t in range(n)])
    L = lib.load()
    qd = torch.as_tensor(q, device="cuda").contiguous()
    out = torch.empty(n, dtype=torch.float64, device="cuda")
    lib.check(L.ny_debug_weno3(lib.context(), lib.ptr(qd), lib.ptr(out), n, lib.stream()))
    assert np.array_equal(out.cpu().numpy(), ref)


# ---------------------------------------------------------------- the linear (non-WENO) upwind branch
@pytest.mark.parametrize("shape", SHAPES + [(9, 70, 34)])
@pytest.mark.parametrize("order", [1, 2, 3, 4, 5])
def test_linear_upwind_branch(K, L, shape, order):
    """fortran_upwind.f90:33-64 with core/interpolate_tracer.f90 (dormant in the shipped reference, linear=.false.):
    ny_upwind_linear against the oracle over the three directions as tracer.py:44-72 drives them, bit-exact."""
    trac, Ux, Uy, Uz = rand_fields(shape, 4, 60 + order)
    ref = np.full_like(trac, 3.0)
    for n, ax in enumerate("ijk"):
        dv = views(ref)[ax]
        if ax == "i":
            dv[...] = 0.0
        K.upwind_linear(views(trac)[ax], views([Ux, Uy, Uz][n])[ax], dv, order)
    g = [dev(a) for a in (trac, Ux, Uy, Uz)]
    out = torch.full(shape, 3.0, dtype=torch.float64, device="cuda")
    L.check(L.load().ny_upwind_linear(L.context(), *[L.ptr(t) for t in g], L.ptr(out), order, L.ext(out), L.stream()))
    assert np.array_equal(ref, host(out))


@pytest.mark.parametrize("shape", SHAPES + [(9, 70, 34)])
@pytest.mark.parametrize("order", [1, 2, 3, 4, 5])
def test_linear_vortex_force_branch(K, L, shape, order):
    """fortran_vortex_force.f90:39-64,118-143 with core/interpolate.f90: ny_vortex_force_linear against the oracle
    driven through the three passes of vortex_force.py:69-81, bit-exact (even orders omit the velocity, as the
    Fortran does)."""
    U, w, du = (rand_fields(shape, 3, 70 + order + n) for n in range(3))
    ref = [a.copy() for a in du]
    comp = dict(zip("ijk", range(3)))
    for k, j, i in ["ikj", "jik", "kji"]:
        K.vortex_force_direc(flip(U[comp[k]], j), flip(w[comp[j]], j), flip(ref[comp[i]], j), order, linear=True)
        K.vortex_force_flip(flip(U[comp[i]], j), flip(w[comp[j]], j), flip(ref[comp[k]], j), order, linear=True)
    gU, gw, gdu = [dev(a) for a in U], [dev(a) for a in w], [dev(a) for a in du]
    L.check(L.load().ny_vortex_force_linear(L.context(), *[L.ptr(t) for t in gU + gw + gdu], order, L.ext(gdu[0]), L.stream()))
    for a, t in zip(ref, gdu):
        assert np.array_equal(a, host(t))
