"""GPU parity, multigrid: hierarchy, operator coefficients, every level operation and complete
solves of libnyles_b200's MG against the CPU restatement of mgfor (oracle/csrc/oracle_mg.c).

Stencil operations are bit-exact (same order, no FMA).  Only the two norms are summed in a
different order on the GPU; they feed the stopping test, so residual histories agree to 1e-12
relative and iteration counts must be identical."""
import numpy as np
import pytest
import torch

from oracle.kernels import OracleMG

pytestmark = pytest.mark.gpu

# (nx, ny, nz, topology)   topology: 1 closed, 5 perio_xy, 6 perio_xyz (mg_enums.f90:5-7)
GRIDS = [(8, 8, 8, 1), (16, 8, 8, 1), (32, 16, 16, 1), (8, 8, 16, 5), (16, 16, 8, 5), (8, 8, 8, 6),
         (16, 16, 16, 6), (32, 32, 16, 6), (4, 4, 4, 6), (16, 16, 8, 6)]
IVARS = dict(x=1, b=2, r=3, y=4, diag=5, idiag=6, msk=7, Rcoef=8, Pcoef=9)


def make(nx, ny, nz, topo):
    from nyles_b200.mgfordriver import MG
    return MG(1, 1, nx, ny, nz, 3, topo), OracleMG(1, 1, nx, ny, nz, 3, topo)


def host(t):
    return t.cpu().numpy()


@pytest.mark.parametrize("grid", GRIDS)
def test_hierarchy_and_coefficients(grid):
    g, o = make(*grid)
    assert g.nlevels == o.nlevels
    for lev in range(1, g.nlevels + 1):
        assert g.get_arrayshape(lev) == o.get_arrayshape(lev)
        for name in ("msk", "diag", "idiag", "Rcoef", "Pcoef", "x", "b"):
            a = host(g.get_array(ivar=IVARS[name], lev=lev))
            b = o.get_array(ivar=IVARS[name], lev=lev)
            assert np.array_equal(a, b), "level %d %s" % (lev, name)


def test_box_domain_operator_identities():
    """diag = 6 in the interior, Rcoef = 0.5, Pcoef = 1/64 away from walls."""
    g, _ = make(32, 16, 16, 1)
    d = host(g.get_array(ivar=5, lev=1))
    assert d[8, 8, 8] == 6.0 and set(np.unique(d[d > 0])) == {3.0, 4.0, 5.0, 6.0}
    assert set(np.unique(host(g.get_array(ivar=8, lev=2)))) == {0.0, 0.5}
    P = host(g.get_array(ivar=9, lev=1))
    assert set(np.unique(1.0 / P[P > 0])) == {27.0, 36.0, 48.0, 64.0}


@pytest.mark.parametrize("grid", GRIDS)
def test_level_operations(grid):
    g, o = make(*grid)
    rng = np.random.default_rng(5)
    for lev in range(1, g.nlevels + 1):
        shape = o.get_arrayshape(lev)
        for name in ("x", "b", "r"):
            a = rng.standard_normal(shape)
            g.set_array(a, ivar=IVARS[name], lev=lev)
            o.set_array(a, ivar=IVARS[name], lev=lev)
    for lev in range(1, g.nlevels + 1):
        ops = ["smooth", "residual"]
        if lev < g.nlevels:
            ops += ["restriction", "prolongation"]
        for name in ops:
            g.op(name, lev)
            o.op(name, lev)
            touched = {"smooth": [(lev, "x")], "residual": [(lev, "r")],
                       "restriction": [(lev + 1, "b"), (lev + 1, "x")], "prolongation": [(lev, "x")]}[name]
            for l2, var in touched:
                a = host(g.get_array(ivar=IVARS[var], lev=l2))
                b = o.get_array(ivar=IVARS[var], lev=l2)
                assert np.array_equal(a, b), "%s at level %d changed %s differently" % (name, lev, var)


def point_sources(shape, seed=0):
    """Two opposite point sources, in the spirit of mgfor/tests.f90:50-57."""
    b = np.zeros(shape)
    nz, ny, nx = shape
    b[3 + (nz - 6) // 4, 3 + (ny - 6) // 4, 3 + (nx - 6) // 4] = -1.0
    b[3 + 3 * (nz - 6) // 4, 3 + 3 * (ny - 6) // 4, 3 + 3 * (nx - 6) // 4] = 1.0
    return b


@pytest.mark.parametrize("grid", GRIDS)
def test_solve_point_sources_and_warm_start(grid):
    g, o = make(*grid)
    shape = o.get_arrayshape(1)
    rng = np.random.default_rng(9)
    for it in range(3):                      # the second and third solves warm-start from x (solvers.f90)
        b = point_sources(shape) if it == 0 else point_sources(shape) + 0.01 * _interior_noise(rng, shape, grid[3])
        xg = torch.zeros(shape, dtype=torch.float64, device="cuda")
        xo = np.zeros(shape)
        g.solve(xg, b)
        o.solve(xo, b)
        assert g.stats["nite"] == o.nite, "V-cycle count differs (%d vs %d)" % (g.stats["nite"], o.nite)
        np.testing.assert_allclose(g.stats["res"], o.reshist, rtol=1e-11, atol=0)
        assert np.array_equal(host(xg), xo), "solution differs after solve %d" % it


def _interior_noise(rng, shape, topo):
    a = np.zeros(shape)
    a[3:-3, 3:-3, 3:-3] = rng.standard_normal(tuple(s - 6 for s in shape))
    return a


def test_zero_rhs_returns_immediately():
    g, o = make(16, 16, 16, 1)
    shape = o.get_arrayshape(1)
    x = torch.zeros(shape, dtype=torch.float64, device="cuda")
    g.solve(x, np.zeros(shape))
    assert g.stats["nite"] == 0 and g.stats["res"] == [0.0] and float(x.abs().max()) == 0.0


def test_grid_that_cannot_be_coarsened_is_refused():
    from nyles_b200.mgfordriver import MG
    from nyles_b200.lib import NylesB200Error
    with pytest.raises(NylesB200Error):
        MG(1, 1, 24, 24, 24, 3, 1)           # 24 -> 12 -> 6 -> 3: never reaches 2 (mg_setup.f90:262)


def test_residual_reduction_property_large():
    """Size-independent property at a size the oracle is not run at: each V-cycle contracts the
    residual and the solve reaches the reference's tolerance within maxite."""
    g, _ = make(128, 128, 128, 1)
    shape = g.get_arrayshape(1)
    b = torch.zeros(shape, dtype=torch.float64, device="cuda")
    gen = torch.Generator(device="cuda").manual_seed(3)
    inner = torch.randn((128, 128, 128), dtype=torch.float64, device="cuda", generator=gen)
    inner -= inner.mean()                    # compatibility condition of the Neumann problem
    b[3:-3, 3:-3, 3:-3] = inner
    x = torch.zeros_like(b)
    g.solve(x, b)
    res = g.stats["res"]
    assert all(r1 < r0 for r0, r1 in zip(res, res[1:]))
    assert g.stats["final_res"] < 1e-6 or g.stats["nite"] == 20


@pytest.mark.parametrize("grid", [(64, 64, 64, 1), (128, 64, 32, 1), (64, 32, 64, 5), (64, 64, 64, 6), (32, 64, 128, 6)])
def test_box_kernels_equal_generic_kernels(grid):
    """The fused analytic-coefficient kernels (hot path) against the one-kernel-per-Fortran-loop path
    that reads the coefficient arrays, at sizes the CPU oracle is not run at: every level operation
    and a complete solve must agree bit for bit (norms: same summation order is not required)."""
    from nyles_b200.mgfordriver import MG
    nx, ny, nz, topo = grid
    fast, slow = MG(1, 1, nx, ny, nz, 3, topo), MG(1, 1, nx, ny, nz, 3, topo)
    assert fast.is_box()
    slow.set_fast_path(False)
    assert not slow.is_box()
    gen = torch.Generator(device="cuda").manual_seed(11)
    for lev in range(1, fast.nlevels + 1):
        shape = fast.get_arrayshape(lev)
        for name in ("x", "b", "r"):
            a = torch.randn(shape, dtype=torch.float64, device="cuda", generator=gen)
            fast.set_array(a, ivar=IVARS[name], lev=lev)
            slow.set_array(a, ivar=IVARS[name], lev=lev)
    for lev in range(1, fast.nlevels + 1):
        ops = ["smooth", "residual"] + (["restriction", "prolongation"] if lev < fast.nlevels else [])
        for name in ops:
            fast.op(name, lev)
            slow.op(name, lev)
            touched = {"smooth": [(lev, "x")], "residual": [(lev, "r")],
                       "restriction": [(lev + 1, "b"), (lev + 1, "x")], "prolongation": [(lev, "x")]}[name]
            for l2, var in touched:
                a, b = fast.get_array(ivar=IVARS[var], lev=l2), slow.get_array(ivar=IVARS[var], lev=l2)
                assert torch.equal(a, b), "%s at level %d: %s differs between box and generic kernels" % (name, lev, var)
    for m in (fast, slow):                    # the fused V-cycle legs rely on consistent halos of x and b
        m.op("fill", 1)
    fast.op("vcycle", 1)
    slow.op("vcycle", 1)
    assert torch.equal(fast.get_array(ivar=1), slow.get_array(ivar=1))
    shape = fast.get_arrayshape(1)
    b = torch.zeros(shape, dtype=torch.float64, device="cuda")
    inner = torch.randn((nz, ny, nx), dtype=torch.float64, device="cuda", generator=gen)
    b[3:-3, 3:-3, 3:-3] = inner - inner.mean()
    xf, xs = torch.zeros_like(b), torch.zeros_like(b)
    fast.solve(xf, b, fill_halo=True)
    slow.solve(xs, b, fill_halo=True)
    assert fast.stats["nite"] == slow.stats["nite"]
    np.testing.assert_allclose(fast.stats["res"], slow.stats["res"], rtol=1e-11, atol=0)
    assert torch.equal(xf, xs)


@pytest.mark.parametrize("split,wide", [(1, 300000), (2, 0), (1, 0), (2, 1 << 40)])
@pytest.mark.parametrize("grid", [(128, 64, 96, 1), (64, 128, 32, 1), (96, 32, 64, 1), (64, 64, 64, 5), (128, 32, 64, 6),
                                  (16, 16, 16, 6), (8, 8, 8, 1), (256, 128, 128, 1), (128, 256, 128, 5), (128, 128, 256, 6)])
def test_fused_legs_equal_single_operator_kernels(grid, split, wide):
    """The TMA-staged fused V-cycle legs (smooth+residual+restriction, prolongation+smooth[+norm]) and the one-launch
    tail of the V-cycle against the one-box-kernel-per-operator V-cycle, bit for bit, on grids whose tiles are ragged
    in every direction: single V-cycles from random x (halos included) and b, then complete solves with warm starts.
    wide: levels of at most that many cells run inside the tail launch as its grid-barrier ("wide") levels -- the
    default, none (every level above 4096 cells through the fused legs), or every level of every grid."""
    from nyles_b200.mgfordriver import MG
    from nyles_b200 import lib
    nx, ny, nz, topo = grid
    # test grids are small: split wherever a wall-free tile exists (a default that multigrids copy at creation)
    lib.load().ny_mg_set_split_tiles(1)
    lib.load().ny_mg_set_wide_cells(wide)
    try:
        fused, plain = MG(1, 1, nx, ny, nz, 3, topo), MG(1, 1, nx, ny, nz, 3, topo)
    finally:
        lib.load().ny_mg_set_split_tiles(148)
        lib.load().ny_mg_set_wide_cells(300000)
    plain.set_fused_legs(False)
    # split 1 (default): tiles away from the x / y walls run the specialised (wall-free) kernel instance,
    # the frame around them the general one; split 2: every tile through the general instance
    fused.set_fused_legs(split)
    gen = torch.Generator(device="cuda").manual_seed(21)
    shape = fused.get_arrayshape(1)
    for rep in range(2):
        x = torch.randn(shape, dtype=torch.float64, device="cuda", generator=gen)
        b = torch.randn(shape, dtype=torch.float64, device="cuda", generator=gen)
        for m in (fused, plain):
            m.set_array(x, ivar=1)
            m.set_array(b, ivar=2)
            m.op("fill", 1)
            m.op("vcycle", 1)
            m.op("vcycle", 1)
        for lev in range(1, fused.nlevels + 1):
            for ivar in (1, 2):
                assert torch.equal(fused.get_array(ivar=ivar, lev=lev), plain.get_array(ivar=ivar, lev=lev)), \
                    "level %d ivar %d after two V-cycles" % (lev, ivar)
    for rep in range(3):
        b = torch.zeros(shape, dtype=torch.float64, device="cuda")
        inner = torch.randn((nz, ny, nx), dtype=torch.float64, device="cuda", generator=gen)
        b[3:-3, 3:-3, 3:-3] = inner - inner.mean()
        xf, xp = torch.zeros_like(b), torch.zeros_like(b)
        fused.solve(xf, b, fill_halo=True)
        plain.solve(xp, b, fill_halo=True)
        assert fused.stats["nite"] == plain.stats["nite"] and fused.stats["nite"] > 0
        np.testing.assert_allclose(fused.stats["res"], plain.stats["res"], rtol=1e-11, atol=0)
        assert torch.equal(xf, xp)


@pytest.mark.parametrize("grid", [(32, 16, 16, 1), (16, 16, 16, 6), (16, 32, 8, 5)])
def test_obstacle_mask_rebuilds_the_operators(grid):
    """a user mask at level 1 (set_pyarray ivar=7, mgfor/tests.f90:207-212) followed by
    setup_fine_msk + setup_operators.  Coarse masks and every coefficient array must equal the oracle's bit for
    bit, the analytic box kernels must step aside, and V-cycles / solves on the masked domain must match."""
    nx, ny, nz, topo = grid
    g, o = make(*grid)
    assert g.is_box()
    msk = o.get_array(ivar=7)
    msk[3 + nz // 4:3 + nz // 2, 3 + ny // 4:3 + ny // 2 + 1, 3 + nx // 2:3 + nx // 2 + 5] = 0.0     # a block
    msk[3:3 + 2, 3 + ny - 3:3 + ny, 3:3 + 3] = 0.0                                                   # a corner step
    g.set_mask(msk)
    o.set_mask(msk)
    assert not g.is_box()
    for lev in range(1, g.nlevels + 1):
        for name in ("msk", "diag", "idiag", "Rcoef", "Pcoef"):
            assert np.array_equal(host(g.get_array(ivar=IVARS[name], lev=lev)), o.get_array(ivar=IVARS[name], lev=lev)), \
                "level %d %s" % (lev, name)
    assert host(g.get_array(ivar=5, lev=1))[3 + nz // 4, 3 + ny // 4, 3 + nx // 2] == 0.0       # solid: no equation
    rng = np.random.default_rng(8)
    shape = g.get_arrayshape(1)
    fluid = o.get_array(ivar=7)[3:-3, 3:-3, 3:-3] > 0
    for rep in range(2):
        b = np.zeros(shape)
        inner = rng.standard_normal((nz, ny, nx)) * fluid
        inner[fluid] -= inner[fluid].mean()                       # compatible right-hand side on the fluid cells
        b[3:-3, 3:-3, 3:-3] = inner
        xg, xo = np.zeros(shape), np.zeros(shape)
        g.solve(xg, b)                  # both sides take b as given (halo planes zero), like the solve test above
        o.solve(xo, b)
        assert g.stats["nite"] == o.nite and o.nite >= 1
        np.testing.assert_allclose(g.stats["res"], o.reshist, rtol=1e-10, atol=0)
        assert np.array_equal(xg, xo)
        assert np.all(xg[3:-3, 3:-3, 3:-3][~fluid] == 0.0)        # nothing is ever written inside the obstacle


def test_c_abi_example_runs(tmp_path):
    """examples/c_abi_example.c -- plain C on the C ABI, no Python in the process -- solves the point-source problem
    and reports the same V-cycle count and residual as the oracle."""
    import re
    import subprocess
    from test_host_logic import build_c_example
    exe = str(tmp_path / "c_abi_example")
    r = build_c_example(exe)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    m = re.search(r"V-cycles (\d+), \|\|r\|\|\^2/\|\|b\|\|\^2 = (\S+),", out.stdout)
    assert m, out.stdout
    n = 64
    o = OracleMG(1, 1, n, n, n, 3, 1)
    b = np.zeros(o.shape)
    b[3 + n // 4, 3 + n // 4, 3 + n // 4] = 1.0
    b[3 + 3 * n // 4, 3 + 3 * n // 4, 3 + 3 * n // 4] = -1.0
    x = np.zeros(o.shape)
    o.solve(x, b)
    assert int(m.group(1)) == o.nite
    assert abs(float(m.group(2)) - o.res) <= 1e-3 * o.res        # printed with 4 significant digits
