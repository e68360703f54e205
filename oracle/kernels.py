"""ORACLE (test infrastructure, NOT product code): ctypes front-end of the C restatement.

Exposes the Fortran routines of the reference with their f2py call signatures
(``core/Makefile:1-2``; SURVEY.md 8b "Boundary 1"): float64 3-D NumPy arrays of any
strides, mutated in place, no return value.  Two builds exist (oracle/Makefile):
``strict`` (bit-reproducible, used by all parity tests) and ``fast`` (OpenMP,
-march=native; only timed as the CPU baseline by bench.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm
may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}

_PD = C.POINTER(C.c_double)
_PS = C.POINTER(C.c_ssize_t)


def build(flavour="strict", force=False):
    """Compile oracle/csrc with the recipe in oracle/Makefile (seconds)."""
    target = "liboracle_%s.so" % flavour
    path = os.path.join(_HERE, target)
    if force and os.path.exists(path):
        os.remove(path)
    if not os.path.exists(path):
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.check_call(["make", "-s", "-C", _HERE, target], env=env)
    return path


def lib(flavour="strict"):
    if flavour not in _LIBS:
        L = C.CDLL(build(flavour))
        L.orc_weno3.restype = C.c_double
        L.orc_weno3.argtypes = [C.c_double] * 3
        L.orc_weno5.restype = C.c_double
        L.orc_weno5.argtypes = [C.c_double] * 5
        L.orc_mg_create.restype = C.c_void_p
        L.orc_mg_create.argtypes = [C.c_int] * 4
        L.orc_mg_norm.restype = C.c_double
        _LIBS[flavour] = L
    return _LIBS[flavour]


def _common(arrays):
    a0 = arrays[0]
    assert a0.dtype == np.float64 and a0.ndim == 3
    for a in arrays[1:]:
        assert a.dtype == np.float64 and a.shape == a0.shape and a.strides == a0.strides, \
            "oracle kernels expect identically shaped/strided float64 views"
    st = (C.c_ssize_t * 3)(*[s // 8 for s in a0.strides])
    return a0.shape, st


def _p(a):
    return a.ctypes.data_as(_PD)


class Kernels(object):
    """The f2py modules of the reference, as attributes (same routine names)."""

    def __init__(self, flavour="strict"):
        self.L = lib(flavour)

    # fortran_vorticity.vorticity(ui, uj, wk)
    def vorticity(self, ui, uj, wk):
        (l, m, n), st = _common([ui, uj, wk])
        self.L.orc_vorticity(_p(ui), _p(uj), _p(wk), l, m, n, st)

    # fortran_upwind.upwind(trac, u, dtrac, order)
    def upwind(self, trac, u, dtrac, order=5):
        (l, m, n), st = _common([trac, u, dtrac])
        assert n >= 5
        self.L.orc_upwind(_p(trac), _p(u), _p(dtrac), l, m, n, st)

    # fortran_vortex_force.vortex_force_direc / _flip (U, vort, res, order).  linear=True: the branch the Fortran
    # takes when its local flag `linear` is .true. (dormant in the reference as shipped)
    def vortex_force_direc(self, U, vort, res, order=5, linear=False):
        (m, n, l), st = _common([U, vort, res])
        assert l >= 5
        if linear:
            self.L.orc_vortex_force_direc_linear(_p(U), _p(vort), _p(res), int(order), m, n, l, st)
        else:
            self.L.orc_vortex_force_direc(_p(U), _p(vort), _p(res), m, n, l, st)

    def vortex_force_flip(self, U, vort, res, order=5, linear=False):
        (m, n, l), st = _common([U, vort, res])
        assert n >= 5
        if linear:
            self.L.orc_vortex_force_flip_linear(_p(U), _p(vort), _p(res), int(order), m, n, l, st)
        else:
            self.L.orc_vortex_force_flip(_p(U), _p(vort), _p(res), m, n, l, st)

    def upwind_linear(self, trac, u, dtrac, order):
        (l, m, n), st = _common([trac, u, dtrac])
        assert n >= 5
        self.L.orc_upwind_linear(_p(trac), _p(u), _p(dtrac), int(order), l, m, n, st)

    def interpolate_vf(self, vU, order):
        """core/interpolate.f90 (vortex-force flavour): returns (qp(0:n-1), qm(1:n)), unassigned entries NaN."""
        vU = np.ascontiguousarray(vU, dtype=np.float64)
        qp, qm = np.full(len(vU), np.nan), np.full(len(vU), np.nan)
        self.L.orc_interpolate_vf(_p(vU), _p(qp), _p(qm), int(order), len(vU))
        return qp, qm

    def interpolate_tr(self, q, order):
        """core/interpolate_tracer.f90: returns (qp(1:n), qm(1:n)), unassigned entries NaN."""
        q = np.ascontiguousarray(q, dtype=np.float64)
        qp, qm = np.full(len(q), np.nan), np.full(len(q), np.nan)
        self.L.orc_interpolate_tr(_p(q), _p(qp), _p(qm), int(order), len(q))
        return qp, qm

    # fortran_kinenergy.kin(u, v, ke, ds2, order)
    def kin(self, u, v, ke, ds2, order=2):
        (l, m, n), st = _common([u, ke])
        self.L.orc_kin(_p(u), _p(ke), C.c_double(ds2), l, m, n, st)

    # fortran_bernoulli.gradke / gradkeandb / div
    def gradke(self, ke, du):
        (l, m, n), st = _common([ke, du])
        self.L.orc_gradke(_p(ke), _p(du), l, m, n, st)

    def gradkeandb(self, ke, b, du, dz):
        (l, m, n), st = _common([ke, b, du])
        self.L.orc_gradkeandb(_p(ke), _p(b), _p(du), C.c_double(dz), l, m, n, st)

    def div(self, d, u, iflag):
        (l, m, n), st = _common([d, u])
        self.L.orc_div(_p(d), _p(u), int(iflag), l, m, n, st)

    # fortran_dissipation.add_laplacian(phi, dphi, coef)
    def add_laplacian(self, phi, dphi, coef):
        (l, m, n), st = _common([phi, dphi])
        self.L.orc_add_laplacian(_p(phi), _p(dphi), C.c_double(coef), l, m, n, st)

    # 1-D helpers for unit tests
    def flux1d(self, u, q):
        u = np.ascontiguousarray(u, dtype=np.float64)
        q = np.ascontiguousarray(q, dtype=np.float64)
        flux = np.zeros_like(u)
        self.L.orc_flux1d(_p(u), _p(q), _p(flux), len(u))
        return flux

    def weno3(self, qm, q0, qp):
        return self.L.orc_weno3(qm, q0, qp)

    def weno5(self, qmm, qm, q0, qp, qpp):
        return self.L.orc_weno5(qmm, qm, q0, qp, qpp)


class OracleMG(object):
    """mgfordriver.MG (core/mgfordriver.py:5-129) on top of the C restatement of mgfor.

    Single process only (npx = npy = 1)."""

    IVAR = dict(x=1, b=2, r=3, y=4, diag=5, idiag=6, msk=7, Rcoef=8, Pcoef=9)

    def __init__(self, npx, npy, nx, ny, nz, nh, topology=1, flavour="strict"):
        assert npx == 1 and npy == 1 and nh == 3
        self.L = lib(flavour)
        self.nh = nh
        self.mg = C.c_void_p(self.L.orc_mg_create(nx, ny, nz, topology))
        if not self.mg:
            raise ValueError("grid %dx%dx%d cannot be coarsened by mgfor" % (nx, ny, nz))
        self.nlevels = self.L.orc_mg_nlevels(self.mg)
        self.shape = self.get_arrayshape()
        self.stats = {"normb": 0, "res": [0], "blowup": False}

    def __del__(self):
        try:
            self.L.orc_mg_free(self.mg)
        except Exception:
            pass

    def get_arrayshape(self, lev=1):
        s = (C.c_int * 3)()
        self.L.orc_mg_shape(self.mg, lev, s)
        return tuple(s)

    def get_idx_from_neighbours(self, neighbours):
        nh = self.nh
        i0 = 0 if (0, 0, -1) in neighbours else nh
        i1 = None if (0, 0, 1) in neighbours else -nh
        j0 = 0 if (0, -1, 0) in neighbours else nh
        j1 = None if (0, 1, 0) in neighbours else -nh
        k0 = 0 if (-1, 0, 0) in neighbours else nh
        k1 = None if (1, 0, 0) in neighbours else -nh
        return (slice(k0, k1), slice(j0, j1), slice(i0, i1))

    def preallocate_for_nyles(self, dx, neighbours, halo):
        self.dx = dx
        self.idx = self.get_idx_from_neighbours(neighbours)
        self.x = np.zeros(self.shape)
        self.b = np.zeros(self.shape)
        self.halo = halo

    def set_array(self, array, ivar=1, lev=1):
        a = np.ascontiguousarray(array, dtype=np.float64)
        assert a.shape == self.get_arrayshape(lev)
        assert self.L.orc_mg_set_array(self.mg, lev, ivar, _p(a)) == 0

    def get_array(self, array=None, ivar=1, lev=1):
        if array is None:
            array = np.zeros(self.get_arrayshape(lev))
        assert array.flags.c_contiguous and array.shape == self.get_arrayshape(lev)
        assert self.L.orc_mg_get_array(self.mg, lev, ivar, _p(array)) == 0
        return array

    def set_mask(self, msk):
        """mgfor/tests.f90:207-212: a user mask at level 1, then setup_fine_msk + setup_operators."""
        self.set_array(msk, ivar=7)
        self.L.orc_mg_setup_fine_msk(self.mg)
        self.L.orc_mg_setup_operators(self.mg)

    def _solve(self):
        self.L.orc_mg_solve(self.mg)
        nite, res, normb = C.c_int(), C.c_double(), C.c_double()
        self.L.orc_mg_stats(self.mg, C.byref(nite), C.byref(res), C.byref(normb))
        hist = np.zeros(64)
        n = self.L.orc_mg_reshist(self.mg, _p(hist))
        self.nite, self.res, self.normb = nite.value, res.value, normb.value
        self.reshist = hist[:n].copy()

    def solve_directly(self, p, div):
        self.halo.fill(div)
        self.b[self.idx] = div
        self.set_array(self.b, ivar=2)
        self._solve()
        self.get_array(self.x, ivar=1)
        p[:, :, :] = self.x[self.idx] * self.dx ** 2

    def solve(self, x, b):
        self.set_array(b, ivar=2)
        self._solve()
        self.get_array(x, ivar=1)

    # single operators for operator-level parity tests
    def op(self, name, lev):
        getattr(self.L, "orc_mg_" + name)(self.mg, lev)
