"""ORACLE (test infrastructure, NOT product code): NumPy restatement of the Nyles
Python drivers around the C kernels of oracle/kernels.py.

Follows, function by function:
  core/variables.py:52-399     Scalar / Vector / State (one canonical (k,j,i) array
                               per field; view('j'), view('k') are transposed aliases
                               with the reference's index order, variables.py:146-150)
  core/mpi/topology.py:72-157, 253-304   neighbours, halo-aware array extents
  core/mpi/halo.py:93-178      26-neighbour halo fill (single process: periodic wrap)
  core/grid.py:13-60           metric terms, cell-centre coordinates
  core/cov_to_contra.py:4-20, vorticity.py:7-34, vortex_force.py:69-81,
  kinenergy.py:7-24, bernoulli.py:14-32, tracer.py:44-77, projection.py:16-87
  core/timescheme.py:113-221   EF / LFAM3 / RK3_SSP
  core/model_les.py:33-175, core/model_les_euler.py (diff: no tracer, no pre-projection fill)
  core/nyles.py:227-260        compute_dt

The orchestration is pinned against the reference's *own* Python modules by
oracle/gen_golden.py (run in the build container, where /root/reference exists).
"""
import itertools

import numpy as np

from .kernels import Kernels, OracleMG

TOPO_ENUM = {"closed": 1, "perio_x": 2, "perio_y": 3, "perio_xy": 5, "perio_xyz": 6}


# ---------------------------------------------------------------- topology
def get_neighbours(loc, procs, topology):
    """core/mpi/topology.py:72-157 with extension=26, incr=1."""
    k, j, i = loc
    nz, ny, nx = procs
    ngs = {}
    for dk, dj, di in itertools.product([-1, 0, 1], repeat=3):
        if dk == dj == di == 0:
            continue
        ok = True
        if "x" not in topology and not (0 <= i + di < nx):
            ok = False
        if "y" not in topology and not (0 <= j + dj < ny):
            ok = False
        if "z" not in topology and not (0 <= k + dk < nz):
            ok = False
        if ok:
            l = [(k + dk) % nz, (j + dj) % ny, (i + di) % nx]
            ngs[(dk, dj, di)] = (l[0] * ny + l[1]) * nx + l[2]
    return ngs


def get_variable_shape(innersize, ngbs, nh):
    """core/mpi/topology.py:253-304."""
    size = list(innersize)
    lo = []
    for ax, key in enumerate([(-1, 0, 0), (0, -1, 0), (0, 0, -1)]):
        if key in ngbs:
            size[ax] += nh
            lo.append(nh)
        else:
            lo.append(0)
    hi = list(size)
    for ax, key in enumerate([(1, 0, 0), (0, 1, 0), (0, 0, 1)]):
        if key in ngbs:
            size[ax] += nh
    return size, (lo[0], hi[0], lo[1], hi[1], lo[2], hi[2])


# ---------------------------------------------------------------- variables
class Scalar(object):
    AXES = {"i": (0, 1, 2), "j": (2, 0, 1), "k": (1, 2, 0)}

    def __init__(self, param, nickname, prognostic=False):
        self.param = param
        self.nickname = nickname
        self.prognostic = prognostic
        shape = [param["nz"], param["ny"], param["nx"]]
        size, self.domainindices = get_variable_shape(shape, param["neighbours"], param["nh"])
        self.shape = shape
        self.size = {"i": size[2], "j": size[1], "k": size[0]}
        self.data = np.zeros(size)
        self.activeview = "i"

    def duplicate(self):
        return Scalar(self.param, self.nickname, self.prognostic)

    def view(self, idx=None):
        return self.data.transpose(self.AXES[idx or "i"])

    def flipview(self, idx):
        return self.view({"i": "j", "j": "k", "k": "i"}[idx])

    def viewlike(self, other):
        return self.view("i")


class Vector(dict):
    def __init__(self, param, nickname, prognostic=False):
        for d in "ijk":
            self[d] = Scalar(param, nickname + "_" + d, prognostic)
        self.param, self.nickname, self.prognostic = param, nickname, prognostic

    def duplicate(self):
        return Vector(self.param, self.nickname, self.prognostic)


MODELVAR = [("b", "scalar", True), ("p", "scalar", False), ("ke", "scalar", False),
            ("div", "scalar", False), ("u", "vector", True), ("U", "vector", False),
            ("vor", "vector", False), ("work", "scalar", False)]


class State(object):
    def __init__(self, variables):
        self.toc = {}
        for v in variables:
            self.toc[v.nickname] = "scalar" if isinstance(v, Scalar) else "vector"
            setattr(self, v.nickname, v)

    def get(self, name):
        if len(name) > 2 and name[-2] == "_" and name[-1] in "ijk":
            return getattr(self, name[:-2])[name[-1]]
        return getattr(self, name)

    def duplicate_prognostic_variables(self):
        return State([getattr(self, n).duplicate() for n in self.toc if getattr(self, n).prognostic])

    def get_prognostic_scalars(self):
        out = []
        for n in self.toc:
            if getattr(self, n).prognostic:
                out += [n] if self.toc[n] == "scalar" else ["%s_%s" % (n, d) for d in "ijk"]
        return out


def get_state(param, extra_tracers=()):
    vs = []
    for nick, kind, prog in MODELVAR:
        vs.append(Scalar(param, nick, prog) if kind == "scalar" else Vector(param, nick, prog))
    for t in extra_tracers:
        vs.append(Scalar(param, t, True))
    return State(vs)


# ---------------------------------------------------------------- halo (one process)
class Halo(object):
    """core/mpi/halo.py:93-178 when every neighbour is the process itself."""

    def __init__(self, param, scalar):
        self.nh = param["nh"]
        self.neighbours = param["neighbours"]
        self.domi = scalar.domainindices
        self.size = scalar.data.shape

    def fill(self, thing):
        if isinstance(thing, Scalar):
            self.fillarray(thing.view("i"))
        elif isinstance(thing, np.ndarray):
            self.fillarray(thing)
        else:
            for d in "ijk":
                self.fillarray(thing[d].view("i"))

    def fillarray(self, x):
        nh, domi = self.nh, self.domi
        iidx, oidx = [], []
        for l in range(3):
            iidx.append({-1: slice(-nh - nh, -nh), 1: slice(nh, nh + nh), 0: slice(domi[2 * l], domi[2 * l + 1])})
            oidx.append({-1: slice(0, nh), 1: slice(self.size[l] - nh, self.size[l]),
                         0: slice(domi[2 * l], domi[2 * l + 1])})
        bufs = {}
        for (dk, dj, di) in self.neighbours:
            bufs[(dk, dj, di)] = x[iidx[0][-dk], iidx[1][-dj], iidx[2][-di]].copy()
        # message sent towards `direc` lands in the halo on the receiver's `-direc` side
        for (dk, dj, di), buf in bufs.items():
            x[oidx[0][-dk], oidx[1][-dj], oidx[2][-di]] = buf


# ---------------------------------------------------------------- grid
class Grid(object):
    def __init__(self, param):
        self.nx, self.ny, self.nz = param["nx"], param["ny"], param["nz"]
        self.npx, self.npy, self.npz = param["npx"], param["npy"], param["npz"]
        self.Lx, self.Ly, self.Lz = param["Lx"], param["Ly"], param["Lz"]
        self.dx = self.Lx / (self.npx * self.nx)
        self.dy = self.Ly / (self.npy * self.ny)
        self.dz = self.Lz / (self.npz * self.nz)
        self.idx2 = 1 / self.dx ** 2
        self.idy2 = 1 / self.dy ** 2
        self.idz2 = 1 / self.dz ** 2
        self.ids2 = {"i": self.idx2, "j": self.idy2, "k": self.idz2}
        s = Scalar(param, "x_b")
        k0, k1, j0, j1, i0, i1 = s.domainindices
        loc = param.get("loc", [0, 0, 0])
        self.x_b_1D = (np.arange(s.size["i"]) + 0.5 - i0) * self.dx + loc[2] * self.nx * self.dx
        self.y_b_1D = (np.arange(s.size["j"]) + 0.5 - j0) * self.dy + loc[1] * self.ny * self.dy
        self.z_b_1D = (np.arange(s.size["k"]) + 0.5 - k0) * self.dz + loc[0] * self.nz * self.dz
        self.z_b, self.y_b, self.x_b = np.meshgrid(self.z_b_1D, self.y_b_1D, self.x_b_1D, indexing="ij")


# ---------------------------------------------------------------- operator modules
def U_from_u(state, grid):
    state.U["i"].view("i")[:] = state.u["i"].view("i") * grid.idx2
    state.U["j"].view("i")[:] = state.u["j"].view("i") * grid.idy2
    state.U["k"].view("i")[:] = state.u["k"].view("i") * grid.idz2


def vorticity(K, state, fparameter):
    perm = {"i": ("k", "j"), "j": ("i", "k"), "k": ("j", "i")}
    for dirk in "ijk":
        dirj, diri = perm[dirk]
        ui = state.u[diri].flipview(dirk)
        uj = state.u[dirj].flipview(dirk)
        wk = state.vor[dirk].flipview(dirk)
        K.vorticity(ui, uj, wk)
        if fparameter > 0.0 and dirk == "k":
            wk[:, :-1, :-1] += fparameter


def vortex_force(K, state, rhs, order=5, linear=False):
    for k, j, i in ["ikj", "jik", "kji"]:
        u_i = rhs.u[i].flipview(j)
        u_k = rhs.u[k].flipview(j)
        U_i = state.U[i].flipview(j)
        U_k = state.U[k].flipview(j)
        w_j = state.vor[j].flipview(j)
        K.vortex_force_direc(U_k, w_j, u_i, order, linear=linear)
        K.vortex_force_flip(U_i, w_j, u_k, order, linear=linear)


def kinenergy(K, state, grid, order=2):
    for d in "ijk":
        u = state.u[d].view(d)
        ke = state.ke.view(d)
        if d == "i":
            ke[...] = 0.0
        K.kin(u, u, ke, grid.ids2[d], order)


def bernoulli(K, state, rhs, grid, euler=False):
    for d in "ijk":
        du = rhs.u[d].view(d)
        ke = state.ke.view(d)
        if d in "ij" or euler:
            K.gradke(ke, du)
        else:
            K.gradkeandb(ke, state.b.view(d), du, grid.dz)


def rhstrac(K, state, rhs, traclist, order=5, diff_coef=None, ids2=None, last=False, linear=False):
    for name in traclist:
        trac, dtrac = state.get(name), rhs.get(name)
        for d in "ijk":
            vel = state.U[d].view(d)
            field, dfield = trac.view(d), dtrac.view(d)
            if d == "i":
                dfield[...] = 0.0
            if linear:
                K.upwind_linear(field, vel, dfield, order)
            else:
                K.upwind(field, vel, dfield, order)
            if diff_coef and last and name in diff_coef:
                K.add_laplacian(field, dfield, diff_coef[name] * ids2[d])


def add_viscosity(K, grid, state, dstate, viscosity):
    for direc in "ijk":
        coef = viscosity * grid.ids2[direc]
        for comp in "ijk":
            K.add_laplacian(state.u[comp].view(direc), dstate.u[comp].view(direc), coef)


def compute_div(K, state):
    for count, d in enumerate("ijk"):
        K.div(state.div.view(d), state.U[d].view(d), count)


def compute_p(K, mg, state, grid):
    compute_div(K, state)
    mg.solve_directly(state.p.view("i"), state.div.view("i"))
    for d in "ijk":
        K.gradke(state.p.view(d), state.u[d].view(d))


# ---------------------------------------------------------------- model + time scheme
def vf_work(K, state, dstate, work, domainindices):
    """core/online_diag.py:4-35: work[...] = sum_d centred(U_d * du_d / 2) with du the vortex force alone;
    returns the interior sum."""
    for d in "ijk":
        dstate.u[d].view("i")[...] = 0.0
    vortex_force(K, state, dstate)
    for d in "ijk":
        u1, u2, res = state.U[d].view(d), dstate.u[d].view(d), work.view(d)
        if d == "i":
            res[:] = 0.0
        product = u1 * u2 * 0.5
        res[:, :, 1:] += product[:, :, 1:] + product[:, :, :-1]
    k0, k1, j0, j1, i0, i1 = domainindices
    return work.view("i")[k0:k1, j0:j1, i0:i1].sum()


class LES(object):
    """model_les.LES / model_les_euler.LES (modelname 'LES' | 'Euler3d' | 'linear')."""

    def __init__(self, param, flavour="strict", linear_upwind=False):
        self.K = Kernels(flavour)
        self.param = param
        # the Fortran's local flag `linear` (fortran_upwind.f90:31, fortran_vortex_force.f90:28,108); .false. as shipped
        self.linear_upwind = linear_upwind
        self.modelname = param.get("modelname", "LES")
        self.euler = self.modelname == "Euler3d"
        self.nonlinear = self.modelname != "linear"
        self.grid = Grid(param)
        self.traclist = [] if self.euler else ["b"] + ["t%d" % i for i in range(param.get("n_tracers", 0))]
        self.state = get_state(param, extra_tracers=self.traclist[1:])
        self.halo = Halo(param, self.state.b)
        self.diff_coef = param.get("diff_coef", {})
        self.forced = param.get("forced", False)
        self.forcing = None
        self.fparameter = param.get("coriolis", 1.0) * self.grid.dx * self.grid.dy if param.get("rotating") else 0.0
        self.mg = OracleMG(1, 1, param["nx"], param["ny"], param["nz"], param["nh"],
                           TOPO_ENUM[param["geometry"]], flavour=flavour)
        self.mg.preallocate_for_nyles(self.grid.dx, param["neighbours"], self.halo)
        self.mg_log = []          # (nite, res, normb) of every solve
        self.ts = Timescheme(param, self.state, self.rhs, self.diagnose_var)

    def diagnose_var(self, state):
        K = self.K
        if not self.euler:
            self.halo.fill(state.b)
            self.halo.fill(state.u)
        U_from_u(state, self.grid)
        compute_p(K, self.mg, state, self.grid)
        self.mg_log.append((self.mg.nite, self.mg.res, self.mg.normb))
        self.halo.fill(state.u)
        U_from_u(state, self.grid)
        if self.nonlinear:
            vorticity(K, state, self.fparameter)
            kinenergy(K, state, self.grid)
            self.halo.fill(state.vor)
            self.halo.fill(state.ke)

    def rhs(self, state, t, dstate, last=False):
        K = self.K
        for name, kind in dstate.toc.items():
            if kind == "scalar":
                dstate.get(name).view("i")[...] = 0.0
            else:
                for d in "ijk":
                    dstate.get(name)[d].view("i")[...] = 0.0
        if not self.euler:
            # model_les.py:133 calls rhstrac(state, dstate) without `last`, so tracer
            # diffusion (tracer.py:74-77) is never active in the LES model
            rhstrac(K, state, dstate, self.traclist, order=self.param.get("orderA", 5), diff_coef=self.diff_coef,
                    ids2=self.grid.ids2, last=False, linear=self.linear_upwind)
        if self.nonlinear:
            vortex_force(K, state, dstate, order=self.param.get("orderVF", 5), linear=self.linear_upwind)
        bernoulli(K, state, dstate, self.grid, euler=self.euler)
        if last and "u" in self.diff_coef:
            add_viscosity(K, self.grid, state, dstate, self.diff_coef["u"])
        if self.forced and self.forcing is not None:
            self.forcing.add(state, dstate, t)

    def forward(self, t, dt):
        self.ts.forward(self.state, t, dt)

    def compute_dt(self):
        """core/nyles.py:227-260."""
        p = self.param
        if not p.get("auto_dt", True):
            return p["dt"]
        U = self.state.U["i"].view("i")
        V = self.state.U["j"].view("i")
        W = self.state.U["k"].view("i")
        U_max = np.sqrt(np.max(U ** 2 + V ** 2 + W ** 2))
        if U_max == 0.0:
            return p["dt_max"]
        return min(p["cfl"] / U_max, p["dt_max"])


class Timescheme(object):
    def __init__(self, param, state, rhs, diagnose_var):
        self.scheme = param.get("timestepping", "LFAM3")
        self.rhs, self.diagnose_var = rhs, diagnose_var
        self.names = state.get_prognostic_scalars()
        self.dstate = state.duplicate_prognostic_variables()
        if self.scheme == "LFAM3":
            self.stateb = state.duplicate_prognostic_variables()
            self.staten = state.duplicate_prognostic_variables()
            self.first = True
        if self.scheme == "RK3_SSP":
            self.ds0 = self.dstate
            self.ds1 = state.duplicate_prognostic_variables()
            self.ds2 = state.duplicate_prognostic_variables()

    def forward(self, state, t, dt):
        getattr(self, {"EF": "EulerForward", "LFAM3": "LFAM3", "RK3_SSP": "RK3_SSP"}[self.scheme])(state, t, dt)

    def EulerForward(self, state, t, dt):
        self.rhs(state, t, self.dstate, last=True)
        for n in self.names:
            s = state.get(n).view("i")
            s += dt * self.dstate.get(n).view("i")
        self.diagnose_var(state)

    def LFAM3(self, state, t, dt):
        self.rhs(state, t, self.dstate)
        if self.first:
            for n in self.names:
                s = state.get(n).view("i")
                ds = self.dstate.get(n).view("i")
                self.staten.get(n).view("i")[:] = s
                self.stateb.get(n).view("i")[:] = s
                s += dt * ds
            self.first = False
            self.diagnose_var(state)
            return
        for n in self.names:
            s = state.get(n).view("i")
            ds = self.dstate.get(n).view("i")
            sb = self.stateb.get(n).view("i")
            sn = self.staten.get(n).view("i")
            sn[:] = s
            s[:] = sb + (2. * dt) * ds
            s[:] = (1. / 12.) * (5. * s + 8. * sn - sb)
            sb[:] = sn
        self.diagnose_var(state)
        self.rhs(state, t + dt * .5, self.dstate, last=True)
        for n in self.names:
            s = state.get(n).view("i")
            s[:] = self.staten.get(n).view("i") + dt * self.dstate.get(n).view("i")
        self.diagnose_var(state)

    def RK3_SSP(self, state, t, dt):
        self.rhs(state, t, self.ds0, last=False)
        for n in self.names:
            s = state.get(n).view("i")
            s += dt * self.ds0.get(n).view("i")
        self.diagnose_var(state)
        self.rhs(state, t + dt, self.ds1, last=False)
        for n in self.names:
            s = state.get(n).view("i")
            ds0, ds1 = self.ds0.get(n).view("i"), self.ds1.get(n).view("i")
            s += (dt / 4.) * (ds1 - 3 * ds0)
        self.diagnose_var(state)
        self.rhs(state, t + dt * 0.5, self.ds2, last=True)
        for n in self.names:
            s = state.get(n).view("i")
            ds0, ds1, ds2 = (self.ds0.get(n).view("i"), self.ds1.get(n).view("i"), self.ds2.get(n).view("i"))
            s += (dt / 12.) * (8 * ds2 - ds0 - ds1)
        self.diagnose_var(state)


# ---------------------------------------------------------------- convenience
def make_param(nx, ny, nz, geometry="closed", Lx=1.0, Ly=1.0, Lz=1.0, **kw):
    """Flat parameter dict as Nyles.__init__ builds it (core/nyles.py:44-74), one process."""
    p = dict(nx=nx, ny=ny, nz=nz, global_nx=nx, global_ny=ny, global_nz=nz, nh=3,
             npx=1, npy=1, npz=1, Lx=Lx, Ly=Ly, Lz=Lz, geometry=geometry, modelname="LES",
             timestepping="LFAM3", auto_dt=True, dt=0.1, cfl=1.0, dt_max=1.0, n_tracers=0,
             rotating=False, forced=False, coriolis=1.0, diff_coef={}, orderA=5, orderVF=5, orderKE=2,
             procs=[1, 1, 1], myrank=0, loc=[0, 0, 0])
    p.update(kw)
    p["neighbours"] = get_neighbours(p["loc"], p["procs"], p["geometry"])
    return p
