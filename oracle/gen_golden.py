"""ORACLE tooling: generate tests/golden/*.npz by running the REFERENCE's own Python
drivers (core/model_les.py, model_les_euler.py, timescheme.py, vorticity.py,
vortex_force.py, tracer.py, kinenergy.py, bernoulli.py, projection.py, cov_to_contra.py,
mgfordriver.py, mpi/halo.py, mpi/topology.py, variables.py, grid.py, parameters.py,
nyles.Nyles.compute_dt) imported unmodified from /root/reference.

The reference's compiled layers cannot be built in this image (no Fortran compiler, no
MPI), so the modules they provide are substituted at import time:
    fortran_*            -> oracle/kernels.py (C restatement, strict build)
    mgmod / libmgmod64   -> oracle/csrc/oracle_mg.c through the same ctypes-style calls
    mpi4py               -> an in-process single-rank fake (tag-matched persistent requests)
    matplotlib, netCDF4, mg (old multigrid) -> empty stand-ins (never called on this path)
What the fixtures therefore pin: every line of reference *Python* on the hot path
(view permutations, call order, halo logic, MG embedding, time schemes).  What they do
not pin: the Fortran arithmetic itself (restated in oracle/csrc, "parity unpinned").

Run in the build container only:   python -m oracle.gen_golden
/root/reference does not exist on the GPU box; tests read only the committed .npz files.
"""
import ctypes
import os
import sys
import types

import numpy as np

REF = "/root/reference/core"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


# ------------------------------------------------------------------ stand-ins
def install_stubs():
    from oracle.kernels import Kernels, lib
    K = Kernels("strict")
    L = lib("strict")

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("fortran_vorticity", vorticity=K.vorticity)
    mod("fortran_vortex_force", vortex_force_direc=K.vortex_force_direc, vortex_force_flip=K.vortex_force_flip)
    mod("fortran_upwind", upwind=K.upwind)
    mod("fortran_kinenergy", kin=K.kin)
    mod("fortran_bernoulli", gradke=K.gradke, gradkeandb=K.gradkeandb, div=K.div)
    mod("fortran_dissipation", add_laplacian=K.add_laplacian)
    mod("mg")
    mod("netCDF4")
    mpl = mod("matplotlib")
    mpl.pyplot = mod("matplotlib.pyplot")
    tk = mod("mpl_toolkits")
    tk.mplot3d = mod("mpl_toolkits.mplot3d", axes3d=None)

    # ---- mpi4py: one rank; persistent requests matched by tag -------------
    mailbox = {}

    class Req(object):
        def __init__(self, kind, buf, tag):
            self.kind, self.buf, self.tag = kind, buf, tag

    class Comm(object):
        def Get_rank(self): return 0
        def Get_size(self): return 1
        def Send_init(self, buf, dest, tag=0): return Req("s", buf, tag)
        def Recv_init(self, buf, source, tag=0): return Req("r", buf, tag)
        def allreduce(self, x, op=None): return x
        def barrier(self): pass
        def Barrier(self): pass

    class Prequest(object):
        @staticmethod
        def Startall(reqs):
            for r in reqs:
                if r.kind == "s":
                    mailbox[r.tag] = r.buf.copy()

        @staticmethod
        def Waitall(reqs):
            for r in reqs:
                if r.kind == "r":
                    r.buf[...] = mailbox.pop(r.tag)

    MPI = types.SimpleNamespace(COMM_WORLD=Comm(), Prequest=Prequest, SUM="sum", MAX="max")
    mod("mpi4py", MPI=MPI)

    # ---- mgmod: the generated ctypes module of libmgmod64.so ----------------
    def val(a):
        return a._obj.value if hasattr(a, "_obj") else a

    def as_np(p, n):
        return np.ctypeslib.as_array(p, shape=(n,))

    class Fn(object):
        def __init__(self, name):
            self.name = name
            outer = self

            class _F(object):          # plain object so callers may set .restype like on a ctypes symbol
                def __call__(self, *a):
                    return outer._f(*a)
            self.f = _F()
        # called with raw python values (build.Function.__call__ converts by itself)

        def __call__(self, *a):
            assert self.name == "get_ptrmg"
            npx, npy, nx, ny, nz, vertices, short, is3d, topology = a
            assert npx == npy == 1 and is3d and not vertices and not short
            L.orc_mg_create.restype = ctypes.c_void_p
            return ctypes.c_void_p(L.orc_mg_create(nx, ny, nz, topology))

        def _f(self, mg, *a):
            if self.name == "print_mginfos":
                return
            if self.name == "solve":
                L.orc_mg_solve(mg)
                return
            if self.name == "get_pyshape":
                lev, shp = val(a[0]), a[1]
                s = (ctypes.c_int * 3)()
                L.orc_mg_shape(mg, lev, s)
                out = as_np(shp, 3)
                out[:] = list(s)
                return
            lev, ivar, n1, n2, n3, x = val(a[0]), val(a[1]), val(a[2]), val(a[3]), val(a[4]), a[5]
            s = (ctypes.c_int * 3)()
            L.orc_mg_shape(mg, lev, s)
            assert (n3, n2, n1) == tuple(s), "array extents passed to mgfor do not match the level"
            if self.name == "set_pyarray":
                assert L.orc_mg_set_array(mg, lev, ivar, x) == 0
            elif self.name == "get_pyarray":
                assert L.orc_mg_get_array(mg, lev, ivar, x) == 0
            else:
                raise KeyError(self.name)

    mod("mgmod", get=lambda name, isfloat32=False: Fn(name))
    for p in (REF, os.path.join(REF, "mpi")):
        if p not in sys.path:
            sys.path.insert(0, p)
    return L


# ------------------------------------------------------------------ cases
CASES = {
    # name: (modelname, geometry, (nx,ny,nz), (Lx,Ly,Lz), extras, nsteps)
    "les_closed": ("LES", "closed", (16, 8, 8), (4.0, 2.0, 2.0), {"uamp": 2.0}, 4),
    "les_perio_xy_rot": ("LES", "perio_xy", (8, 8, 16), (1.0, 1.0, 2.0), {"rotating": True, "coriolis": 3.0}, 3),
    "euler_perio_xyz": ("Euler3d", "perio_xyz", (8, 8, 8), (2 * np.pi,) * 3, {"uamp": 1.0, "dt_max": 1.0}, 4),
    "les_closed_rk3": ("LES", "closed", (8, 16, 8), (1.0, 2.0, 1.0), {"timestepping": "RK3_SSP"}, 2),
    "les_closed_ef_diff": ("LES", "closed", (8, 8, 8), (1.0, 1.0, 1.0),
                           {"timestepping": "EF", "diff_coef": {"u": 1e-3, "b": 2e-3}}, 3),
    # modelname "linear": LES(linear=True) -- no vortex force, no vorticity / kinetic energy (nyles.py:93-100)
    "linear_closed": ("linear", "closed", (8, 8, 16), (1.0, 1.0, 2.0), {"uamp": 0.5}, 3),
    # one passive tracer next to the buoyancy (model_les.py:36-43, tracer.py:44-72)
    "les_closed_tracer": ("LES", "closed", (16, 8, 8), (2.0, 1.0, 1.0), {"n_tracers": 1, "uamp": 1.0}, 3),
    # rotating frame + a user forcing object, as experiments/forced_convection/forced_plume.py sets them up
    "les_closed_forced_rot": ("LES", "closed", (8, 8, 16), (1.0, 1.0, 2.0),
                              {"forced": True, "rotating": True, "coriolis": 1.0, "forcing": "plume", "uamp": 0.5}, 3),
    # LFAM3 with viscosity and tracer diffusion (diff_coef): the corrector adds the Laplacians (model_les.py:141-142,
    # tracer.py:74-77), the leapfrog predictor does not -- the buffers of both must stay consistent across steps
    "les_closed_lfam3_visc": ("LES", "closed", (8, 8, 16), (1.0, 1.0, 2.0),
                              {"diff_coef": {"u": 5e-3, "b": 2e-3}, "uamp": 1.0}, 6),
}


def initial_fields(case, shape, seed=1234, uamp=1e-3):
    """Seeded, smooth-plus-noise initial state in canonical (k,j,i) order (whole arrays)."""
    rng = np.random.default_rng(seed)
    nz, ny, nx = shape
    z, y, x = np.meshgrid(np.linspace(0, 1, nz), np.linspace(0, 1, ny), np.linspace(0, 1, nx), indexing="ij")
    f = {}
    f["b"] = np.tanh((x - 0.5 + 0.05 * rng.standard_normal(shape)) * 6) + 0.1 * z
    f["u_i"] = uamp * (np.sin(2 * np.pi * x) * np.cos(2 * np.pi * y) + 0.1 * rng.standard_normal(shape))
    f["u_j"] = uamp * (-np.cos(2 * np.pi * x) * np.sin(2 * np.pi * y) + 0.1 * rng.standard_normal(shape))
    f["u_k"] = uamp * (0.3 * np.sin(2 * np.pi * z) + 0.1 * rng.standard_normal(shape))
    return f


def run_case(name):
    import parameters
    import topology as topo
    import grid as grid_module
    import model_les
    import model_les_euler
    import nyles as nyles_module  # noqa: F401  (compute_dt is borrowed below)

    modelname, geometry, (nx, ny, nz), (Lx, Ly, Lz), extra, nsteps = CASES[name]
    up = parameters.UserParameters()
    up.model["modelname"] = modelname
    up.model["geometry"] = geometry
    up.model["Lx"], up.model["Ly"], up.model["Lz"] = Lx, Ly, Lz
    up.discretization["global_nx"], up.discretization["global_ny"], up.discretization["global_nz"] = nx, ny, nz
    up.time["cfl"], up.time["dt_max"] = 0.8, extra.get("dt_max", 0.05)
    for k, v in extra.items():
        for cat in ("model", "physics", "time", "discretization"):
            if k in getattr(up, cat):
                getattr(up, cat)[k] = v
    up.check()
    param = up.view_parameters()
    param["nx"], param["ny"], param["nz"] = nx, ny, nz
    topo.topology = geometry
    procs = [1, 1, 1]
    loc = topo.rank2loc(0, procs)
    param.update(procs=procs, myrank=0, loc=loc, neighbours=topo.get_neighbours(loc, procs))
    grid = grid_module.Grid(param)
    if modelname == "Euler3d":
        model = model_les_euler.LES(param, grid)
    elif modelname == "linear":
        model = model_les.LES(param, grid, linear=True)
    else:
        model = model_les.LES(param, grid)
    if extra.get("forcing") == "plume":
        sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tests"))
        from golden_cases import PlumeForcing
        model.forcing = PlumeForcing(param, grid)          # "the user must attach the forcing to the model"
    st = model.state
    tracers = ["t%d" % i for i in range(param["n_tracers"])] if modelname != "Euler3d" else []
    shape = st.b.view("i").shape
    ic = initial_fields(name, shape, uamp=extra.get("uamp", 1e-3))
    st.b.view("i")[:] = ic["b"] if modelname != "Euler3d" else 0.0
    for d in "ijk":
        st.u[d].view("i")[:] = ic["u_" + d]
    for n, nick in enumerate(tracers):
        zz, yy, xx = np.meshgrid(*[np.linspace(0, 1, m) for m in shape], indexing="ij")
        ic[nick] = np.cos((2 + n) * np.pi * xx) * np.sin(np.pi * zz) + 0.2 * yy
        st.get(nick).view("i")[:] = ic[nick]

    fake = types.SimpleNamespace(auto_dt=True, model=model, cfl=param["cfl"], dt_max=param["dt_max"], dt0=param["dt"])
    out = {"shape": np.array(shape), "nsteps": nsteps}
    for k, v in ic.items():
        out["ic_" + k] = v

    def snap(tag):
        for sname in ["b", "p", "ke", "div"] + tracers:
            out["%s_%s" % (tag, sname)] = getattr(st, sname).view("i").copy()
        for vname in ("u", "U", "vor"):
            for d in "ijk":
                out["%s_%s_%s" % (tag, vname, d)] = getattr(st, vname)[d].view("i").copy()

    model.diagnose_var(st)
    snap("diag0")
    # one bare RHS evaluation (operator-level pin)
    ds = st.duplicate_prognostic_variables()
    model.rhs(st, 0.0, ds, last=True)
    out["rhs0_b"] = ds.b.view("i").copy()
    for nick in tracers:
        out["rhs0_" + nick] = ds.get(nick).view("i").copy()
    for d in "ijk":
        out["rhs0_u_" + d] = ds.u[d].view("i").copy()

    t, dts = 0.0, []
    for n in range(nsteps):
        dt = nyles_module.Nyles.compute_dt(fake)
        model.forward(t, dt)
        t += dt
        dts.append(dt)
    out["dts"] = np.array(dts)
    snap("final")
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name, "shape", shape, "t=%.5f" % t)


def main():
    if not os.path.isdir(REF):
        raise SystemExit("reference tree not present; golden fixtures can only be generated in the build container")
    install_stubs()
    # parameters.InextensibleDict.freeze is class-wide: never call it here
    for name in CASES:
        run_case(name)


if __name__ == "__main__":
    main()
