"""LES model, API of core/model_les.py:31-175 (and, with euler=True, core/model_les_euler.py).

At each step the model does not suppose that the previous velocity was divergence free; it
projects after every update (Ferziger p.180), exactly like the reference:
    rhs          : tracer advection, vortex force, Bernoulli term (+ viscosity, forcing)
    diagnose_var : halo fills, U from u, pressure projection (multigrid), vorticity, kinetic energy
All fields are CUDA tensors in one canonical layout; the operator modules call libnyles_b200.so.
"""
import pickle

from . import variables as var
from . import tracer
from . import timescheme as ts
from . import vortex_force as vortf
from . import vorticity as vort
from . import bernoulli as bern
from . import kinenergy as kinetic
from . import viscosity as visc
from . import projection
from . import cov_to_contra
from . import halo
from . import lib
from .mgfordriver import MG
from .timing import timing

TOPOLOGY = {"closed": 1, "perio_xy": 5, "perio_xyz": 6}      # core/model_les.py:80-88


class LES(object):
    euler = False

    def __init__(self, param, grid, linear=False, fused=True):
        import nyles_b200
        lib.set_arith(nyles_b200.FAST_ARITH)
        self.nonlinear = not linear
        # one fused RHS launch pair instead of the per-operator calls; the linear-upwind branch has per-operator
        # kernels only
        self.fused = fused and not nyles_b200.LINEAR_UPWIND
        self.grid = grid
        self.traclist = [] if self.euler else ["b"]
        modelvar = dict(var.modelvar)
        if not self.euler:
            for i in range(param["n_tracers"]):
                nick = "t{}".format(i)
                self.traclist.append(nick)
                modelvar[nick] = var.ModelVariable("scalar", "tracer{}".format(i), "", True)
        self.state = var.get_state(param, modelvar)
        self.halo = halo.set_halo(param, self.state)
        self.neighbours = param["neighbours"]
        self.timescheme = ts.Timescheme(param, self.state)
        self.timescheme.set(self.rhs, self.diagnose_var, self.rhs_step, self._rest_rhs)
        self.orderA, self.orderVF, self.orderKE = param["orderA"], param["orderVF"], param["orderKE"]
        self.rotating = param["rotating"]
        self.forced = param["forced"]
        self.forcing = None
        self.diff_coef = param["diff_coef"]
        self.add_viscosity = "u" in self.diff_coef.keys()
        if self.add_viscosity:
            self.viscosity = self.diff_coef["u"]
        self.tracer = tracer.Tracer_numerics(param, grid, self.traclist, self.orderA, self.diff_coef)
        # Coriolis parameter as a covariant quantity: f * horizontal cell area (model_les.py:61-67)
        self.fparameter = param["coriolis"] * self.grid.dx * self.grid.dy if self.rotating else 0.

        geometry = param["geometry"]
        if geometry not in TOPOLOGY or (self.euler and geometry == "perio_xy"):
            raise ValueError("geometry %r is not supported by the multigrid (model_les.py:80-88)" % geometry)
        self.mg = MG(grid.npx, grid.npy, param["nx"], param["ny"], param["nz"], param["nh"], TOPOLOGY[geometry],
                     npz=param["npz"])
        self.mg.preallocate_for_nyles(grid.dx, param["neighbours"], self.halo)
        self.stats = []
        self._umax_key = None

    def cached_max_speed2(self):
        """max(U^2+V^2+W^2) left on the device by the last fused diagnose_var, or None if state.U has been
        written since (torch-level writes bump the tensors' version counters; the library's own writers of U
        bump lib.u_epoch)."""
        key = self._umax_key
        if key is None:
            return None
        state, epoch, versions, ptrs = key
        U = state.U
        if state is not self.state or epoch != lib.u_epoch or versions is None or versions != _versions(U) or \
                ptrs != tuple(U[d].tensor.data_ptr() for d in "ijk"):
            return None
        out = lib.C.c_double()
        lib.check(lib.load().ny_diag_post_max_speed2(lib.context(U["i"].tensor.device), lib.C.byref(out), lib.stream()))
        return out.value

    # ------------------------------------------------------------------
    @timing
    def diagnose_var(self, state):
        if not self.euler:                    # model_les_euler.py:98-99 has these two fills commented out
            # model_les.py:105-106: fill(b), fill(u) -- one exchange for the four arrays
            self.halo.fillarrays([state.b.tensor] + [state.u[d].tensor for d in "ijk"])
        if self.fused:
            # same statements as below in three launches around the solve plus one for the diagnostics
            self.mg.project(state, self.grid)
            # div, vor and ke are evaluated on every plane of the slab from exchanged u: the planes the RHS
            # reads (vor: 3 below / 2 above, ke: 1 above) already hold what the neighbour computes, bit for
            # bit, so only the periodic wraps of this rank are applied; the outermost halo plane of vor / ke
            # and the z halo of div (never read) are not refreshed from the neighbour
            lazy = self.halo.z_remote
            self.halo.fill(state.div, local_only=lazy)   # keeps the model's div array as the reference leaves it
            self.halo.fill(state.u)
            if self.nonlinear:
                u, U, w = state.u, state.U, state.vor
                t = u["i"].tensor
                lib.check(lib.load().ny_diag_post(
                    lib.context(t.device), lib.ptr(u["i"].tensor), lib.ptr(u["j"].tensor), lib.ptr(u["k"].tensor),
                    lib.ptr(U["i"].tensor), lib.ptr(U["j"].tensor), lib.ptr(U["k"].tensor),
                    lib.ptr(w["i"].tensor), lib.ptr(w["j"].tensor), lib.ptr(w["k"].tensor), lib.ptr(state.ke.tensor),
                    self.grid.idx2, self.grid.idy2, self.grid.idz2, float(self.fparameter), lib.ext(t), lib.stream()))
                # the launch also left max(U^2+V^2+W^2) on the device; valid until U is written again
                lib.u_epoch += 1
                self._umax_key = (state, lib.u_epoch, _versions(U),
                                  tuple(U[d].tensor.data_ptr() for d in "ijk"))
                self.halo.fill(state.vor, local_only=lazy)
                self.halo.fill(state.ke, local_only=lazy)
            else:
                cov_to_contra.U_from_u(state, self.grid)
            return
        cov_to_contra.U_from_u(state, self.grid)
        projection.compute_p(self.mg, state, self.grid, self.neighbours)
        self.halo.fill(state.u)
        cov_to_contra.U_from_u(state, self.grid)
        if self.nonlinear:
            vort.vorticity(state, self.fparameter)
            kinetic.kinenergy(state, self.grid, self.orderKE)
            self.halo.fill(state.vor)
            self.halo.fill(state.ke)

    @timing
    def rhs(self, state, t, dstate, last=False):
        extra_tracers = len(self.traclist) > 1
        if self.fused:
            self._rhs_fused(state, dstate)
            if extra_tracers:
                saved, self.tracer.traclist = self.tracer.traclist, self.traclist[1:]
                self.tracer.rhstrac(state, dstate)
                self.tracer.traclist = saved
        else:
            reset_state(dstate)
            if not self.euler:
                self.tracer.rhstrac(state, dstate)     # model_les.py:133: `last` is not forwarded
            if self.nonlinear:
                vortf.vortex_force(state, dstate, self.orderVF)
            if self.euler:
                bern.bernoulli_euler(state, dstate, self.grid)
            else:
                bern.bernoulli(state, dstate, self.grid)
        if last and self.add_viscosity:
            visc.add_viscosity(self.grid, state, dstate, self.viscosity)
        if self.forced:
            self.forcing.add(state, dstate, t)

    def _rhs_fused(self, state, dstate):
        U, w, du = state.U, state.vor, dstate.u
        t = U["i"].tensor
        flags = (1 if self.euler else 0) | (0 if self.nonlinear else 2)
        lib.check(lib.load().ny_rhs(
            lib.context(t.device), lib.ptr(None if self.euler else state.b.tensor),
            lib.ptr(U["i"].tensor), lib.ptr(U["j"].tensor), lib.ptr(U["k"].tensor),
            lib.ptr(w["i"].tensor), lib.ptr(w["j"].tensor), lib.ptr(w["k"].tensor), lib.ptr(state.ke.tensor),
            lib.ptr(None if self.euler else dstate.b.tensor),
            lib.ptr(du["i"].tensor), lib.ptr(du["j"].tensor), lib.ptr(du["k"].tensor),
            self.grid.dz, flags, lib.ext(t), lib.stream()))
        if self.euler:
            dstate.b.tensor.zero_()

    def rhs_step(self, state, t, mode, dt, stateb, staten, last=False):
        """rhs() and the time-scheme update of b and u in the two RHS launches themselves (ny_rhs_step;
        mode 1: LFAM3 start-up, 2: predictor, 3: corrector): no tendency is stored and each field is written
        once, into the buffer that the scheme no longer needs; the buffers of state / stateb / staten are
        then rotated (core/timescheme.py:131-175 leaves the same values in the same-named arrays).
        Returns the names it has updated, or None -- having done nothing -- when something must see the
        tendencies first or the fused path is off:
          * viscosity: the corrector adds the Laplacian of u to du (model_les.py:141-142).  The decision is
            taken per MODEL, not per call: the rotating predictor defers `stateb := state n` to the rotation
            of the corrector, so a step must not mix a fused predictor with an unfused corrector;
          * a forcing object that only offers the reference's `add(state, dstate, t)`.  A forcing object that
            also offers `device_tendencies(state, t) -> {nickname: tensor}` (what it would add to dstate,
            canonical (k,j,i) CUDA tensors, e.g. {"b": Q}) keeps the fused path: the kernels add those arrays
            to the finished tendencies exactly where `forcing.add` would."""
        if not self.fused or self.add_viscosity:
            return None
        add = {}
        if self.forced:
            fn = getattr(self.forcing, "device_tendencies", None)
            if fn is None:
                return None
            add = fn(state, t) or {}
        U, w = state.U, state.vor
        t0 = U["i"].tensor
        flags = (1 if self.euler else 0) | (0 if self.nonlinear else 2)
        names = ([] if self.euler else ["b"]) + ["u_i", "u_j", "u_k"]
        outs = stateb if mode == 3 else staten

        def ptr4(st):
            p = [lib.ptr(None if self.euler else st.b.tensor).value] + [lib.ptr(st.u[d].tensor).value for d in "ijk"]
            return lib.C.byref((lib.C.c_void_p * 4)(*p))
        addp = None
        if add:
            unknown = set(add) - set(names)
            if unknown:
                raise ValueError("device_tendencies: %s is not a prognostic field of the fused step" % sorted(unknown))
            tens = [add.get(n) for n in (["b"] if not self.euler else [None]) + ["u_i", "u_j", "u_k"]]
            tens = [getattr(a, "tensor", a) for a in tens]
            for a in tens:
                if a is not None and tuple(a.shape) != tuple(t0.shape):
                    raise ValueError("device_tendencies: arrays must have the shape of the fields, %s" % (tuple(t0.shape),))
            addp = lib.C.byref((lib.C.c_void_p * 4)(*[lib.ptr(a).value for a in tens]))
        lib.check(lib.load().ny_rhs_step(
            lib.context(t0.device), lib.ptr(U["i"].tensor), lib.ptr(U["j"].tensor), lib.ptr(U["k"].tensor),
            lib.ptr(w["i"].tensor), lib.ptr(w["j"].tensor), lib.ptr(w["k"].tensor), lib.ptr(state.ke.tensor),
            ptr4(state), ptr4(stateb), ptr4(staten), ptr4(outs), addp, mode, dt, self.grid.dz, flags, lib.ext(t0),
            lib.stream()))
        for name in names:
            S, B, N = state.get(name), stateb.get(name), staten.get(name)
            if mode == 3:                     # new state <- sb's buffer; sb <- sn (state at n); sn <- scratch
                S.tensor, B.tensor, N.tensor = B.tensor, N.tensor, S.tensor
            else:                             # new state <- sn's buffer; sn <- the old state array itself
                S.tensor, N.tensor = N.tensor, S.tensor
                if mode == 1:
                    B.tensor.copy_(N.tensor)
        return names

    def _rest_rhs(self, state, dstate):
        """Tendencies of the prognostic scalars that rhs_step does not update itself: the passive tracers
        (tracer.py:44-72) and, in the Euler model, the buoyancy that nothing moves (db = 0)."""
        if self.euler:
            dstate.b.tensor.zero_()
        if len(self.traclist) > 1:
            saved, self.tracer.traclist = self.tracer.traclist, self.traclist[1:]
            self.tracer.rhstrac(state, dstate)
            self.tracer.traclist = saved

    @timing
    def forward(self, t, dt):
        self.timescheme.forward(self.state, t, dt)
        return self.mg.stats["blowup"]

    def update_stats(self):
        stats = dict(self.mg.stats)
        stats["maxdiv"] = self.state.div.tensor.abs().max().item()
        self.stats += [stats]

    def write_stats(self, path):
        with open("%s/stats.pkl" % path, "bw") as fid:
            pickle.dump(self.stats, fid)


def _versions(U):
    """Write counters of the three U tensors (torch bumps `_version` on every in-place write; it is the
    counter autograd's own saved-tensor check reads).  A torch build without it gives None, which never
    matches: the cached maximum is then simply not used."""
    v = tuple(getattr(U[d].tensor, "_version", None) for d in "ijk")
    return None if None in v else v


def reset_state(state):
    for var_name, var_type in state.toc.items():
        if var_type == "scalar":
            state.get(var_name).tensor.zero_()
        else:
            for i in "ijk":
                state.get(var_name)[i].tensor.zero_()
