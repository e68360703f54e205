"""Time schemes, API of core/timescheme.py (Timescheme(param, state).set(rhs, diagnose_var).forward).

EF, LFAM3 (default) and RK3_SSP; the elementwise updates that the reference writes as NumPy
expressions over whole arrays (timescheme.py:113-221) are single fused kernels per field with the
same operation order.
"""
from . import lib


class Timescheme(object):
    def __init__(self, param, state):
        timestepping = param["timestepping"]
        self.prognostic_scalars = state.get_prognostic_scalars()
        table = {"EF": self.EulerForward, "LFAM3": self.LFAM3, "RK3_SSP": self.RK3_SSP}
        try:
            self.forward = table[timestepping]
        except KeyError:
            raise ValueError("unknown time scheme " + repr(timestepping))
        self.dstate = state.duplicate_prognostic_variables()
        if timestepping == "LFAM3":
            self.stateb = state.duplicate_prognostic_variables()
            self.state = state.duplicate_prognostic_variables()
            self.first = True
        if timestepping == "RK3_SSP":
            self.ds0 = self.dstate
            self.ds1 = state.duplicate_prognostic_variables()
            self.ds2 = state.duplicate_prognostic_variables()
        self.L = lib.load()
        # optional callable(state), run after the last update of the prognostic fields of a step and before the
        # diagnose_var that ends it (Nyles.step_host starts the download of the finished fields there)
        self.before_last_diagnose = None

    def _last_diagnose(self, state):
        if self.before_last_diagnose is not None:
            self.before_last_diagnose(state)
        self.diagnose_var(state)

    def set(self, rhs, diagnose_var, rhs_step=None, tracer_rhs=None):
        self.rhs = rhs
        self.diagnose_var = diagnose_var
        # optional: rhs + update of b and u inside the RHS launches, with buffer rotation (LES.rhs_step);
        # tracer_rhs(state, dstate) then supplies the tendencies of the remaining (passive) tracers
        self.rhs_step = rhs_step
        self.tracer_rhs = tracer_rhs

    def _rhs_and_update(self, state, t, mode, dt, last, others, fn):
        """The right-hand side and the update `fn` of every prognostic scalar; b and the velocity components
        are updated by the model's fused launches when it offers them (rhs_step), else by `fn`."""
        skip = None
        if self.rhs_step is not None:
            skip = self.rhs_step(state, t, mode, dt, self.stateb, self.state, last=last)
            if skip is not None and len(skip) < len(self.prognostic_scalars):
                self.tracer_rhs(state, self.dstate)
        if skip is None:
            self.rhs(state, t, self.dstate, last=last)
        for name in self.prognostic_scalars:
            if skip and name in skip:
                continue
            fn(*([state.get(name).tensor] + [o.get(name).tensor for o in others]))

    def _each(self, state, *others):
        for name in self.prognostic_scalars:
            s = state.get(name).tensor
            yield (s,) + tuple(o.get(name).tensor for o in others)

    # ----------------------------------------
    def EulerForward(self, state, t, dt, **kwargs):
        self.rhs(state, t, self.dstate, last=True)
        for s, ds in self._each(state, self.dstate):
            lib.check(self.L.ny_ts_axpy(lib.context(s.device), lib.ptr(s), lib.ptr(ds), dt, s.numel(), lib.stream()))
        self._last_diagnose(state)

    # ----------------------------------------
    def LFAM3(self, state, t, dt, **kwargs):
        L = self.L
        three = (self.dstate, self.stateb, self.state)
        if self.first:                                               # Euler forward start-up
            self._rhs_and_update(state, t, 1, dt, False, three, lambda s, ds, sb, sn: lib.check(
                L.ny_ts_lfam3_first(lib.context(s.device), lib.ptr(s), lib.ptr(ds), lib.ptr(sb), lib.ptr(sn), dt,
                                    s.numel(), lib.stream())))
            self.first = False
            self._last_diagnose(state)
            return
        self._rhs_and_update(state, t, 2, dt, False, three, lambda s, ds, sb, sn: lib.check(     # predictor
            L.ny_ts_lfam3_pred(lib.context(s.device), lib.ptr(s), lib.ptr(ds), lib.ptr(sb), lib.ptr(sn), dt,
                               s.numel(), lib.stream())))
        self.diagnose_var(state)
        self._rhs_and_update(state, t + dt * .5, 3, dt, True, (self.dstate, self.state),        # corrector at n+1/2
                             lambda s, ds, sn: lib.check(
            L.ny_ts_lfam3_corr(lib.context(s.device), lib.ptr(s), lib.ptr(ds), lib.ptr(sn), dt, s.numel(),
                               lib.stream())))
        self._last_diagnose(state)

    # ----------------------------------------
    def RK3_SSP(self, state, t, dt, **kwargs):
        L = self.L
        self.rhs(state, t, self.ds0, last=False)
        for s, d0 in self._each(state, self.ds0):
            lib.check(L.ny_ts_axpy(lib.context(s.device), lib.ptr(s), lib.ptr(d0), dt, s.numel(), lib.stream()))
        self.diagnose_var(state)
        self.rhs(state, t + dt, self.ds1, last=False)
        for s, d0, d1 in self._each(state, self.ds0, self.ds1):
            lib.check(L.ny_ts_rk3_stage2(lib.context(s.device), lib.ptr(s), lib.ptr(d0), lib.ptr(d1), dt,
                                         s.numel(), lib.stream()))
        self.diagnose_var(state)
        self.rhs(state, t + dt * 0.5, self.ds2, last=True)
        for s, d0, d1, d2 in self._each(state, self.ds0, self.ds1, self.ds2):
            lib.check(L.ny_ts_rk3_stage3(lib.context(s.device), lib.ptr(s), lib.ptr(d0), lib.ptr(d1), lib.ptr(d2),
                                         dt, s.numel(), lib.stream()))
        self._last_diagnose(state)
