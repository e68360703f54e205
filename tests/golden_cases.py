"""Shared description of the committed golden fixtures (tests/golden/*.npz).

The fixtures were produced by oracle/gen_golden.py, i.e. by the reference's own Python
drivers running on the C restatement of its Fortran kernels.  Keep in sync with CASES there.
"""
import numpy as np

CASES = {
    # name: (modelname, geometry, (nx,ny,nz), (Lx,Ly,Lz), extras, nsteps)
    "les_closed": ("LES", "closed", (16, 8, 8), (4.0, 2.0, 2.0), {}, 4),
    "les_perio_xy_rot": ("LES", "perio_xy", (8, 8, 16), (1.0, 1.0, 2.0), {"rotating": True, "coriolis": 3.0}, 3),
    "euler_perio_xyz": ("Euler3d", "perio_xyz", (8, 8, 8), (2 * np.pi,) * 3, {"dt_max": 1.0}, 4),
    "les_closed_rk3": ("LES", "closed", (8, 16, 8), (1.0, 2.0, 1.0), {"timestepping": "RK3_SSP"}, 2),
    "les_closed_ef_diff": ("LES", "closed", (8, 8, 8), (1.0, 1.0, 1.0),
                           {"timestepping": "EF", "diff_coef": {"u": 1e-3, "b": 2e-3}}, 3),
    "linear_closed": ("linear", "closed", (8, 8, 16), (1.0, 1.0, 2.0), {}, 3),
    "les_closed_tracer": ("LES", "closed", (16, 8, 8), (2.0, 1.0, 1.0), {"n_tracers": 1}, 3),
    # the plume set-up of experiments/forced_convection/forced_plume.py: rotating frame + user forcing object
    "les_closed_forced_rot": ("LES", "closed", (8, 8, 16), (1.0, 1.0, 2.0),
                              {"forced": True, "rotating": True, "coriolis": 1.0, "forcing": "plume"}, 3),
    "les_closed_lfam3_visc": ("LES", "closed", (8, 8, 16), (1.0, 1.0, 2.0), {"diff_coef": {"u": 5e-3, "b": 2e-3}}, 6),
}

SCALARS = ("b", "p", "ke", "div")
VECTORS = ("u", "U", "vor")


def tracers(name):
    """Nicknames of the passive tracers of a case."""
    return ["t%d" % i for i in range(CASES[name][4].get("n_tracers", 0))]


def flat_param(name):
    modelname, geometry, (nx, ny, nz), (Lx, Ly, Lz), extra, nsteps = CASES[name]
    kw = dict(modelname=modelname, cfl=0.8, dt_max=0.05)
    kw.update({k: v for k, v in extra.items() if k != "forcing"})
    return dict(nx=nx, ny=ny, nz=nz, geometry=geometry, Lx=Lx, Ly=Ly, Lz=Lz, **kw), nsteps


class PlumeForcing(object):
    """The user forcing object of experiments/forced_convection/forced_plume.py:68-80 (a Gaussian column heat
    source added to db), written against the reference's grid / state API so that the same class drives the
    reference's drivers (oracle/gen_golden.py), the oracle and nyles_b200."""

    def __init__(self, param, grid):
        def coord(a):              # Scalar-like in the reference and in nyles_b200, a plain (k,j,i) array in the oracle
            return a if isinstance(a, np.ndarray) else np.asarray(a.view("i"))
        x = coord(grid.x_b) / param["Lx"] - 0.5
        y = coord(grid.y_b) / param["Ly"] - 0.5
        z = coord(grid.z_b) / param["Lz"]
        d = np.sqrt(x ** 2 + y ** 2)
        msk = 0.5 * (1. - np.tanh(d / 0.1))
        self.Q = 1e-1 * np.exp(-z / 0.02) * msk

    def add(self, state, dstate, time):
        db = dstate.b.view("i")
        db += self.Q


class DevicePlumeForcing(PlumeForcing):
    """The same forcing, additionally offering nyles_b200's `device_tendencies` protocol (the arrays `add` would
    add to dstate, resident on the GPU), which keeps the model on its fused RHS + time-scheme launches."""

    def device_tendencies(self, state, time):
        import torch
        if getattr(self, "_Qd", None) is None:
            self._Qd = torch.as_tensor(np.ascontiguousarray(self.Q), dtype=torch.float64).to(state.b.tensor.device)
        return {"b": self._Qd}


def forcing_of(name, param, grid, device=False):
    kind = CASES[name][4].get("forcing")
    if kind != "plume":
        return None
    return DevicePlumeForcing(param, grid) if device else PlumeForcing(param, grid)
