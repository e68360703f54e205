"""The Nyles main class, API of core/nyles.py:22-260: Nyles(user_param), .run(), .compute_dt().

One process per GPU (torchrun); the process grid is [npz, 1, 1] -- slabs along z.
"""
import pickle
import sys
from time import time

import numpy as np
import torch

from . import grid as grid_module
from . import lib
from . import model_les
from . import model_les_euler
from . import mpitools
from .online_diag import VFwork
from . import nylesIO
from . import timing
from . import topology as topo


class Nyles(object):
    def __init__(self, user_param):
        user_param.check()
        user_param.freeze()
        param = user_param.view_parameters()
        self.param = param
        npx, npy, npz = param["npx"], param["npy"], param["npz"]
        if npx != 1 or npy != 1:
            raise NotImplementedError("nyles_b200 decomposes along z only: set npx = npy = 1, npz = #GPUs")
        param["nx"] = param["global_nx"] // npx
        param["ny"] = param["global_ny"] // npy
        param["nz"] = param["global_nz"] // npz

        topo.topology = param["geometry"]
        procs = [npz, npy, npx]
        myrank = mpitools.get_myrank(procs)
        loc = topo.rank2loc(myrank, procs)
        param["procs"] = procs
        param["myrank"] = myrank
        param["neighbours"] = topo.get_neighbours(loc, procs)
        param["loc"] = loc
        self.myrank = myrank
        if myrank == 0 and param.get("verbose", True):
            self.banner()

        self.grid = grid_module.Grid(param)
        self.IO = nylesIO.NylesIO(param)
        if myrank == 0 and self.IO.enabled:
            with open(self.IO.output_directory + "/param.pkl", "wb") as fid:
                pickle.dump({k: v for k, v in param.items()}, fid)
        self.initiate(param)

    def initiate(self, param):
        name = param["modelname"]
        if name == "LES":
            self.model = model_les.LES(param, self.grid)
        elif name == "Euler3d":
            self.model = model_les_euler.LES(param, self.grid)
        elif name == "linear":
            self.model = model_les.LES(param, self.grid, linear=True)
        else:
            raise NotImplementedError("modelname %r is outside the LES hot path of nyles_b200" % name)
        self.tend = param["tend"]
        self.auto_dt = param["auto_dt"]
        self.dt0 = param["dt"]
        self.cfl = param["cfl"]
        self.dt_max = param["dt_max"]
        self.plotting = None
        self.gridcellpersubdom = param["nx"] * param["ny"] * param["nz"]
        self.diag = VFwork(self.model, self.grid)                      # core/nyles.py:102

    def run(self, max_steps=None, quiet=False):
        t, n = 0.0, 0
        self.model.diagnose_var(self.model.state)
        self.IO.init(self.model.state, self.grid, t, n)
        if self.myrank == 0 and self.IO.enabled:
            self.IO.backup_scriptfile(sys.argv[0])
        time_length = len(str(int(self.tend))) + 3
        time_string = "\r" + ", ".join([
            "n = {:3d}", "t = {:" + str(time_length) + ".2f}/{:" + str(time_length) + ".2f}",
            "dt = {:.4f}", "perf = {:.2e}"])
        stop = False
        realtime0 = time()
        while not stop:
            dt = self.compute_dt()
            blowup = self.model.forward(t, dt)
            t += dt
            n += 1
            if self.IO.enabled and t >= self.IO.t_next_hist and self.model.nonlinear:      # core/nyles.py:167-171
                self.diag.compute()
                Kdiss = mpitools.global_sum(self.diag.worksum)
                if self.myrank == 0 and not quiet:
                    print("\nKdiss = {}".format(Kdiss))
            stop = self.IO.write(self.model.state, t, n)
            if self.myrank == 0 and not quiet:
                realtime = time()
                # wall time per iteration per grid cell of the subdomain (nyles.py:174-182)
                perf = (realtime - realtime0) / self.gridcellpersubdom
                realtime0 = realtime
                print(time_string.format(n, t, self.tend, dt, perf), end="")
            localmood = 1. if (blowup or not np.isfinite(dt)) else 0.
            if mpitools.global_sum(localmood) > 0.:
                if self.myrank == 0:
                    print("\nBLOW UP!")
                self.IO.t_next_hist = t
                self.IO.write(self.model.state, t, n)
                stop = True
            if t >= self.tend or stop or (max_steps is not None and n >= max_steps):
                break
        if self.myrank == 0 and not quiet:
            print()
            print("Job is aborted" if stop else "Job completed as expected")
        self.IO.finalize(self.model.state, t, n)
        if self.IO.enabled:
            self.model.write_stats(self.IO.output_directory)
            timing.write_timings(self.IO.output_directory)
        self.t, self.n = t, n

    def compute_dt(self):
        """dt = min(cfl / max|U|, dt_max) from the contravariant velocity (nyles.py:227-260)."""
        if not self.auto_dt:
            return self.dt0
        U = self.model.state.U
        t = U["i"].tensor
        umax2 = self.model.cached_max_speed2()            # reduced while diagnose_var wrote U, if still valid
        if umax2 is None:
            out = lib.C.c_double()
            lib.check(lib.load().ny_max_speed2(lib.context(t.device), lib.ptr(U["i"].tensor), lib.ptr(U["j"].tensor),
                                               lib.ptr(U["k"].tensor), t.numel(), lib.C.byref(out), lib.stream()))
            umax2 = out.value
        U_max = mpitools.global_max(float(np.sqrt(umax2)))
        if U_max == 0.0:
            return self.dt_max
        return min(self.cfl / U_max, self.dt_max)

    # ------------------------------------------------------------------ host-buffer interface
    def prognostic_tensors(self):
        st = self.model.state
        return [st.get(name).tensor for name in st.get_prognostic_scalars()]

    def allocate_host_state(self):
        """Pinned host arrays (canonical (k,j,i) order) for the prognostic fields b, u_i, u_j, u_k, ...:
        the layout in which a caller of the reference's f2py/ctypes boundary holds its NumPy state."""
        return [torch.empty(t.shape, dtype=t.dtype, device="cpu", pin_memory=True) for t in self.prognostic_tensors()]

    def step_host(self, t, host_state, modified=False):
        """One model step on HOST buffers: upload the prognostic state, compute dt, step, download the
        new state into the same buffers.  Returns dt.  This is the call whose cost bench.py reports as
        `e2e`: what a user pays who keeps the state in host memory, as the reference does.

        modified=False (default): the caller promises that host_state still holds what the previous
        step_host (or the copy from prognostic_tensors()) left there.  The diagnostic fields U, vor, ke, p,
        the multigrid's warm start and the LFAM3 history on the device are then consistent with the upload
        and the step costs what a step costs.
        modified=True: the caller has edited the buffers (a restart, an assimilation step, a forcing applied on
        the host).  The upload is followed by diagnose_var -- halo fills, projection, U, vorticity, kinetic
        energy, from a zero first guess of the pressure -- exactly as Nyles.run() starts from a freshly written
        state (core/nyles.py:125), the cached
        max|U|^2 is dropped and the LFAM3 scheme restarts with its Euler start-up step
        (core/timescheme.py:131-139), because its stored time levels no longer belong to this state.
        The projection that ends a step only changes u, so on a domain without halos the other fields
        (b, passive tracers) start their way back over PCIe on a second stream as soon as the time scheme
        has written them, underneath that projection."""
        import torch
        dev = self.prognostic_tensors()
        for h, d in zip(host_state, dev):
            d.copy_(h, non_blocking=True)
        if modified:
            self.model._umax_key = None
            # a run that starts from this state has no previous pressure to warm-start from
            mg = self.model.mg
            mg.set_array(torch.zeros(mg.get_arrayshape(1), dtype=torch.float64, device=dev[0].device), ivar=1)
            self.model.diagnose_var(self.model.state)
            if hasattr(self.model.timescheme, "first"):
                self.model.timescheme.first = True
        dt = self.compute_dt()
        names = self.model.state.get_prognostic_scalars()
        early = []
        ts = self.model.timescheme
        if not self.param["neighbours"]:              # with halos the closing diagnose_var still fills them
            if getattr(self, "_copy_stream", None) is None:
                self._copy_stream = torch.cuda.Stream()
            main = torch.cuda.current_stream()

            def start_download(state):
                ready = torch.cuda.Event()
                ready.record(main)
                self._copy_stream.wait_event(ready)
                with torch.cuda.stream(self._copy_stream):
                    for name, h in zip(names, host_state):
                        if name not in ("u_i", "u_j", "u_k"):
                            h.copy_(state.get(name).tensor, non_blocking=True)
                            early.append(name)
            ts.before_last_diagnose = start_download
        try:
            self.model.forward(t, dt)
        finally:
            ts.before_last_diagnose = None
        # the fused step rotates the buffers of the prognostic fields: ask again where they live
        for name, h, d in zip(names, host_state, self.prognostic_tensors()):
            if name not in early:
                h.copy_(d, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        if early:
            self._copy_stream.synchronize()
        return dt

    def banner(self):
        print("-" * 80 + "\nnyles_b200: B200-native LES time step with the Nyles API", file=sys.stderr)
