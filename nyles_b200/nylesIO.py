"""History output, interface of core/nylesIO.py as used by Nyles.run() (init / write / finalize,
t_next_hist, hist_path, output_directory).

The reference writes netCDF4 per rank; netCDF4 is not part of this build's scope (SURVEY.md 2,
row 14), so snapshots go to NumPy .npz files with the same cadence and variable selection.
Output is disabled entirely with param["IO"]["datadir"] = "" (benchmarks).
"""
import os
import shutil

import numpy as np

from . import mpitools


class NylesIO(object):
    def __init__(self, param):
        self.myrank = param.get("myrank", 0)
        self.enabled = bool(param["datadir"])
        self.dt_hist = param["timestep_history"]
        self.include_halo = param["include_halo"]
        self.variables = param["variables_in_history"]
        self.t_next_hist = 0.0
        self.n_hist = 0
        self.output_directory = None
        self.hist_path = None
        self.script_path = None
        if not self.enabled:
            return
        root = os.path.expanduser(param["datadir"])
        expname = param["expname"]
        out = os.path.join(root, expname)
        if param["mode"] == "count":
            n = 0
            while os.path.isdir("%s_%02i" % (out, n)):
                n += 1
            out = "%s_%02i" % (out, n)
        elif param["mode"] == "continue":
            raise NotImplementedError("mode 'continue' is not implemented (neither in the reference, nylesIO.py:185-193)")
        if self.myrank == 0:
            os.makedirs(out, exist_ok=True)
        mpitools.barrier()
        self.output_directory = out
        self.hist_path = os.path.join(out, "%s_%02i_hist" % (expname, self.myrank))
        self.script_path = os.path.join(out, expname + ".py")

    def _names(self, state):
        v = self.variables
        if v == "all":
            return list(state.toc)
        if v == "prognostic":
            return state.get_prognostic_variables()
        if v == "p+p":
            return state.get_prognostic_variables() + ["p"]
        return list(v)

    def init(self, state, grid, t, n):
        self.domainindices = state.b.domainindices
        if self.enabled:
            self.write(state, t, n)

    def backup_scriptfile(self, filename):
        if self.enabled and filename and os.path.isfile(filename):
            shutil.copyfile(filename, self.script_path)

    def write_githashnumber(self):
        return None

    def write(self, state, t, n):
        """Write a snapshot when t has reached the next history time.  Returns the stop flag."""
        if not self.enabled or t < self.t_next_hist:
            return False
        k0, k1, j0, j1, i0, i1 = self.domainindices
        sl = (slice(None),) * 3 if self.include_halo else (slice(k0, k1), slice(j0, j1), slice(i0, i1))
        out = {"t": t, "n": n}
        for name in self._names(state):
            if state.toc[name] == "scalar":
                out[name] = state.get(name).tensor[sl].cpu().numpy()
            else:
                for d in "ijk":
                    out["%s_%s" % (name, d)] = state.get(name)[d].tensor[sl].cpu().numpy()
        np.savez("%s_%05i.npz" % (self.hist_path, self.n_hist), **out)
        self.n_hist += 1
        self.t_next_hist += self.dt_hist
        return False

    def finalize(self, state, t, n):
        return None
