"""History output, API of core/nylesIO.py:86-620 (NylesIO.init / write / finalize / save_array_3D,
t_next_hist, hist_path, output_directory, script_path).

File layout is the reference's (nylesIO.py:365-585): one netCDF file per rank
``<expname>_<rank>_hist.nc`` whose global attributes are the experiment parameters, an unlimited
dimension ``t``, the 1-D coordinates (``x``/``y``/``z`` with simplified_grid, else one triple per
staggering point: ``x_b .. z_vor_k``), the step counter ``n``, the time ``t`` and one record variable
``(t, z, y, x)`` per model field, named as in the reference (b, p, u/v/w, U/V/W, vor_i ..).
The reference writes them through the netCDF4 library; that package is not in this image, so the
files are written as netCDF-3 (64-bit offset) with scipy.io.netcdf_file -- same names, dimensions,
attributes and values, readable by the same tools (ncdump, xarray, netCDF4, tools/join.py's reader).
Without scipy the snapshots go to ``.npz`` files with the same cadence.

Fields are fetched from the GPU one at a time (interior slice, device -> pinned host) only when a
snapshot is due.  Output is disabled entirely with param["datadir"] = "" (benchmarks).
"""
import os
import shutil

import numpy as np

from . import mpitools

try:
    from scipy.io import netcdf_file as _netcdf_file
except Exception:                                   # pragma: no cover - scipy is part of the image
    _netcdf_file = None

MAX_LENGTH_ATTRIBUTE = 100
_POINTS = ["b", "u", "v", "w", "vor_i", "vor_j", "vor_k"]


def _attribute(value):
    """nylesIO.py:136-160: what can be stored as a global attribute, and how."""
    if isinstance(value, bool):
        return str(value)
    if isinstance(value, (int, float, str)):
        return value
    text = "{}: {}".format(type(value), repr(value))[:MAX_LENGTH_ATTRIBUTE + 1]
    if len(text) > MAX_LENGTH_ATTRIBUTE:
        text = text[:MAX_LENGTH_ATTRIBUTE - 3] + "..."
    return text


class NylesIO(object):
    def __init__(self, param):
        self.myrank = param.get("myrank", 0)
        self.enabled = bool(param["datadir"])
        self.disk_limit = param.get("disk_space_warning", 0.0)
        self.simplified_grid = param.get("simplified_grid", False)
        self.variables_in_history = param["variables_in_history"]
        self.dt_hist = param["timestep_history"]
        self.include_halo = param["include_halo"]
        self.unitT = param.get("unit_duration", "s")
        self.unitL = param.get("unit_length", "m")
        self.n_tracers = param.get("n_tracers", 0)
        self.unit = self.generate_units
        self.t_next_hist = 0.0
        self.n_hist = 0
        self.last_saved_frame = None
        self.output_directory = None
        self.hist_path = None
        self.script_path = None
        self.hist_variables = {}
        self.format = "netcdf" if _netcdf_file is not None else "npz"
        if not self.enabled:
            return
        self.experiment_parameters = {k: _attribute(v) for k, v in param.items()}
        datadir = os.path.expanduser(param["datadir"])
        expname = param["expname"]
        out_dir = os.path.join(datadir, expname)
        mode = param["mode"]
        if mode == "overwrite":
            pass
        elif mode == "count":                                   # nylesIO.py:170-178
            counter = 0
            full = "{}_{:04d}".format(expname, counter)
            while os.path.exists(os.path.join(datadir, full)):
                counter += 1
                full = "{}_{:04d}".format(expname, counter)
            # every rank must pick the same directory: rank 0 decides
            counter = int(mpitools.global_max(float(counter) if self.myrank == 0 else 0.0))
            expname = "{}_{:04d}".format(expname, counter)
            out_dir = os.path.join(datadir, expname)
        elif mode == "continue":
            raise NotImplementedError("mode 'continue' is not implemented (neither in the reference, nylesIO.py:185-193)")
        else:
            raise ValueError("unknown mode: {}".format(mode))
        ext = "_hist.nc" if self.format == "netcdf" else "_hist"
        self.hist_path = os.path.join(out_dir, expname + "_%02i" % self.myrank + ext)
        self.script_path = os.path.join(out_dir, expname + ".py")
        self.output_directory = out_dir
        if self.myrank == 0:
            os.makedirs(out_dir, exist_ok=True)
        mpitools.barrier()

    # ------------------------------------------------------------------ helpers
    def generate_units(self, dimensions):
        """nylesIO.py:612-618: 'L.T^-2' -> 'm s-2'."""
        units = dimensions
        for k, v in {"T": self.unitT, "L": self.unitL, "^": "", ".": " "}.items():
            units = units.replace(k, v)
        return units

    def _select_variables(self, state):
        """nylesIO.py:216-250."""
        v = self.variables_in_history
        if v == "all":
            names = list(state.toc.keys())
        elif v == "prognostic":
            names = list(state.get_prognostic_variables())
        elif v == "p+p":
            names = list(state.get_prognostic_variables())
            if "p" not in state.toc:
                raise ValueError("pressure is not a variable of this model; please change the parameter "
                                 "variables_in_history.")
            names.append("p")
        else:
            names = list(v)
            if any(name not in state.toc for name in names):
                raise ValueError("a variable chosen for the history file does not exist in this model.  "
                                 "The available variables are " + str(state.toc))
            for i in range(self.n_tracers):
                nick = "t{}".format(i)
                if nick not in names:
                    names.append(nick)
        if len(names) == 0:
            raise ValueError("no variables are selected for the history file.  This means, all calculation "
                             "results are lost.  I refuse to waste energy like this.")
        return names

    @staticmethod
    def _host(field, idx):
        """Interior (or full) block of a field as a NumPy array, (z, y, x) order."""
        t = field.tensor[idx]
        return t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)

    # ------------------------------------------------------------------ public
    def init(self, state, grid, t=0.0, n=0):
        self.domainindices = state.b.domainindices
        if not self.enabled:
            return
        names = self._select_variables(state)
        self.variables_in_history = names
        variables = [state.get(v) for v in names]
        mpitools.barrier()
        assert os.path.isdir(self.output_directory)
        self.create_history_file(grid, variables)
        self.write_history_file(state, t, n)
        self.t_next_hist = t + self.dt_hist

    def create_history_file(self, grid, variables):
        """nylesIO.py:365-560."""
        k0, k1, j0, j1, i0, i1 = self.domainindices
        self.idx = {"x": slice(i0, i1), "y": slice(j0, j1), "z": slice(k0, k1)}
        sizes = {}
        for x, i in zip("xyz", "ijk"):
            if self.include_halo:
                sizes[x] = grid.size[i]
                self.idx[x] = slice(None)
            else:
                sizes[x] = getattr(grid, "n" + x)
        # scalar -> history name; vector component -> history name   (nylesIO.py:512-560)
        self.hist_variables = {}
        layout = []                       # (hist_name, nickname, point, long_name, dimension)
        for var in variables:
            nature = var.get_nature()
            if nature == "scalar":
                layout.append((var.nickname, var.nickname, "b", var.name, var.dimension))
            elif nature == "velocity":
                if var.nickname not in ("u", "U"):
                    raise NotImplementedError("unknown kind of velocity: " + var.nickname)
                mod = str.lower if var.nickname == "u" else str.upper
                for i, u in zip("ijk", "uvw"):
                    layout.append((mod(u), var[i].nickname, u, var[i].name, var[i].dimension))
            elif nature == "vorticity":
                for i in "ijk":
                    layout.append((var[i].nickname, var[i].nickname, "vor_" + i, var[i].name, var[i].dimension))
            else:
                raise ValueError("unknown nature", nature, "of variable", var.nickname)
        for hist_name, nickname, _, _, _ in layout:
            self.hist_variables[hist_name] = nickname
        if self.format != "netcdf":
            return
        f = _netcdf_file(self.hist_path, "w", mmap=False, version=2)
        for key, value in self.experiment_parameters.items():
            # not setattr: parameter names such as "mode" would shadow attributes of the file object itself
            f._attributes[key.replace(" ", "_")] = value
        f.createDimension("t", None)
        coords = {"x": "x_%s_1D", "y": "y_%s_1D", "z": "z_%s_1D"}
        where = {"b": "cell centers", "u": "cell faces x-component", "v": "cell faces y-component",
                 "w": "cell faces z-component", "vor_i": "cell edges x-component",
                 "vor_j": "cell edges y-component", "vor_k": "cell edges z-component"}
        if self.simplified_grid:
            for x in "xyz":
                f.createDimension(x, sizes[x])
        else:
            for x in "xyz":
                for p in _POINTS:
                    f.createDimension("{}_{}".format(x, p), sizes[x])
        v = f.createVariable("n", "i", ("t",))
        v.long_name = "integration step in the model run"
        v = f.createVariable("t", "d", ("t",))
        v.long_name = "time in the model run"
        v.units = self.unit("T")
        if self.simplified_grid:
            for x in "xyz":
                v = f.createVariable(x, "d", (x,))
                v.long_name = "{} at cell centers".format(x)
                v.units = self.unit("L")
                v[:] = getattr(grid, coords[x] % "b")[self.idx[x]]
        else:
            for p in _POINTS:
                for x in "xyz":
                    name = "{}_{}".format(x, p)
                    v = f.createVariable(name, "d", (name,))
                    v.long_name = "{} at {}".format(x, where[p])
                    v.units = self.unit("L")
                    v[:] = getattr(grid, coords[x] % p)[self.idx[x]]
        for hist_name, nickname, p, long_name, dimension in layout:
            dims = ("t", "z", "y", "x") if self.simplified_grid else ("t", "z_" + p, "y_" + p, "x_" + p)
            v = f.createVariable(hist_name, "d", dims)
            v.long_name = long_name
            v.units = self.unit(dimension)
        handles = dict(f.variables)                 # close() forgets them; their header positions are set by it
        f.close()
        # The header and the coordinates are on disk with zero records.  Records are appended by hand
        # (netCDF classic record layout: per record, one slab of every record variable in definition
        # order, big-endian, each padded to 4 bytes) and `numrecs` in the header is patched, so a
        # snapshot costs one pass over its own bytes however long the history already is.
        block = sizes["z"] * sizes["y"] * sizes["x"]
        self._records = [("n", ">i4", 1), ("t", ">f8", 1)] + [(h, ">f8", block) for h, _, _, _, _ in layout]
        self._record_start = os.path.getsize(self.hist_path)
        # a writer that has seen no record leaves vsize = 0 and a common `begin` for the record variables:
        # put the per-record slab size and the offset of each variable inside the first record there
        offset = self._record_start
        with open(self.hist_path, "r+b") as fp:
            for name, dtype, count in self._records:
                vsize = np.dtype(dtype).itemsize * count
                vsize += -vsize % 4
                fp.seek(handles[name].__dict__["_begin"] - 4)
                fp.write(np.array([min(vsize, 0xFFFFFFFF)], dtype=">u4").tobytes())
                fp.write(np.array([offset], dtype=">i8").tobytes())
                offset += vsize
        self._record_bytes = offset - self._record_start

    def write_history_file(self, state, t, n):
        """Append the state as record n_hist (nylesIO.py:562-580)."""
        idx = (self.idx["z"], self.idx["y"], self.idx["x"])
        if self.format == "netcdf":
            with open(self.hist_path, "r+b") as f:
                f.seek(self._record_start + self.n_hist * self._record_bytes)
                for name, dtype, count in self._records:
                    if name == "n":
                        data = np.array([n])
                    elif name == "t":
                        data = np.array([t])
                    else:
                        data = self._host(state.get(self.hist_variables[name]), idx)
                        assert data.size == count
                    f.write(np.ascontiguousarray(data, dtype=dtype).tobytes())
                f.seek(4)
                f.write(np.array([self.n_hist + 1], dtype=">i4").tobytes())      # numrecs
        else:
            out = {"t": t, "n": n}
            for hist_name, nickname in self.hist_variables.items():
                out[hist_name] = self._host(state.get(nickname), idx)
            np.savez("%s_%05i.npz" % (self.hist_path, self.n_hist), **out)
        self.n_hist += 1
        self.last_saved_frame = n

    def write(self, state, t, n):
        """Write a snapshot when t has reached the next history time; returns the stop flag
        (nylesIO.py:262-335; the interactive low-disk-space prompt becomes a one-line warning)."""
        if not self.enabled or t < self.t_next_hist:
            return False
        self.write_history_file(state, t, n)
        self.t_next_hist += self.dt_hist
        if self.disk_limit > 0:
            try:
                free = self.get_disk_space_in_GB()
            except Exception as e:                                  # noqa: BLE001
                print("\nCannot determine available disk space (%s); disabling further checks." % e)
                self.disk_limit = 0
                return False
            if free < self.disk_limit:
                print("\nWarning, low disk space: %.2f GB remaining in %s" % (free, self.output_directory))
                self.disk_limit = 0
        return False

    def finalize(self, state, t, n):
        if not self.enabled:
            return
        if n != self.last_saved_frame:
            self.write_history_file(state, t, n)

    def save_array_3D(self, data, name, description=""):
        """A 3-D array in [z, y, x] convention saved next to the history (nylesIO.py:337-356)."""
        if not self.enabled:
            return
        data = np.asarray(data.tensor.detach().cpu().numpy() if hasattr(data, "tensor") else data)
        if self.format == "netcdf":
            # the history file's header is final once records exist: extra arrays get a file of their own
            base = self.hist_path[:-len("_hist.nc")]
            f = _netcdf_file("%s_%s.nc" % (base, name), "w", mmap=False, version=2)
            for x, size in zip("zyx", data.shape):
                f.createDimension(x, size)
            v = f.createVariable(name, "d", ("z", "y", "x"))
            if description:
                v.long_name = description
            v[:] = data
            f.close()
        else:
            np.save(os.path.join(self.output_directory, name + ".npy"), data)

    def get_disk_space_in_GB(self):
        statvfs = os.statvfs(self.output_directory)
        return statvfs.f_frsize * statvfs.f_bavail / 1e9

    def backup_scriptfile(self, filename):
        if self.enabled and filename and os.path.isfile(filename):
            shutil.copyfile(filename, self.script_path)

    def write_githashnumber(self):
        return None
