#!/usr/bin/env python
"""Join the per-rank history files of one experiment into a single global file for ONE variable
(the job of the reference's tools/join.py:36-156, for the files nyles_b200/nylesIO.py writes).

    python tools/join.py <output directory of the experiment> <variable>      -> <directory>/<variable>.nc

nyles_b200 cuts the domain into slabs along z only, so rank r owns global planes [r*nz, (r+1)*nz);
halo points stored with include_halo are dropped.  Output: dimensions t, z, y, x, the variable as
(t, z, y, x) in single precision like the reference's joined file, and the three coordinate axes.
"""
import glob
import os
import re
import sys

import numpy as np
from scipy.io import netcdf_file

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nyles_b200 import topology as topo  # noqa: E402

_POINT = {"u": "u", "v": "v", "w": "w", "U": "u", "V": "v", "W": "w", "vor_i": "vor_i", "vor_j": "vor_j", "vor_k": "vor_k"}


def _open(path):
    f = netcdf_file(path, "r", mmap=False)
    f.__dict__["mode"] = "r"          # the experiment parameter "mode" is an attribute of the file
    return f


def read_param(path):
    """Global attributes back to Python values (tools/join.py:15-33)."""
    f = _open(path)
    param = {}
    for key, val in f._attributes.items():
        if isinstance(val, bytes):
            val = val.decode()
            if val in ("True", "False"):
                val = val == "True"
            elif val.startswith("<class 'list'>:"):
                body = val.split(">:", 1)[1].strip().strip("[]")
                items = [e.strip() for e in body.split(",") if e.strip() and e.strip() != "..."]
                val = [int(e) if re.fullmatch(r"-?\d+", e) else e.strip("'") for e in items]
        elif isinstance(val, np.ndarray) and val.size == 1:
            val = val.item()
        param[key] = val
    f.close()
    return param


def join(directory, varname, out=None):
    files = sorted(glob.glob(os.path.join(directory, "*_hist.nc")))
    if not files:
        raise SystemExit("no *_hist.nc file in " + directory)
    param = read_param(files[0])
    procs, nh = param["procs"], param["nh"]
    nx, ny, nz = param["nx"], param["ny"], param["nz"]
    if procs[1] != 1 or procs[2] != 1 or len(files) != procs[0]:
        raise SystemExit("expected %d z-slab files, found %d" % (procs[0], len(files)))
    halo = bool(param["include_halo"])
    simple = bool(param["simplified_grid"])
    point = "b" if varname not in _POINT else _POINT[varname]
    axis = {x: x if simple else "%s_%s" % (x, point) for x in "xyz"}
    out = out or os.path.join(directory, varname + ".nc")
    first = _open(files[0])
    if varname not in first.variables:
        raise SystemExit("call join.py with one of: " + ", ".join(k for k, v in first.variables.items() if len(v.shape) == 4))
    nt = first.variables["t"].shape[0]
    g = netcdf_file(out, "w", mmap=False, version=2)
    g.createDimension("t", nt)
    g.createDimension("x", param["global_nx"])
    g.createDimension("y", param["global_ny"])
    g.createDimension("z", param["global_nz"])
    v = g.createVariable(varname, "f", ("t", "z", "y", "x"))
    v.long_name = first.variables[varname].long_name
    v.units = first.variables[varname].units
    tv = g.createVariable("t", "f", ("t",))
    tv.long_name = first.variables["t"].long_name
    tv.units = first.variables["t"].units
    tv[:] = first.variables["t"][:]
    coords = {}
    for x in "xyz":
        c = g.createVariable(x, "f", (x,))
        c.long_name = first.variables[axis[x]].long_name
        c.units = first.variables[axis[x]].units
        coords[x] = c
    first.close()
    for k, path in enumerate(files):
        loc = [k, 0, 0]
        ngs = topo.get_neighbours(loc, procs, topo=param["geometry"])
        _, (k0, k1, j0, j1, i0, i1) = topo.get_variable_shape([nz, ny, nx], ngs, nh)
        if not halo:
            k0, k1, j0, j1, i0, i1 = 0, nz, 0, ny, 0, nx
        f = _open(path)
        assert f.variables["t"].shape[0] == nt, "rank files hold different numbers of snapshots"
        ka, kb = k * nz, (k + 1) * nz
        v[:, ka:kb, :, :] = f.variables[varname][:, k0:k1, j0:j1, i0:i1]
        coords["z"][ka:kb] = f.variables[axis["z"]][k0:k1]
        if k == 0:
            coords["y"][:] = f.variables[axis["y"]][j0:j1]
            coords["x"][:] = f.variables[axis["x"]][i0:i1]
        f.close()
    g.close()
    return out


if __name__ == "__main__":
    if len(sys.argv) != 3:
        raise SystemExit(__doc__)
    print("%s has been joined into '%s'" % (sys.argv[2], join(sys.argv[1], sys.argv[2])))
