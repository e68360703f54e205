"""Grid metrics and coordinates, API of core/grid.py:13-160.

Coordinates stay on the host (experiment scripts do NumPy math on them).  The reference
materialises 21 full 3-D coordinate arrays; here they are zero-stride broadcasts of the 1-D
axes, created on demand, so a 512^3 grid costs kilobytes instead of ~20 GB.
"""
import numpy as np

from .variables import Scalar  # noqa: F401  (re-exported for API parity)
from . import topology as topo

_AXES = {"i": (0, 1, 2), "j": (2, 0, 1), "k": (1, 2, 0)}


class HostField(object):
    """Read-only coordinate field with the Scalar.view/flipview interface."""

    def __init__(self, z1, y1, x1, which):
        self._axes1d = (z1, y1, x1)
        self._which = which
        self.size = {"i": len(x1), "j": len(y1), "k": len(z1)}

    def _full(self):
        z1, y1, x1 = self._axes1d
        shape = (len(z1), len(y1), len(x1))
        src = {"z": z1[:, None, None], "y": y1[None, :, None], "x": x1[None, None, :]}[self._which]
        return np.broadcast_to(src, shape)

    def view(self, idx=None):
        return self._full().transpose(_AXES[idx or "i"])

    def flipview(self, idx):
        return self.view({"i": "j", "j": "k", "k": "i"}[idx])


class HostVector(dict):
    pass


class Grid(object):
    def __init__(self, param):
        self.nx, self.ny, self.nz = param["nx"], param["ny"], param["nz"]
        self.npx, self.npy, self.npz = param["npx"], param["npy"], param["npz"]
        self.Lx, self.Ly, self.Lz = param["Lx"], param["Ly"], param["Lz"]
        self.dx = self.Lx / (self.npx * self.nx)
        self.dy = self.Ly / (self.npy * self.ny)
        self.dz = self.Lz / (self.npz * self.nz)
        self.dx2, self.dy2, self.dz2 = self.dx ** 2, self.dy ** 2, self.dz ** 2
        self.idx, self.idy, self.idz = 1 / self.dx, 1 / self.dy, 1 / self.dz
        self.idx2, self.idy2, self.idz2 = 1 / self.dx ** 2, 1 / self.dy ** 2, 1 / self.dz ** 2
        self.vol = self.dx * self.dy * self.dz
        self.ids2 = {"i": self.idx2, "j": self.idy2, "k": self.idz2}
        self.vol_per_ds2 = {"i": self.vol / self.dx2, "j": self.vol / self.dy2, "k": self.vol / self.dz2}

        size, self.domainindices = topo.get_variable_shape(
            [self.nz, self.ny, self.nx], param["neighbours"], param["nh"])
        self.size = {"i": size[2], "j": size[1], "k": size[0]}
        k0, k1, j0, j1, i0, i1 = self.domainindices
        loc = param.get("loc", [0, 0, 0])
        x0, y0, z0 = loc[2] * self.nx * self.dx, loc[1] * self.ny * self.dy, loc[0] * self.nz * self.dz
        self.x_b_1D = (np.arange(size[2]) + 0.5 - i0) * self.dx + x0
        self.y_b_1D = (np.arange(size[1]) + 0.5 - j0) * self.dy + y0
        self.z_b_1D = (np.arange(size[0]) + 0.5 - k0) * self.dz + z0
        hx, hy, hz = self.dx / 2, self.dy / 2, self.dz / 2
        # staggering offsets (x,y,z) of each location, grid.py:84-135
        self._offsets = {
            "b": (0, 0, 0), "vel_i": (hx, 0, 0), "vel_j": (0, hy, 0), "vel_k": (0, 0, hz),
            "vor_i": (0, hy, hz), "vor_j": (hx, 0, hz), "vor_k": (hx, hy, 0),
        }
        for name, (ox, oy, oz) in self._offsets.items():
            tag = {"b": "b", "vel_i": "u", "vel_j": "v", "vel_k": "w"}.get(name, name)
            setattr(self, "x_%s_1D" % tag, self.x_b_1D + ox)
            setattr(self, "y_%s_1D" % tag, self.y_b_1D + oy)
            setattr(self, "z_%s_1D" % tag, self.z_b_1D + oz)

    def _field(self, loc, which):
        ox, oy, oz = self._offsets[loc]
        return HostField(self.z_b_1D + oz, self.y_b_1D + oy, self.x_b_1D + ox, which)

    def _vector(self, kind, which):
        v = HostVector()
        for d in "ijk":
            v[d] = self._field("%s_%s" % (kind, d), which)
        return v

    def __getattr__(self, name):
        # x_b, y_b, z_b ; x_vel, y_vel, z_vel ; x_vor, y_vor, z_vor   (grid.py:51-160)
        if len(name) >= 3 and name[0] in "xyz" and name[1] == "_":
            tail = name[2:]
            if tail == "b":
                return self._field("b", name[0])
            if tail in ("vel", "vor"):
                return self._vector(tail, name[0])
        raise AttributeError(name)
