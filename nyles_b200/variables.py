"""Scalar / Vector / State on device memory, API of core/variables.py:52-399.

Re-design: the reference keeps three NumPy copies of every
field, one per inner direction, and transposes between them (variables.py:146-182,404-411).
Here a field is ONE contiguous float64 CUDA tensor in the canonical (k,j,i) order;
``view('j')`` and ``view('k')`` return permuted aliases with the reference's index order
((i,k,j) and (j,i,k)), so scripts that read or write through any view keep working and no
transpose is ever executed.  Views are thin proxies (`FieldView`) that accept NumPy arrays on
assignment and convert to NumPy on ``np.asarray`` -- experiment scripts do
``b[:] = np.tanh(...)`` (experiments/lockechange/lockexchange.py:65-72).
"""
from collections import namedtuple

import numpy as np
import torch

from . import topology as topo

ModelVariable = namedtuple("ModelVariable", ["type", "name", "dimension", "prognostic"])

# core/variables.py:39-48
modelvar = {
    "b": ModelVariable("scalar", "buoyancy", "L.T^-2", True),
    "p": ModelVariable("scalar", "pressure", "L^2.T^-2", False),
    "ke": ModelVariable("scalar", "kinetic energy", "L^2.T^-2", False),
    "div": ModelVariable("scalar", "divergence", "T^-1", False),
    "u": ModelVariable("velocity", "covariant velocity", "L^2.T^-1", True),
    "U": ModelVariable("velocity", "contravariant velocity", "T^-1", False),
    "vor": ModelVariable("vorticity", "vorticity", "L^2.T^-1", False),
    "work": ModelVariable("scalar", "work", "L^2.T^-2", False),
}

_AXES = {"i": (0, 1, 2), "j": (2, 0, 1), "k": (1, 2, 0)}


def default_device(param=None):
    if param is not None and param.get("device") is not None:
        return torch.device(param["device"])
    if not torch.cuda.is_available():
        raise RuntimeError("nyles_b200 fields live in GPU memory and no CUDA device is visible "
                           "(host-logic tests may pass param['device']='cpu')")
    return torch.device("cuda", torch.cuda.current_device())


class FieldView(object):
    """Alias of (a permutation of) a field tensor with NumPy-friendly assignment."""

    __array_priority__ = 100

    def __init__(self, tensor=None, owner=None, axes=None, index=()):
        # A view taken from a Scalar is bound to the Scalar, not to its current buffer: the fused time
        # step rotates the buffers of the prognostic fields (model_les.LES.rhs_step), and a view kept
        # across steps -- as NumPy views of the reference may be -- must keep showing the field.  The same
        # holds for sub-views (b.view("i")[k0:k1]): they keep the chain of indices and re-apply it.
        self._tensor, self._owner, self._axes, self._index = tensor, owner, axes, tuple(index)

    @property
    def tensor(self):
        if self._owner is not None:
            t = self._owner.tensor.permute(*self._axes)
            for idx in self._index:
                t = t[idx]
            return t
        return self._tensor

    # --- conversions
    def _coerce(self, value):
        if isinstance(value, FieldView):
            return value.tensor
        if isinstance(value, np.ndarray):
            return torch.as_tensor(np.ascontiguousarray(value), dtype=torch.float64).to(self.tensor.device)
        return value

    def __array__(self, dtype=None, copy=None):
        a = self.tensor.detach().cpu().numpy()
        return a.astype(dtype) if dtype is not None else a

    def numpy(self):
        return self.__array__()

    def copy(self):
        return self.__array__().copy()

    @property
    def shape(self):
        return tuple(self.tensor.shape)

    @property
    def ndim(self):
        return self.tensor.ndim

    # --- indexing
    def __getitem__(self, idx):
        out = self.tensor[idx]
        if out.ndim == 0:
            return out.item()
        if self._owner is not None:
            return FieldView(owner=self._owner, axes=self._axes, index=self._index + (idx,))
        return FieldView(out)

    def __setitem__(self, idx, value):
        self.tensor[idx] = self._coerce(value)

    # --- in-place arithmetic used by experiment scripts (u *= dx, db += Q, wk[..] += f)
    def __iadd__(self, v):
        self.tensor.add_(self._coerce(v)); return self

    def __isub__(self, v):
        self.tensor.sub_(self._coerce(v)); return self

    def __imul__(self, v):
        self.tensor.mul_(self._coerce(v)); return self

    def __itruediv__(self, v):
        self.tensor.div_(self._coerce(v)); return self

    # --- out-of-place arithmetic returns NumPy (host), like the reference's views
    def _np(self, v):
        return np.asarray(v) if isinstance(v, FieldView) else v

    def __add__(self, v): return self.numpy() + self._np(v)
    def __radd__(self, v): return self._np(v) + self.numpy()
    def __sub__(self, v): return self.numpy() - self._np(v)
    def __rsub__(self, v): return self._np(v) - self.numpy()
    def __mul__(self, v): return self.numpy() * self._np(v)
    def __rmul__(self, v): return self._np(v) * self.numpy()
    def __truediv__(self, v): return self.numpy() / self._np(v)
    def __rtruediv__(self, v): return self._np(v) / self.numpy()
    def __pow__(self, v): return self.numpy() ** v
    def __neg__(self): return -self.numpy()

    def sum(self): return self.tensor.sum().item()
    def max(self): return self.tensor.max().item()
    def min(self): return self.tensor.min().item()

    def __repr__(self):
        return "FieldView(shape=%s, device=%s)" % (self.shape, self.tensor.device)


class Scalar(object):
    """core/variables.py:52-224."""

    def __init__(self, param, name, nickname, dimension, prognostic=False):
        required = ["nx", "ny", "nz", "nh", "neighbours"]
        self.param = {k: param[k] for k in required}
        self.param["device"] = param.get("device")
        self.name, self.nickname, self.dimension, self.prognostic = name, nickname, dimension, prognostic
        nx, ny, nz, nh = param["nx"], param["ny"], param["nz"], param["nh"]
        shape = [nz, ny, nx]
        self.shape = shape
        size, domainindices = topo.get_variable_shape(shape, self.param["neighbours"], nh)
        nzl, nyl, nxl = size
        self.size = {"i": nxl, "j": nyl, "k": nzl}
        self.domainindices = domainindices
        k0, k1, j0, j1, i0, i1 = domainindices
        self.mg_idx = (slice(max(0, k0 - 1), min(nzl, k1 + 1)), slice(max(0, j0 - 1), min(nyl, j1 + 1)),
                       slice(max(0, i0 - 1), min(nxl, i1 + 1)))
        # the single canonical (k,j,i) buffer
        self.tensor = torch.zeros((nzl, nyl, nxl), dtype=torch.float64, device=default_device(self.param))
        self.activeview = "i"

    def duplicate(self):
        return Scalar(self.param, self.name, self.nickname, self.dimension, self.prognostic)

    def view(self, idx=None):
        """Alias of the data with idx as inner direction: 'i' -> (k,j,i), 'j' -> (i,k,j), 'k' -> (j,i,k)."""
        if idx is None:
            idx = self.activeview
        if idx not in _AXES:
            raise ValueError("argument idx of Scalar.view must be in ['i','j','k'], not %r" % (idx,))
        self.activeview = idx
        return FieldView(owner=self, axes=_AXES[idx])

    def flipview(self, idx):
        """Alias with idx as OUTER direction (variables.py:184-204)."""
        try:
            return self.view({"i": "j", "j": "k", "k": "i"}[idx])
        except KeyError:
            raise ValueError('argument idx of Scalar.flipview must be in ["i","j","k"], not ' + repr(idx))

    def viewlike(self, scalar):
        return self.view(scalar.activeview)

    @staticmethod
    def get_nature():
        return "scalar"


class Vector(dict):
    """core/variables.py:229-293."""

    def __init__(self, param, name, nickname, dimension, prognostic=False, is_velocity=True):
        dirname = {"i": "x", "j": "y", "k": "z"}
        for d in "ijk":
            self[d] = Scalar(param, name + " %s-component" % dirname[d], nickname + "_" + d, dimension, prognostic)
        self.param = self["i"].param
        self.name, self.nickname, self.dimension = name, nickname, dimension
        self.prognostic, self.is_velocity = prognostic, is_velocity

    def duplicate(self):
        return Vector(self.param, self.name, self.nickname, self.dimension, self.prognostic, self.is_velocity)

    def get_nature(self):
        return "velocity" if self.is_velocity else "vorticity"

    def tensors(self):
        return self["i"].tensor, self["j"].tensor, self["k"].tensor


class State(object):
    """core/variables.py:298-399."""

    def __init__(self, listvar):
        self.toc = {}
        for var in listvar:
            self.toc[var.nickname] = var.get_nature()
            setattr(self, var.nickname, var)

    def __str__(self):
        return "\n".join("{:10}: {!r}".format(v, getattr(self, v)) for v in self.toc)

    def duplicate_prognostic_variables(self):
        return State([getattr(self, n).duplicate() for n in self.toc if getattr(self, n).prognostic])

    def get(self, variable):
        if len(variable) > 2 and variable[-2] == "_" and variable[-1] in "ijk":
            return getattr(self, variable[:-2])[variable[-1]]
        return getattr(self, variable)

    def get_prognostic_variables(self):
        return [v for v in self.toc if getattr(self, v).prognostic]

    def get_prognostic_scalars(self):
        out = []
        for nickname in self.get_prognostic_variables():
            if self.toc[nickname] == "scalar":
                out.append(nickname)
            else:
                out += ["{}_{}".format(nickname, d) for d in "ijk"]
        return out


def get_state(param, variables=None):
    """core/variables.py:414-438."""
    listvar = []
    for nickname, var in (variables or modelvar).items():
        if var.type == "scalar":
            listvar.append(Scalar(param, var.name, nickname, var.dimension, var.prognostic))
        else:
            listvar.append(Vector(param, var.name, nickname, var.dimension, var.prognostic,
                                  is_velocity=(var.type == "velocity")))
    return State(listvar)


def get_work(param):
    return Scalar(param, "work", "w", "any", True)
