"""User parameters, API and checks of core/parameters.py:12-247.

Experiments do ``param = UserParameters(); param.model["geometry"] = ...; Nyles(param)``.
Unknown keys and writes after the creation of Nyles raise UserParameterError.
"""
import copy
import datetime

from .defaults import DEFAULTS

CATEGORIES = ["model", "physics", "IO", "animation", "time", "discretization", "MPI", "multigrid"]
ATTRIBUTES = ["type", "default", "avail", "doc"]


class UserParameterError(Exception):
    """An error occured with the user-set parameters."""


class InextensibleDict(dict):
    """dict whose key set is fixed; `freeze()` blocks every further assignment (class-wide, as in
    the reference, parameters.py:12-41)."""

    frozen = False

    @classmethod
    def freeze(cls):
        cls.frozen = True

    @classmethod
    def unfreeze(cls):
        """Not in the reference: lets one process build several Nyles objects (tests, benchmarks)."""
        cls.frozen = False

    def __setitem__(self, key, item):
        if self.frozen:
            raise UserParameterError("not possible to modify parameters after the creation of Nyles.")
        if key not in self:
            raise UserParameterError("not possible to add new key {!r} to the parameters.".format(key))
        dict.__setitem__(self, key, item)


class UserParameters(object):
    TYPES = {"str": str, "int": int, "float": (float, int), "bool": bool, "dict": dict,
             "list or string": (list, tuple, str)}

    def __init__(self):
        self.documentations, self.options, self.types = {}, {}, {}
        for cat in CATEGORIES:
            values = {}
            for name, (typ, default, avail, doc) in DEFAULTS[cat].items():
                values[name] = copy.deepcopy(default)
                self.documentations[name], self.options[name], self.types[name] = doc, avail, typ
            setattr(self, cat, InextensibleDict(values))

    def help(self, parameter):
        if parameter in self.documentations:
            return self.documentations[parameter]
        raise ValueError("invalid parameter: {!r}".format(parameter))

    def possible_values(self, parameter):
        if parameter in self.options:
            return self.options[parameter]
        raise ValueError("invalid parameter: {!r}".format(parameter))

    def view_parameters(self):
        out = {}
        for cat in CATEGORIES:
            out.update(getattr(self, cat))
        return out

    def freeze(self):
        InextensibleDict.freeze()

    def check(self):
        """Raise UserParameterError if any value has the wrong type or range (parameters.py:149-247)."""
        powers = [2 ** n for n in range(datetime.datetime.now().year - 2000)]
        for name, value in self.view_parameters().items():
            typ = self.types[name]
            if not isinstance(value, self.TYPES[typ]) or (typ in ("int", "float") and isinstance(value, bool)):
                raise UserParameterError("parameter {} must be {} {}, not {}".format(
                    name, "an" if typ == "int" else "a", typ, type(value)))
            avail = self.options[name]
            if isinstance(avail, list):
                if value not in avail:
                    raise UserParameterError("parameter {} must be one of {}, not {!r}".format(name, avail, value))
            elif avail == "> 0.0":
                if not value > 0.0:
                    raise UserParameterError("parameter {} must be positive".format(name))
            elif avail in (">= 0.0", ">= 0"):
                if not value >= 0.0:
                    raise UserParameterError("parameter {} must be non-negative".format(name))
            elif avail == ">= 1":
                if not value >= 1:
                    raise UserParameterError("parameter {} must be at least 1".format(name))
            elif avail in ("2^n", "[3 *] 2^n"):
                ok = value in powers or (avail.startswith("[3") and value / 3 in powers)
                if not ok:
                    if value < max(powers):
                        raise UserParameterError("parameter {} must be a power of 2{}".format(
                            name, " or 3 times a power of 2" if avail.startswith("[3") else ""))
                    raise UserParameterError("parameter {} is very large".format(name))
            elif avail == "any valid filename":
                if "/" in value:
                    raise UserParameterError('parameter {} must not contain a "/"'.format(name))
            elif name == "variables_in_history":
                if not isinstance(value, (list, tuple)) and value not in ["all", "prognostic", "p+p"]:
                    raise UserParameterError("value {!r} of parameter {} not understood".format(value, name))
        for x in "xyz":
            if self.MPI["np" + x] > self.discretization["global_n" + x]:
                raise UserParameterError("parameter np{} cannot be larger than global_n{}".format(x, x))
