"""@timing decorator, API of core/timing.py:15-42 (wall time per decorated call, first 100 steps).

Kernel launches are asynchronous, so unsynchronised wall times would be meaningless; the
decorator is therefore inert unless NYLES_TIMING=1, in which case it synchronises the device
around the call (diagnostic use only, never inside a benchmark's timed region).
"""
import os
import pickle
from functools import wraps
from time import time

import torch

from . import mpitools

stats = {}
ENABLED = os.environ.get("NYLES_TIMING", "0") == "1"


def timing(f):
    if not ENABLED:
        return f

    @wraps(f)
    def wrap(*args, **kw):
        if len(stats.get("forward", ())) >= 100:
            return f(*args, **kw)
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        ts = time()
        result = f(*args, **kw)
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        stats.setdefault(f.__name__, []).append(time() - ts)
        return result
    return wrap


def write_timings(path):
    if mpitools.get_myrank() == 0 and stats:
        with open("%s/timing.pkl" % path, "bw") as fid:
            pickle.dump(stats, fid)


def analyze_timing(path):
    """The reference draws a log-log PNG here (timing.py:44-86); plotting is out of scope."""
    return None
