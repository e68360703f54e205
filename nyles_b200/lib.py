"""ctypes binding of libnyles_b200.so (include/nyles_b200.h).

This is the only place where the host layer crosses into native code; it plays the role of
the f2py extension modules (core/Makefile:1-2) and of the generated ``mgmod`` ctypes module
(core/build.py:261-284) of the reference.  There is no CPU fallback: if the shared library is
missing or a call fails, an exception is raised.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NYLES_B200_LIB") or os.path.join(_HERE, "libnyles_b200.so")

_lib = None
_ctx = {}
u_epoch = 0          # bumped by every library call that writes the contravariant velocity state.U


class NylesB200Error(RuntimeError):
    pass


class ny_ext(C.Structure):
    _fields_ = [("nz", C.c_int), ("ny", C.c_int), ("nx", C.c_int)]


class ny_mg_stats(C.Structure):
    _fields_ = [("nite", C.c_int), ("nres", C.c_int), ("res", C.c_double), ("normb", C.c_double),
                ("reshist", C.c_double * 32)]


_P = C.c_void_p      # device pointers and opaque handles
_D = C.c_double
_I = C.c_int
_LL = C.c_longlong

_PROTOS = {
    "ny_init": ([_I, C.POINTER(_P)], _I),
    "ny_free": ([_P], None),
    "ny_last_error": ([], C.c_char_p),
    "ny_version": ([], _I),
    "ny_launch_count": ([_P], _LL),
    "ny_launch_count_reset": ([_P], None),
    "ny_prof_start": ([_P, C.c_ulonglong], _I),
    "ny_prof_collect": ([_P, C.POINTER(_D), C.POINTER(_LL)], _I),
    "ny_prof_name": ([_I], C.c_char_p),
    "ny_set_arith": ([_P, _I], _I),
    "ny_get_arith": ([_P], _I),
    "ny_set_momentum_variant": ([_P, _I], _I),
    "ny_vorticity": ([_P] + [_P] * 6 + [ny_ext, _D, _P], _I),
    "ny_upwind": ([_P] + [_P] * 5 + [ny_ext, _P], _I),
    "ny_upwind_diff": ([_P] + [_P] * 5 + [_D, _D, _D, ny_ext, _P], _I),
    "ny_vortex_force": ([_P] + [_P] * 9 + [ny_ext, _P], _I),
    "ny_vortex_force_linear": ([_P] + [_P] * 9 + [_I, ny_ext, _P], _I),
    "ny_upwind_linear": ([_P] + [_P] * 5 + [_I, ny_ext, _P], _I),
    "ny_kin": ([_P] + [_P] * 4 + [_D, _D, _D, ny_ext, _P], _I),
    "ny_bernoulli": ([_P] + [_P] * 5 + [_D, _I, ny_ext, _P], _I),
    "ny_div": ([_P] + [_P] * 4 + [ny_ext, _P], _I),
    "ny_gradp": ([_P] + [_P] * 4 + [ny_ext, _P], _I),
    "ny_U_from_u": ([_P] + [_P] * 6 + [_D, _D, _D, ny_ext, _P], _I),
    "ny_add_laplacian": ([_P, _P, _P, _D, _D, _D, ny_ext, _P], _I),
    "ny_rhs": ([_P] + [_P] * 12 + [_D, _I, ny_ext, _P], _I),
    "ny_rhs_update_u": ([_P] + [_P] * 9 + [C.POINTER(_P * 3)] * 3 + [_I, _D, _D, _I, ny_ext, _P], _I),
    "ny_rhs_step": ([_P] + [_P] * 7 + [C.POINTER(_P * 4)] * 4 + [_P] + [_I, _D, _D, _I, ny_ext, _P], _I),
    "ny_ts_axpy": ([_P, _P, _P, _D, _LL, _P], _I),
    "ny_ts_lfam3_first": ([_P, _P, _P, _P, _P, _D, _LL, _P], _I),
    "ny_ts_lfam3_pred": ([_P, _P, _P, _P, _P, _D, _LL, _P], _I),
    "ny_ts_lfam3_corr": ([_P, _P, _P, _P, _D, _LL, _P], _I),
    "ny_ts_rk3_stage2": ([_P, _P, _P, _P, _D, _LL, _P], _I),
    "ny_ts_rk3_stage3": ([_P, _P, _P, _P, _P, _D, _LL, _P], _I),
    "ny_max_speed2": ([_P, _P, _P, _P, _LL, C.POINTER(_D), _P], _I),
    "ny_halo_fill_self": ([_P, _P, ny_ext, _I, C.POINTER(_I * 3), _P], _I),
    "ny_comm_unique_id": ([C.c_char_p], _I),
    "ny_comm_init": ([_P, _I, _I, C.c_char_p, C.POINTER(_P)], _I),
    "ny_comm_free": ([_P], None),
    "ny_comm_stats": ([_P, C.POINTER(_LL), C.POINTER(_LL), _I], _I),
    "ny_comm_size": ([_P], _I),
    "ny_comm_rank": ([_P], _I),
    "ny_comm_allreduce_host": ([_P, C.POINTER(_D), _I, _I, _P], _I),
    "ny_halo_exchange": ([_P, _P, C.POINTER(_P), _I, ny_ext, _I, _I, _I, _I, _I, _P], _I),
    "ny_mg_create_slab": ([_P, _P, _I, _I, _I, _I, C.POINTER(_P)], _I),
    "ny_mg_set_gather_cells": ([_LL], None),
    "ny_mg_set_overlap_cells": ([_LL], None),
    "ny_mg_set_split_tiles": ([_LL], None),
    "ny_mg_set_tail_cells": ([_LL], None),
    "ny_mg_set_wide_cells": ([_LL], None),
    "ny_mg_is_box": ([_P], _I),
    "ny_mg_set_fast_path": ([_P, _I], _I),
    "ny_mg_set_fused_legs": ([_P, _I], _I),
    "ny_mg_first_gathered_level": ([_P], _I),
    "ny_mg_create": ([_P, _I, _I, _I, _I, C.POINTER(_P)], _I),
    "ny_mg_destroy": ([_P], None),
    "ny_mg_nlevels": ([_P], _I),
    "ny_mg_shape": ([_P, _I, C.POINTER(_I * 3)], _I),
    "ny_mg_set_param": ([_P, _I, _D, _D], _I),
    "ny_mg_set_array": ([_P, _I, _I, _P, _P], _I),
    "ny_mg_get_array": ([_P, _I, _I, _P, _P], _I),
    "ny_mg_setup_operators": ([_P, _P], _I),
    "ny_mg_solve": ([_P, C.POINTER(ny_mg_stats), _P], _I),
    "ny_mg_solve_directly": ([_P, _P, _P, ny_ext, C.POINTER(_I * 3), _D, C.POINTER(ny_mg_stats), _P], _I),
    "ny_mg_project": ([_P] + [_P] * 5 + [_D, _D, _D, ny_ext, C.POINTER(_I * 3), _D, C.POINTER(ny_mg_stats), _P], _I),
    "ny_diag_post": ([_P] + [_P] * 10 + [_D, _D, _D, _D, ny_ext, _P], _I),
    "ny_mg_op": ([_P, _I, _I, _P], _I),
    "ny_diag_post_max_speed2": ([_P, C.POINTER(_D), _P], _I),
    "ny_debug_weno5": ([_P, _P, _P, _LL, _P], _I),
    "ny_debug_weno3": ([_P, _P, _P, _LL, _P], _I),
    "ny_debug_fp64_peak": ([_P, _D, C.POINTER(_D), _P], _I),
    "ny_debug_div": ([_P, _P, _P, _P, _LL, C.POINTER(_LL), _P], _I),
}

EXPORTED = sorted(_PROTOS)


def load():
    """Load the shared library (no device needed); raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NylesB200Error(
                "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C nyles_b200/csrc` (there is no CPU fallback)" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (argtypes, restype) in _PROTOS.items():
            fn = getattr(lib, name)          # AttributeError if the header and the library disagree
            fn.argtypes = argtypes
            fn.restype = restype
        _lib = lib
    return _lib


def check(status):
    if status != 0:
        raise NylesB200Error("libnyles_b200 error %d: %s" % (status, load().ny_last_error().decode()))


def context(device=None):
    """The per-GPU ny_ctx* (created on first use)."""
    if not torch.cuda.is_available():
        raise NylesB200Error("nyles_b200 needs a CUDA device (B200, sm_100a); none is visible")
    if device is None:
        device = torch.cuda.current_device()
    device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _ctx:
        h = _P()
        check(load().ny_init(idx, C.byref(h)))
        _ctx[idx] = h
    return _ctx[idx]


def stream():
    return _P(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a dense float64 CUDA tensor."""
    if t is None:
        return _P(0)
    if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
        raise NylesB200Error("expected a contiguous float64 CUDA tensor, got %s %s contiguous=%s"
                             % (t.device, t.dtype, t.is_contiguous()))
    return _P(t.data_ptr())


def ext(t):
    nz, ny, nx = t.shape
    return ny_ext(nz, ny, nx)


def launch_count():
    return sum(load().ny_launch_count(h) for h in _ctx.values())


def launch_count_reset():
    for h in _ctx.values():
        load().ny_launch_count_reset(h)


NY_PROF_NTAGS = 18


def prof_start(mask=(1 << NY_PROF_NTAGS) - 1, device=None):
    """Time the enabled kernel families with CUDA events on their launching stream."""
    check(load().ny_prof_start(context(device), mask))


def prof_collect(device=None):
    """{family: (total_ms, groups)} since prof_start (synchronises the device)."""
    ms = (_D * NY_PROF_NTAGS)()
    n = (_LL * NY_PROF_NTAGS)()
    check(load().ny_prof_collect(context(device), ms, n))
    return {load().ny_prof_name(t).decode(): (ms[t], n[t]) for t in range(NY_PROF_NTAGS)}


def set_arith(fast, device=None):
    """WENO arithmetic: False = bit-exact source order, True = re-associated smooth part (see nyles_b200.h)."""
    check(load().ny_set_arith(context(device), 1 if fast else 0))
