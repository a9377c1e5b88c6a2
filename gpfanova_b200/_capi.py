"""ctypes binding of include/gpfanova_b200.h -- the only way the Python layer reaches the GPU.

There is no CPU fallback: if the shared library is missing it is built with nvcc; if that fails, or no CUDA
device is present when a compute call is made, the call raises.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

OK, ERR_ARG, ERR_CUDA, ERR_UNSUPPORTED, ERR_NOT_PD = 0, 1, 2, 3, 4
STATUS_NONPOS_DIAG, STATUS_NOT_PD = -1, -2

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


class ModelDesc(C.Structure):
    _fields_ = [("n", C.c_int), ("p", C.c_int), ("m", C.c_int), ("f", C.c_int), ("P", C.c_int),
                ("chains", C.c_int), ("device", C.c_int), ("chain_offset", C.c_int), ("seed", C.c_uint64),
                ("h_x", c_dp), ("h_y", c_dp), ("h_X", c_dp), ("h_group_of_fn", c_ip), ("h_slice_w", c_dp),
                ("h_slice_m", c_ip), ("offset", C.c_double), ("prior_jitter", C.c_double),
                ("prior_lb", C.c_double), ("prior_ub", C.c_double)]


# name -> (restype, argtypes); every symbol include/gpfanova_b200.h declares
SIGNATURES = {
    "gpf_version": (C.c_int, []),
    "gpf_build_info": (C.c_char_p, []),
    "gpf_rbf_K": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "gpf_jitchol": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gpf_kinv": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gpf_rbf_dist": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "gpf_rbf_dK": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "gpf_gp_loglik": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_double,
                                C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gpf_gp_sample": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_ulonglong,
                                C.c_uint, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gpf_function_derivative": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                          C.c_void_p, C.c_ulonglong, C.c_uint, C.c_double, C.c_int, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p]),
    "gpf_compose_y": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_ulonglong, C.c_uint,
                                C.c_void_p, C.c_void_p]),
    "gpf_fanova_transform": (C.c_int, [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gpf_chain_diagnostics": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p]),
    "gpf_row_sort": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "gpf_function_interval": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gpf_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(ModelDesc)]),
    "gpf_destroy": (None, [C.c_void_p]),
    "gpf_last_error": (C.c_char_p, [C.c_void_p]),
    "gpf_stream": (C.c_void_p, [C.c_void_p]),
    "gpf_sync": (C.c_int, [C.c_void_p]),
    "gpf_set_state": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "gpf_get_state": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "gpf_memcpy_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "gpf_memcpy_d2h_event": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]),
    "gpf_record_event": (C.c_int, [C.c_void_p, C.c_int]),
    "gpf_wait_event": (C.c_int, [C.c_void_p, C.c_int]),
    "gpf_device_F": (C.c_void_p, [C.c_void_p]),
    "gpf_device_hyper": (C.c_void_p, [C.c_void_p]),
    "gpf_invalidate": (C.c_int, [C.c_void_p]),
    "gpf_n_samplers": (C.c_int, [C.c_void_p]),
    "gpf_get_status": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "gpf_get_counters": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "gpf_group_kinv": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "gpf_function_step": (C.c_int, [C.c_void_p, C.c_int, C.c_uint, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
    "gpf_obs_loglik": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "gpf_prior_loglik": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "gpf_slice_step": (C.c_int, [C.c_void_p, C.c_int, C.c_uint, C.c_uint, C.c_void_p, C.c_int, C.c_void_p]),
    "gpf_sweep": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_uint, C.c_void_p, C.c_void_p, C.c_longlong,
                            C.POINTER(C.c_longlong)]),
    "gpf_set_option": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "gpf_get_option": (C.c_int, [C.c_void_p, C.c_int]),
    "gpf_set_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "gpf_get_profile": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "gpf_debug_trace": (C.c_int, [C.c_void_p, C.c_void_p]),
}

_lib = None


def lib():
    """load (building first if needed) libgpfanova_b200.so"""
    global _lib
    if _lib is None:
        path = os.environ.get("GPFANOVA_B200_LIB") or _build.LIB      # (override: A/B builds of the same ABI)
        if path == _build.LIB and _build.stale():
            try:
                path = _build.build()
            except Exception:
                if not os.path.exists(path):
                    raise
        L = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class GpfError(RuntimeError):
    pass


def check(rc, ctx=None):
    if rc == OK:
        return
    msg = lib().gpf_last_error(ctx)
    msg = msg.decode() if msg else ""
    if rc == ERR_NOT_PD:
        raise np.linalg.LinAlgError(msg or "not positive definite, even with jitter.")
    if rc == ERR_ARG:
        raise ValueError(msg)
    raise GpfError("gpfanova_b200 error %d: %s" % (rc, msg))
