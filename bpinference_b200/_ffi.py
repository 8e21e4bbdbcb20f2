"""ctypes binding of libbpcuda's C ABI (include/bpcuda.h) — the Python twin of the R `.Call` shim.

Nothing here computes: it translates Python objects to the plain structs of bpcuda.h and raises
`BpcError` carrying bpc_last_error() when a call fails.  The shared library must exist in-tree
(bpinference_b200/libbpcuda.so, built by `__graft_entry__.build()` / csrc/Makefile): there is no
fallback path.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbpcuda.so")

MULTITYPE, MULTITYPE_NOISE, SIMPLE_BD = 0, 1, 2
ST_NOT_PD, ST_RATE_NONFINITE, ST_ONETYPE_SINGULAR, ST_PRIOR_SUPPORT = 1, 2, 4, 8
E_INVALID, E_EXPR, E_PRIOR, E_COMPILE, E_CUDA, E_UNSUPPORTED = 1, 2, 3, 4, 5, 6


class BpcError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(message)
        self.code = code


class Prior(C.Structure):
    _fields_ = [("name", C.c_char_p), ("nparams", C.c_int), ("params", C.c_double * 3), ("has_bounds", C.c_int),
                ("lower", C.c_double), ("upper", C.c_double)]


class ModelDesc(C.Structure):
    _fields_ = [("ntypes", C.c_int), ("nevents", C.c_int), ("nparams", C.c_int), ("ndep", C.c_int),
                ("e_mat", C.POINTER(C.c_double)), ("p_vec", C.POINTER(C.c_int)), ("func_deps", C.POINTER(C.c_char_p)),
                ("priors", C.POINTER(Prior)), ("kind", C.c_int)]


class StanData(C.Structure):
    _fields_ = [("ndatapts", C.c_int), ("ntimes_unique", C.c_int), ("ndep_levels", C.c_int), ("ndep", C.c_int),
                ("pop_vec", C.POINTER(C.c_double)), ("init_pop", C.POINTER(C.c_double)), ("times", C.POINTER(C.c_double)),
                ("times_idx", C.POINTER(C.c_int)), ("function_var", C.POINTER(C.c_double)), ("var_idx", C.POINTER(C.c_int))]


class SampleArgs(C.Structure):
    _fields_ = [("chains", C.c_int), ("iter", C.c_int), ("warmup", C.c_int), ("adapt_delta", C.c_double),
                ("max_treedepth", C.c_int), ("seed", C.c_ulonglong), ("chain_id_offset", C.c_int),
                ("inits", C.POINTER(C.c_double)), ("pooled_adaptation", C.c_int), ("stepsize", C.c_double), ("init_opt_steps", C.c_int), ("metric", C.c_int), ("init_opt_data", C.c_void_p)]


# every symbol include/bpcuda.h declares: name -> (restype, argtypes)
_dp, _ip, _vp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p
SYMBOLS = {
    "bpc_abi_version": (C.c_int, []),
    "bpc_last_error": (C.c_char_p, []),
    "bpc_device_count": (C.c_int, []),
    "bpc_fp64_peak": (C.c_int, [C.c_int, C.c_double, _dp, _dp]),
    "bpc_fp64_tensor_peak": (C.c_int, [C.c_int, C.c_double, _dp]),
    "bpc_check_expr": (C.c_int, [C.c_char_p, C.c_int, C.c_int]),
    "bpc_expr_to_stan": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_int]),
    "bpc_expr_deparse": (C.c_int, [C.c_char_p, C.c_char_p, C.c_int]),
    "bpc_check_prior": (C.c_int, [C.POINTER(Prior)]),
    "bpc_model_create": (C.c_int, [C.POINTER(ModelDesc), C.POINTER(_vp)]),
    "bpc_model_free": (None, [_vp]),
    "bpc_model_dim": (C.c_int, [_vp]),
    "bpc_model_cuda_source": (C.c_char_p, [_vp]),
    "bpc_model_cubin": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(C.c_ulonglong)]),
    "bpc_model_warnings": (C.c_char_p, [_vp]),
    "bpc_data_create": (C.c_int, [_vp, C.POINTER(StanData), C.c_int, C.c_int, C.POINTER(_vp)]),
    "bpc_data_free": (None, [_vp]),
    "bpc_data_ngroups": (C.c_int, [_vp]),
    "bpc_data_is_grouped": (C.c_int, [_vp]),
    "bpc_data_path": (C.c_int, [_vp, C.c_int, _ip, _ip, C.POINTER(C.c_ulonglong)]),
    "bpc_log_prob": (C.c_int, [_vp, _vp, _dp, C.c_int, C.c_int, _dp, _dp, _ip]),
    "bpc_log_prob_dev": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "bpc_log_lik": (C.c_int, [_vp, _vp, _dp, C.c_int, _dp]),
    "bpc_moments": (C.c_int, [_vp, _vp, _dp, C.c_int, _dp, _dp]),
    "bpc_transformed_parameters": (C.c_int, [_vp, _vp, _dp, C.c_int, _dp]),
    "bpc_data_nlevels": (C.c_int, [_vp]),
    "bpc_flop_model": (C.c_int, [_vp, _dp, _dp, _dp]),
    "bpc_flop_model_centred": (C.c_int, [_vp, _dp, _dp, _dp]),
    "bpc_constrain": (C.c_int, [_vp, _dp, C.c_int, _dp]),
    "bpc_unconstrain": (C.c_int, [_vp, _dp, C.c_int, _dp]),
    "bpc_launch_count": (C.c_longlong, [C.c_int]),
    "bpc_kernel_timing": (C.c_int, [_vp, C.c_int]),
    "bpc_kernel_time_ms": (C.c_int, [_vp, C.c_int, _dp, C.POINTER(C.c_longlong)]),
    "bpc_count_flops": (C.c_int, [_vp, _vp, _dp, C.c_int, _dp]),
    "bpc_sampler_create": (C.c_int, [_vp, _vp, C.POINTER(SampleArgs), C.POINTER(_vp)]),
    "bpc_sampler_free": (None, [_vp]),
    "bpc_sampler_run": (C.c_int, [_vp, C.c_longlong, _ip]),
    "bpc_sampler_pool_stats": (C.c_int, [_vp, _vp, C.c_int]),
    "bpc_sampler_pool_len": (C.c_int, [_vp]),
    "bpc_sampler_pool_apply": (C.c_int, [_vp, _vp]),
    "bpc_sampler_chain_moments": (C.c_int, [_vp, _vp, C.c_int, _dp]),
    "bpc_rhat": (C.c_int, [_dp, C.c_int, C.c_int, _dp]),
    "bpc_sampler_passes": (C.c_longlong, [_vp]),
    "bpc_sampler_progress": (C.c_int, [_vp, _ip, _dp, _dp, _ip]),
    "bpc_sampler_draws": (C.c_int, [_vp, _dp, _dp, _ip, _ip, _ip, _dp, _dp]),
    "bpc_sampler_total_leapfrogs": (C.c_longlong, [_vp]),
    "bpc_sample": (C.c_int, [_vp, _vp, C.POINTER(SampleArgs), _dp, _dp, _ip, _ip, _ip, _dp, _dp]),
    "bpc_sample_multi": (C.c_int, [_vp, C.POINTER(StanData), _ip, C.c_int, C.POINTER(SampleArgs), _dp, _dp, _ip, _ip, _ip, _dp, _dp, _dp,
                                   C.POINTER(C.c_longlong)]),
    "bpc_nccl_version": (C.c_int, []),
    "bpc_simulate": (C.c_int, [C.c_int, C.c_int, _dp, _ip, _dp, _dp, _dp, C.c_int, C.c_int, C.c_ulonglong, C.c_int, _dp]),
}

_lib = None


def lib():
    """loads libbpcuda.so (fails loudly when it has not been built: no fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(make -C bpinference_b200/csrc).  bpinference_b200 has no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise BpcError(rc, lib().bpc_last_error().decode("utf-8", "replace"))


def dptr(a):
    return a.ctypes.data_as(_dp)


def iptr(a):
    return a.ctypes.data_as(_ip)
