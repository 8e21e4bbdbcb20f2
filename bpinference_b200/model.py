"""bp_model objects — Python mirror of R/bp_model.R (same constructor names, argument meaning and error text).

Reference: new_bp_model R/bp_model.R:12-17, validate_bp_model :24-68, bp_model :80-83,
bp_model_simple_birth_death :93-97.  Errors are raised as ValueError carrying the reference's stop()
message; R warnings become Python warnings.  Rate expressions are parsed and validated by libbpcuda's
front end (the C++ twin of check_valid, R/stan_utils.R:289-364).
"""
from __future__ import annotations

import numbers
import warnings

import numpy as np

from . import _ffi


class BpModel(dict):
    """list(e_mat, p_vec, func_deps, nparams, ndep) with class "bp_model" (R/bp_model.R:15-16).
    `func_deps` holds the canonical deparse() spelling of each parsed rate expression."""

    r_class = "bp_model"

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


def _is_numeric(x) -> bool:
    if isinstance(x, (bool, np.bool_)):
        return False
    if isinstance(x, numbers.Number):
        return True
    if isinstance(x, np.ndarray):
        return np.issubdtype(x.dtype, np.number)
    if isinstance(x, (list, tuple)):
        return len(x) > 0 and all(isinstance(v, numbers.Number) and not isinstance(v, bool) for v in np.asarray(x, dtype=object).ravel())
    return False


def deparse(expr: str) -> str:
    """R's deparse(parse(text = expr)[[1]])."""
    buf = _ffi.C.create_string_buffer(4096 + 4 * len(expr))
    _ffi.check(_ffi.lib().bpc_expr_deparse(expr.encode(), buf, len(buf)))
    return buf.value.decode()


def check_valid(exprn: str, max_params: int, max_dep: int) -> None:
    """R/stan_utils.R:289-364 — silent when valid, ValueError with the reference's message otherwise."""
    rc = _ffi.lib().bpc_check_expr(str(exprn).encode(), int(max_params), int(max_dep))
    if rc:
        raise ValueError(_ffi.lib().bpc_last_error().decode())


def exp_to_stan(exprn: str, max_params: int, max_dep: int) -> str:
    """R/stan_utils.R:205-281 — the Stan spelling of a rate expression (kept for drop-in users and tests)."""
    buf = _ffi.C.create_string_buffer(8192 + 16 * len(str(exprn)))
    rc = _ffi.lib().bpc_expr_to_stan(str(exprn).encode(), int(max_params), int(max_dep), buf, len(buf))
    if rc:
        raise ValueError(_ffi.lib().bpc_last_error().decode())
    return buf.value.decode()


def new_bp_model(e_mat, p_vec, func_deps, nparam, ndep) -> BpModel:
    return BpModel(e_mat=e_mat, p_vec=p_vec, func_deps=func_deps, nparams=nparam, ndep=ndep)


def validate_bp_model(bpm: BpModel) -> BpModel:
    if not (_is_numeric(bpm["p_vec"]) and _is_numeric(bpm["e_mat"]) and _is_numeric(bpm["nparams"]) and _is_numeric(bpm["ndep"])):
        raise ValueError("e_mat, p_mat, ndep, and nparams must all be numeric!")
    p_vec = np.asarray(bpm["p_vec"], dtype=float)
    if p_vec.ndim != 1 or not np.all(np.round(p_vec) == p_vec) or not np.all(p_vec > 0):
        raise ValueError("p_vec must be a vector of parents for each bith event, where each entry is an positive integer "
                         "corresponding to the parent type")
    e_mat = np.asarray(bpm["e_mat"], dtype=float)
    if e_mat.ndim == 1:
        warnings.warn("e_mat is a vector, not a matrix. Converting to a one-column matrix")
        e_mat = e_mat.reshape(-1, 1)
    if e_mat.ndim != 2 or not np.all(np.round(e_mat) == e_mat) or not np.all(e_mat >= 0):
        raise ValueError("e_mat must be a matrix of birth events, where each entry is an integer number of offspring of a given type")
    fd = bpm["func_deps"]
    if isinstance(fd, str):
        fd = [fd]
    if not isinstance(fd, (list, tuple, np.ndarray)) or not all(isinstance(s, str) for s in fd):
        raise ValueError("func_deps must be a vector of strings relating brith rates to dependent variables")
    parsed = [deparse(s) for s in fd]          # parse(text = .)[[1]]; syntax errors surface here
    if e_mat.shape[0] != len(p_vec) or e_mat.shape[0] != len(parsed):
        raise ValueError("Dimensions of e_mat, p_vec, and func_deps do not agree")
    nparams, ndep = bpm["nparams"], bpm["ndep"]
    if round(nparams) != nparams or nparams <= 0:
        raise ValueError("nparams should be an integer number of parameters")
    if round(ndep) != ndep or ndep < 0:
        raise ValueError("ndep should be an integer number of depedent variables")
    # the reference tests `p_vec <= bpm$ntypes`, a field that does not exist (R/bp_model.R:29), so an
    # out-of-range parent is never caught there; here it must be, the kernels index by it
    if np.any(p_vec > e_mat.shape[1]):
        raise ValueError("p_vec must be a vector of parents for each bith event, where each entry is an positive integer "
                         "corresponding to the parent type")
    for s in parsed:
        check_valid(s, int(nparams), int(ndep))
    out = BpModel(e_mat=e_mat, p_vec=p_vec.astype(np.int32), func_deps=parsed, nparams=int(nparams), ndep=int(ndep))
    return out


def bp_model(e_mat, p_vec, func_deps, nparam, ndep) -> BpModel:
    return validate_bp_model(new_bp_model(e_mat, p_vec, func_deps, nparam, ndep))


def bp_model_simple_birth_death(func_deps, nparam, ndep) -> BpModel:
    return validate_bp_model(new_bp_model(np.array([[2.0], [0.0]]), [1, 1], func_deps, nparam, ndep))
