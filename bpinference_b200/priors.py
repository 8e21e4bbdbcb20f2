"""prior_dist objects — Python mirror of R/stan_utils.R:372-487 (same names, validation order and messages)."""
from __future__ import annotations

import numbers

import numpy as np

from .model import deparse

# ALLOWED_DISTRIBUTIONS of R/sysdata.rda: "name nparams constraints" ('+' = parameter must be > 0)
ALLOWED_DISTRIBUTIONS = [
    "normal 2 _+", "exp_mod_normal 3 _++", "skew_normal 3 _+_", "student_t 3 +_+", "cauchy 2 _+", "double_exponential 2 _+",
    "logistic 2 _+", "gumbel 2 _+", "lognormal 2 _+", "chi_square 1 +", "inv_chi_square 1 +", "scaled_inv_chi_square 2 ++",
    "exponential 1 +", "gamma 2 ++", "inv_gamma 2 ++", "weibull 2 ++", "frechet 2 ++", "rayleigh 1 +", "pareto 2 ++",
    "pareto_type_2 3 _++", "beta 2 ++", "uniform 2 __"]
_TABLE = {s.split()[0]: (int(s.split()[1]), s.split()[2]) for s in ALLOWED_DISTRIBUTIONS}


class PriorDist(dict):
    r_class = "prior_dist"

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


def r_num(x) -> str:
    """paste(x): R's 15-significant-digit spelling of a number."""
    return deparse(repr(float(x)))


def _is_na(b) -> bool:
    if b is None:
        return True
    if isinstance(b, str):
        return False
    try:
        a = np.asarray(b, dtype=float)
    except (TypeError, ValueError):
        return False
    return bool(np.any(np.isnan(a)))


def _numeric_vec(x):
    if isinstance(x, (str, bytes, bool)) or x is None:
        return None
    if isinstance(x, numbers.Number):
        return [float(x)]
    try:
        a = np.asarray(x)
    except Exception:
        return None
    if a.dtype.kind not in "iuf":
        return None
    return [float(v) for v in a.ravel()]


def _check_core(name, params):
    if name not in _TABLE:
        raise ValueError("Invalid prior name!")
    nparm, cons = _TABLE[name]
    if nparm != len(params):
        raise ValueError("Incorrect number of parameters for prior distribution")
    for i in range(nparm):
        if params[i] <= 0 and cons[i] == "+":
            raise ValueError("Positive parameter required here!")


def _check_bounds(bounds):
    if not _is_na(bounds):
        b = _numeric_vec(bounds)
        if b is None or len(b) != 2:
            raise ValueError("bounds should be a list of 2 numbers!")
        if b[0] >= b[1]:
            raise ValueError("Error: lower bounder greater or equal to upper bound!")
        return b
    return None


def new_prior_dist(name, params, bounds=None) -> PriorDist:
    return PriorDist(name=name, params=params, bounds=bounds)


def validate_prior_dist(prior) -> PriorDist:
    if getattr(prior, "r_class", None) != "prior_dist":
        raise ValueError("validate_prior_dist required an object of type prior_dist")
    params = _numeric_vec(prior["params"])
    _check_core(prior["name"], params)
    if prior["name"] == "uniform" and params[1] <= params[0]:
        raise ValueError("upper bound must be greater than lower in uniform distribution!")
    _check_bounds(prior["bounds"])
    return prior


def prior_dist(name, params, bounds=None) -> PriorDist:
    if not isinstance(name, str):
        raise ValueError("prior name must be a string!")
    if _numeric_vec(params) is None:
        raise ValueError("parameters should be numeric!")
    return validate_prior_dist(new_prior_dist(name, params, bounds))


def parse_prior(prior, k: int):
    """[prior statement, parameter declaration] exactly as R/stan_utils.R:372-412 emits them.  Accepts plain
    dicts as the reference accepts plain lists (tests/testthat/test_generation.R:60)."""
    params = _numeric_vec(prior["params"])
    _check_core(prior["name"], params)
    prior_str = "\tTheta%d ~ %s(%s);\n" % (k, prior["name"], ",".join(r_num(p) for p in params))
    b = _check_bounds(prior.get("bounds"))
    if b is not None:
        decl = "\treal<lower=%s, upper=%s> Theta%d;\n" % (r_num(b[0]), r_num(b[1]), k)
    else:
        decl = "\treal Theta%d;\n" % k
    return [prior_str, decl]


def prior_fields(prior):
    """(name, params list, bounds or None) of a prior_dist / plain dict, validated the way generate() does."""
    params = _numeric_vec(prior["params"])
    _check_core(prior["name"], params)
    return prior["name"], params, _check_bounds(prior.get("bounds"))
