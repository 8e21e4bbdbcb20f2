"""calculate_moments — Python mirror of R/calculate_moments.R:11-88 (same arguments, validation order and messages,
same returned names mu_mat / sigma_mat), evaluated by libbpcuda's bp_moments kernel.

Two reference quirks are not reproduced: the R code integrates the second moments over seq(0, tf) and so stops at
floor(tf) for a non-integer tf, and it divides by the total rate of each type (NaN for a type without events);
here the moments are evaluated at tf itself and a type without events simply stays constant."""
from __future__ import annotations

import numbers
import warnings

import numpy as np


def _numeric(x):
    if isinstance(x, (str, bytes)) or x is None:
        return False
    try:
        return np.asarray(x).dtype.kind in "iuf"
    except Exception:
        return False


def calculate_moments(e_mat, p_vec, r_vec, z0_vec, tf):
    if not (_numeric(e_mat) and _numeric(r_vec) and _numeric(z0_vec) and _numeric(tf)):
        raise ValueError("all parameters should be numeric!")
    e_mat = np.asarray(e_mat, dtype=float)
    if e_mat.ndim == 1:
        warnings.warn("e_mat is a vector, not a matrix. Converting to a one-column matrix.")
        e_mat = e_mat.reshape(-1, 1)
    p_vec, r_vec, z0_vec = np.asarray(p_vec, dtype=float), np.asarray(r_vec, dtype=float), np.asarray(z0_vec, dtype=float)
    if p_vec.ndim != 1 or r_vec.ndim != 1 or z0_vec.ndim != 1:
        raise ValueError("p_vec, r_vec, and z0_vec should all be vectors!")
    if tf < 0:
        raise ValueError("tf cannot be negative!")
    ntype = len(z0_vec)
    if e_mat.shape[0] != len(r_vec):
        raise ValueError("Incorrect number of rate parameters provided for model!")
    if e_mat.shape[1] != ntype:
        raise ValueError("Dimensions of birth events matrix disagree with dimension of initial population vector!")
    if e_mat.shape[0] != len(p_vec):
        raise ValueError("Dimensions of birth events matrix disagree with dimension of parent vector!")
    if np.any((p_vec > ntype) | (p_vec < 1) | (np.round(p_vec) != p_vec)):
        raise ValueError("Parent vector must have integer entries between 1 and the number of types!")
    from .engine import stan_model
    from .model import bp_model
    from .stan_data import create_stan_data
    E = e_mat.shape[0]
    mod = bp_model(e_mat, p_vec, ["c[%d]" % (e + 1) for e in range(E)], E, 0)
    eng = stan_model(mod, [dict(name="normal", params=[0.0, 1.0], bounds=None)] * E)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        dat = create_stan_data(mod, z0_vec.reshape(1, -1), z0_vec.reshape(1, -1), [float(tf)])
    mu, sig = eng.moments(dat, r_vec.reshape(1, -1))
    return dict(mu_mat=mu[0, 0].reshape(-1, 1), sigma_mat=sig[0, 0])


def template_moments(e_mat, p_vec, r_vec, tf):
    """the Stan template's own state at duration tf (multitype_template.stan:85-94, 137-140): m_t [n, n] with
    m_t[l, j] = E[Z_j | one type-l ancestor] and d_t [n, n*n] with d_t[l, j + n*k] = E[Z_j Z_k | one type-l ancestor]
    (raw second moments, column-major to_matrix layout).  Evaluated on the device: one unit-ancestor observation per type."""
    e_mat = np.asarray(e_mat, dtype=float)
    if e_mat.ndim == 1:
        e_mat = e_mat.reshape(-1, 1)
    E, n = e_mat.shape
    from .engine import stan_model
    from .model import bp_model
    from .stan_data import create_stan_data
    mod = bp_model(e_mat, np.asarray(p_vec, dtype=float), ["c[%d]" % (e + 1) for e in range(E)], E, 0)
    eng = stan_model(mod, [dict(name="normal", params=[0.0, 1.0], bounds=None)] * E)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        dat = create_stan_data(mod, np.eye(n), np.eye(n), [float(tf)] * n)
    mu, sig = eng.moments(dat, np.asarray(r_vec, dtype=float).reshape(1, -1))
    m_t = mu[0]                                                     # row l: mean vector from one type-l ancestor
    d_t = np.empty((n, n * n))
    for l in range(n):
        d_t[l] = (sig[0, l] + np.outer(m_t[l], m_t[l])).reshape(-1, order="F")
    return m_t, d_t
