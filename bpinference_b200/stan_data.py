"""Stan data lists — Python mirror of create_stan_data / stan_data_from_simulation / uniform_initialize
(R/stan_utils.R:12-66, 96-120, 74-87).  The returned dict has the reference's keys, shapes, 1-based index
vectors and the dummy `function_var = matrix(0,1,1)` when the model has no dependent variables.

One deliberate difference: the reference stores `times = array(times_unique, 1)` (R/stan_utils.R:58), i.e. only
the first distinct duration, which is inconsistent with `ntimes_unique` as soon as there are two; here all
distinct durations are kept (identical whenever the reference's list is usable).
"""
from __future__ import annotations

import warnings

import numpy as np


def _as_matrix(x, name):
    a = np.asarray(x, dtype=float)
    if a.ndim == 1:
        warnings.warn("%s is a vector, not a matrix. Converting to a one-column matrix" % name)
        a = a.reshape(-1, 1)
    return a


def _unique_first(a):
    """unique() in order of first appearance, and match() indices (1-based)."""
    seen, order, idx = {}, [], np.empty(len(a), dtype=np.int32)
    for i, v in enumerate(a):
        key = v if not isinstance(v, np.ndarray) else tuple(v.tolist())
        if key not in seen:
            seen[key] = len(order) + 1
            order.append(v)
        idx[i] = seen[key]
    return order, idx


def create_stan_data(model, final_pop, init_pop, times, c_mat=None, simple_bd=False) -> dict:
    final_pop = _as_matrix(final_pop, "final_pop")
    init_pop = _as_matrix(init_pop, "init_pop")
    times = np.asarray(times, dtype=float).ravel()
    if final_pop.shape[0] != init_pop.shape[0] or init_pop.shape[0] != len(times):
        raise ValueError("times, initial populations, and final populations must all have same number of rows!")
    e_mat = np.asarray(model["e_mat"], dtype=float)
    nevents, ntypes = e_mat.shape
    ndatapts = final_pop.shape[0]
    tu, times_idx = _unique_first(times.tolist())
    times_unique = np.asarray(tu, dtype=float)
    no_c = c_mat is None or (np.ndim(c_mat) == 0 and np.isnan(c_mat))
    if model["ndep"] > 0 and (no_c or np.asarray(c_mat).reshape(ndatapts, -1).shape[1] < model["ndep"]):
        raise ValueError("c_mat does not contain enough dependent variables for the model")
    if model["ndep"] == 0 and no_c:
        c_mat_unique = np.zeros((1, 1))
        ndep_levels, ndep = 1, 1
        var_idx = np.ones(ndatapts, dtype=np.int32)
    else:
        cm = np.asarray(c_mat, dtype=float).reshape(ndatapts, -1)
        rows, var_idx = _unique_first([r for r in cm])
        c_mat_unique = np.asarray(rows, dtype=float).reshape(len(rows), cm.shape[1])
        ndep_levels, ndep = c_mat_unique.shape
    if not simple_bd:
        return dict(ntypes=ntypes, nevents=nevents, ndatapts=ndatapts, ntimes_unique=len(times_unique), ndep_levels=ndep_levels,
                    ndep=ndep, nparams=model["nparams"], e_mat=e_mat, p_vec=np.asarray(model["p_vec"], dtype=np.int32),
                    pop_vec=final_pop, init_pop=init_pop, times=times_unique, times_idx=times_idx, function_var=c_mat_unique,
                    var_idx=var_idx)
    return dict(pop_vec=final_pop.ravel(order="F"), init_pop=init_pop.ravel(order="F"), times=times, function_var=c_mat_unique,
                var_idx=var_idx, ndep_levels=ndep_levels, ndep=ndep, nparams=model["nparams"], ndatapts=ndatapts)


def uniform_initialize(ranges, nchains, rng=None):
    """list (one dict per chain) of Theta%d = runif(lower, upper) — R/stan_utils.R:74-87 (without its debug print)."""
    rng = np.random.default_rng() if rng is None else rng
    ranges = np.asarray(ranges, dtype=float).reshape(-1, 2)
    return [{"Theta%d" % (i + 1): float(rng.uniform(lo, hi)) for i, (lo, hi) in enumerate(ranges)} for _ in range(nchains)]


def stan_data_from_simulation(sim_data, model, simple_bd=False) -> dict:
    """sim_data: the data frame bpsims returns (columns by POSITION: times, rep, [variable_state, dep...], pops...)."""
    ntypes = np.asarray(model["e_mat"]).shape[1]
    ndep = int(model["ndep"])
    offset = ndep + (1 if ndep > 0 else 0)
    df = sim_data.reset_index(drop=True)
    cols = list(df.columns)
    times = df[cols[0]].to_numpy(dtype=float)
    rep = df[cols[1]].to_numpy()
    pops = df[cols[2 + offset: 2 + offset + ntypes]].to_numpy(dtype=float)
    same = np.zeros(len(df), dtype=bool)
    same[1:] = rep[1:] == rep[:-1]          # dplyr::lag within group_by(rep): first row of each replicate is NA
    keep = np.nonzero(same)[0]
    init_pop, final_pop = pops[keep - 1], pops[keep]
    dtimes = times[keep] - times[keep - 1]
    if ndep > 0:
        c_mat = df[cols[3: 3 + ndep]].to_numpy(dtype=float)[keep]
        return create_stan_data(model, final_pop, init_pop, dtimes, c_mat, simple_bd=simple_bd)
    return create_stan_data(model, final_pop, init_pop, dtimes, simple_bd=simple_bd)
