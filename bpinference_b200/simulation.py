"""Branching-process simulator — Python mirror of bp / bpsims (R/simulation.R:9-56, 69-131).

Data generator for parity datasets; not on the log-density path.  Same arguments, same returned data-frame
column order (times, rep, [variable_state, dependent variables...], populations...) which
stan_data_from_simulation indexes by position.  Two reference quirks are not reproduced: the extinction
branch that returns an undefined variable (R/simulation.R:28-34) fills the remaining time points with the
extinct state instead, and populations are reshaped with ntypes columns rather than ncol(c_mat)
(R/simulation.R:121).  The device version for large ensembles is libbpcuda's bpc_simulate.
"""
from __future__ import annotations

import math
import re

import numpy as np
import pandas as pd


def eval_rate(expr: str, c, x=None) -> float:
    """eval(model$func_deps[[j]], envir = list(c = theta, x = c_mat[i, ]))  (R/simulation.R:89,114) — R arithmetic,
    so `1/2` is 0.5 here (Stan's integer division applies only to the inference path)."""
    py = re.sub(r"\^", "**", expr)
    py = re.sub(r"\b([cx])\[(\d+)\]", lambda m: "%s[%d]" % (m.group(1), int(m.group(2)) - 1), py)
    env = {"exp": math.exp, "log": math.log, "c": list(c), "x": list(x) if x is not None else []}
    return float(eval(py, {"__builtins__": {}}, env))


def bp(e_mat, r_vec, p_vec, z0_vec, times, rng=None):
    rng = np.random.default_rng() if rng is None else rng
    e_mat = np.asarray(e_mat, dtype=float)
    times = np.asarray(times, dtype=float)
    p = np.asarray(p_vec, dtype=int) - 1
    r_vec = np.asarray(r_vec, dtype=float)
    z = np.asarray(z0_vec, dtype=float).copy()
    out = np.zeros((len(times), len(z)))
    out[0] = z
    t, tf = 0.0, times[-1]
    while t < tf:
        rates = r_vec * z[p]
        lam = rates.sum()
        if lam <= 0:
            out[times > t] = z
            return out
        dt = rng.exponential(1.0 / lam)
        ev = int(np.searchsorted(np.cumsum(rates), rng.uniform(0.0, lam), side="right"))
        ev = min(ev, len(rates) - 1)
        z = z + e_mat[ev]
        z[p[ev]] -= 1
        out[(t < times) & (t + dt >= times)] = z
        t += dt
    return out


def bp_cuda(e_mat, r_vec, p_vec, z0_vec, times, nrep, seed=1, device=0):
    """nrep replicates of bp() on the GPU (libbpcuda's bpc_simulate): returns [nrep, ntimes, ntypes].  Unlike bp(), `times`
    must be increasing; the first row of every replicate is the initial population, as in the reference."""
    from . import _ffi
    e_mat = np.asfortranarray(np.asarray(e_mat, dtype=np.float64))
    times = np.ascontiguousarray(np.atleast_1d(np.asarray(times, dtype=np.float64)))
    p = np.ascontiguousarray(np.asarray(p_vec, dtype=np.int32))
    r = np.ascontiguousarray(np.asarray(r_vec, dtype=np.float64))
    z0 = np.ascontiguousarray(np.atleast_1d(np.asarray(z0_vec, dtype=np.float64)))
    out = np.empty((int(nrep), len(times), e_mat.shape[1]))
    _ffi.check(_ffi.lib().bpc_simulate(e_mat.shape[1], e_mat.shape[0], _ffi.dptr(e_mat), _ffi.iptr(p), _ffi.dptr(r), _ffi.dptr(z0), _ffi.dptr(times),
                                       len(times), int(nrep), int(seed), int(device), _ffi.dptr(out)))
    return out


def bpsims(model, theta, z0_vec, times, reps, c_mat=None, rng=None, device=None, seed=1):
    """`device` = a CUDA device index runs the replicates with libbpcuda's Gillespie kernel instead of the host loop."""
    if getattr(model, "r_class", None) != "bp_model":
        raise ValueError("model must be a bp_model object!")
    rng = np.random.default_rng() if rng is None else rng
    e_mat = np.asarray(model["e_mat"], dtype=float)
    ntypes = e_mat.shape[1]
    no_c = c_mat is None
    if model["ndep"] > 0 and (no_c or np.asarray(c_mat).reshape(-1, np.asarray(c_mat).shape[-1] if np.ndim(c_mat) > 1 else 1).shape[1] != model["ndep"]):
        raise ValueError("Incorrect number of dependent variables were provided for the model!")
    theta = np.atleast_1d(np.asarray(theta, dtype=float))
    if len(theta) != model["nparams"]:
        raise ValueError("Wrong number of parameters were provided for the model!")
    z0_vec = np.atleast_1d(np.asarray(z0_vec, dtype=float))
    if len(z0_vec) != ntypes:
        raise ValueError("Initial popualation vector has wrong size!")
    times = np.asarray(times, dtype=float)
    pop_names = ["X%d" % (i + 1) for i in range(ntypes)]
    frames = []
    if model["ndep"] == 0 and no_c:
        r_vec = [eval_rate(f, theta) for f in model["func_deps"]]
        batch = bp_cuda(e_mat, r_vec, model["p_vec"], z0_vec, times, int(reps), seed=seed, device=device) if device is not None else None
        for i in range(int(reps)):
            pop = batch[i] if batch is not None else bp(e_mat, r_vec, model["p_vec"], z0_vec, times, rng)
            d = {"times": times, "rep": np.full(len(times), i + 1)}
            d.update({nm: pop[:, j] for j, nm in enumerate(pop_names)})
            frames.append(pd.DataFrame(d))
        return pd.concat(frames, ignore_index=True)
    cm = np.asarray(c_mat, dtype=float).reshape(-1, model["ndep"])
    reps = np.atleast_1d(np.asarray(reps, dtype=int))
    if len(reps) != cm.shape[0]:
        raise ValueError("reps should be a vector with a different number of replications for each dependent variable condition")
    dep_names = ["X%d" % (i + 1) for i in range(cm.shape[1])]
    pop_names = [nm if nm not in dep_names else nm + ".1" for nm in pop_names]     # data.frame(check.names) de-duplication
    curr = 1
    for i in range(cm.shape[0]):
        r_vec = [eval_rate(f, theta, cm[i]) for f in model["func_deps"]]
        batch = bp_cuda(e_mat, r_vec, model["p_vec"], z0_vec, times, int(reps[i]), seed=seed + 7919 * i, device=device) if device is not None else None
        for k in range(int(reps[i])):
            pop = batch[k] if batch is not None else bp(e_mat, r_vec, model["p_vec"], z0_vec, times, rng)
            d = {"times": times, "rep": np.full(len(times), curr), "variable_state": np.full(len(times), i + 1)}
            d.update({nm: np.full(len(times), cm[i, j]) for j, nm in enumerate(dep_names)})
            d.update({nm: pop[:, j] for j, nm in enumerate(pop_names)})
            frames.append(pd.DataFrame(d))
            curr += 1
    return pd.concat(frames, ignore_index=True)
