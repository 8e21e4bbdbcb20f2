"""Compares libbpcuda with rstan: reads the JSON written by r-package/parity_dump.R (rstan::log_prob / grad_log_prob at fixture
points for the five configurations of BASELINE.json, built with the reference's own bpsims + generate) and evaluates the same
points on the GPU.  Needs a machine that has both the dump and a B200.   python tools/compare_rstan_dump.py out.json

rstan integrates the moment ODE with RK45 at rel/abs tolerance 1e-6 (integrate_ode_rk45 defaults), libbpcuda evaluates the same
moments to double-precision round-off; tests/test_oracle.py measures that gap at 1e-10 .. 1e-8 relative in lp — expect agreement at
that level, not at 1e-15."""
import json
import sys

import numpy as np

import bpinference_b200 as bp


def _mat(x, nrow=None):
    a = np.asarray(x, dtype=float)
    return a if nrow is None or a.ndim == 2 else a.reshape(nrow, -1)


def compare(entry):
    m, dat = entry["model"], dict(entry["dat"])
    e_mat = _mat(m["e_mat"])
    func_deps = [m["func_deps"]] if isinstance(m["func_deps"], str) else list(m["func_deps"])
    mod = bp.bp_model(e_mat, np.atleast_1d(m["p_vec"]).tolist(), func_deps, int(m["nparams"]), int(m["ndep"]))
    priors = []
    for pr in entry["priors"]:
        b = pr.get("bounds")
        has = isinstance(b, list) and len(b) == 2 and all(v is not None for v in b)
        priors.append(dict(name=pr["name"], params=np.atleast_1d(pr["params"]).tolist(), bounds=b if has else None))
    eng = bp.stan_model(mod, priors, simple_bd=bool(entry["simple_bd"]), noise_model=bool(entry["noise_model"]))
    # jsonlite's auto_unbox turns every length-one vector into a scalar (`times` of a one-duration design, ...): back to arrays
    for k in ("pop_vec", "init_pop", "function_var", "e_mat", "times"):
        if k in dat:
            dat[k] = np.atleast_1d(np.asarray(dat[k], dtype=float))
    for k in ("p_vec", "times_idx", "var_idx"):
        if k in dat:
            dat[k] = np.atleast_1d(np.asarray(dat[k], dtype=int))
    if "function_var" in dat and dat["function_var"].ndim == 1:
        dat["function_var"] = dat["function_var"].reshape(-1, max(1, int(dat.get("ndep", 1)) or 1))
    upars = np.asarray(entry["upars"], dtype=float).T                 # R writes dim x points column-major
    lp, grad, st = eng.log_prob_grad(dat, upars)
    lp_nj, _, _ = eng.log_prob_grad(dat, upars, adjust_transform=False)
    lp_r, lp_rn, grad_r = np.asarray(entry["lp"], dtype=float), np.asarray(entry["lp_nojac"], dtype=float), np.asarray(entry["grad"], dtype=float).T
    out = {"name": entry["name"], "status": st.tolist()}
    # Stan drops additive constants under `~`: compare log-density DIFFERENCES between points, and gradients directly
    out["lp_diff_rel"] = float(np.max(np.abs((lp - lp[0]) - (lp_r - lp_r[0])) / np.maximum(1.0, np.abs(lp_r - lp_r[0]))))
    out["lp_nojac_diff_rel"] = float(np.max(np.abs((lp_nj - lp_nj[0]) - (lp_rn - lp_rn[0])) / np.maximum(1.0, np.abs(lp_rn - lp_rn[0]))))
    out["grad_rel"] = float(np.max(np.abs(grad - grad_r) / np.maximum(1e-300, np.abs(grad_r).max(axis=1, keepdims=True))))
    if "log_lik" in entry:
        ll = eng.log_lik(dat, upars[:1])[0]
        out["log_lik_rel"] = float(np.max(np.abs(ll - np.asarray(entry["log_lik"], dtype=float)) / np.abs(ll)))
    return out


if __name__ == "__main__":
    j = json.load(open(sys.argv[1]))
    entries = j.values() if isinstance(j, dict) and "dat" not in j else [j]
    for e in entries:
        print(json.dumps(compare(e)))
