"""Generates tests/golden/lp_grad_golden.json: small seeded instances of the five configs with the log density, gradient
and status computed by the CPU oracle in long double (precision=1).  Run from the repo root:  python tests/golden/make_golden.py

The reference itself cannot produce these numbers here (R + rstan are not installed), so the golden
file pins (a) the reference's printed vignette values and (b) the oracle's own output at fixed inputs — a regression anchor
that both the oracle (-m "not gpu") and the CUDA path (-m gpu) are tested against."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bpinference_b200 import synthetic as syn   # noqa: E402
from oracle import bp_oracle as bo              # noqa: E402

CASES = [("cfg1", None, 3), ("cfg2", 30, 3), ("cfg3", 60, 3), ("cfg4", 25, 3), ("cfg5U", 24, 3), ("cfg5G", 36, 3), ("cfg5K", 24, 3)]


def tolist(x):
    return np.asarray(x).tolist()


def main():
    out = {"vignettes": {
        "two_type": {"e_mat": [[2, 0], [1, 1], [1, 1], [0, 2], [0, 0], [0, 0]], "p_vec": [1, 1, 2, 2, 1, 2], "rates": [0.20, 0.10, 0.05, 0.20, 0.25, 0.20],
                     "z0": [100, 50], "t": 1, "mean_printed": [97.8054257205992, 59.8852387358044],
                     "cov_printed_desolve_1e-6": [45.0418961482491, 3.02422137683959, 3.02422137683959, 32.0727485045799],
                     "source": "vignettes/two-type-inference.pdf, chunk two-type-inference.Rmd:38-44"},
        "one_type": {"rates": [0.25, 0.10], "z0": [1000], "t": 1, "mean_printed": [1161.83424272828], "var_printed_desolve_1e-6": [438.725421279098],
                     "source": "vignettes/first-model.pdf, chunk first-model.Rmd:63-67"}}, "cases": []}
    for name, nobs, chains in CASES:
        cfg = syn.make_config(name, seed=7, chains=chains, nobs=nobs)
        om = bo.OracleModel(cfg["kind"], cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"], cfg["priors"])
        u = np.round(cfg["upars"], 6)
        res = {}
        for jac in (True, False):
            lp, g, st = bo.log_prob(om, cfg["data"], u, jacobian=jac, precision=1)
            res["jacobian_%d" % jac] = {"lp": tolist(lp), "grad": tolist(g), "status": tolist(st)}
        d = {k: (tolist(v) if isinstance(v, np.ndarray) else v) for k, v in cfg["data"].items()}
        out["cases"].append({"name": name, "kind": cfg["kind"], "e_mat": tolist(cfg["e_mat"]), "p_vec": tolist(cfg["p_vec"]), "func_deps": cfg["func_deps"],
                             "nparams": cfg["nparams"], "ndep": cfg["ndep"], "priors": cfg["priors"], "data": d, "upars": tolist(u), "expected": res})
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "lp_grad_golden.json"), "w") as f:
        json.dump(out, f)
    print("wrote", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
