"""Compares libbpcuda with rstan: reads the JSON written by r-package/parity_dump.R (rstan::log_prob / grad_log_prob at
fixture points for the two-type example built with the reference's own bpsims + generate) and evaluates the same points
on the GPU.  Needs a machine that has both the dump and a B200.   python tools/compare_rstan_dump.py out.json"""
import json
import sys

import numpy as np

import bpinference_b200 as bp

j = json.load(open(sys.argv[1]))
dat = j["dat"]
mod = bp.bp_model(np.asarray(dat["e_mat"], dtype=float), dat["p_vec"], ["c[%d]" % k for k in range(1, 7)], 6, 0)
eng = bp.stan_model(mod, [dict(name="normal", params=[0, .25], bounds=[0, 5])] * 6)
upars = np.asarray(j["upars"], dtype=float).T                 # R writes 6 x 16 column-major
lp, grad, st = eng.log_prob_grad(dat, upars)
lp_r, grad_r = np.asarray(j["lp"], dtype=float), np.asarray(j["grad"], dtype=float).T
print("status", st.tolist())
# Stan drops additive constants under `~`: compare log-density DIFFERENCES between points, and gradients directly
print("max |d(lp - lp[0])| relative:", np.max(np.abs((lp - lp[0]) - (lp_r - lp_r[0])) / np.maximum(1.0, np.abs(lp_r - lp_r[0]))))
print("max gradient relative difference:", np.max(np.abs(grad - grad_r) / np.maximum(1e-300, np.abs(grad_r).max(axis=1, keepdims=True))))
