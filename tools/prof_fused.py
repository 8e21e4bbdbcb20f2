"""profiling driver: a few launches of the fused kernel on the bench workload (for ncu captures)."""
import sys
import numpy as np
import bpinference_b200 as bp
from bpinference_b200 import synthetic as syn
name = sys.argv[1] if len(sys.argv) > 1 else "cfg5U"
chains = int(sys.argv[2]) if len(sys.argv) > 2 else 512
cfg = syn.make_config(name, chains=chains, nobs=100000 if name.startswith("cfg5") else None)
mod = bp.bp_model(cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"])
eng = bp.stan_model(mod, cfg["priors"], simple_bd=cfg["kind"] == "simple_bd", noise_model=cfg["kind"] == "multitype_noise")
d = eng.data(cfg["data"])
for i in range(4):
    lp, g, st = eng.log_prob_grad(d, cfg["upars"])
print("ok", float(lp[0]), int((st != 0).sum()))
