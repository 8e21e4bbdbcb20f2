"""small octet-path call for compute-sanitizer (memcheck / racecheck): python tools/sanitize_small.py [cfg] [chains] [nobs]"""
import sys
import numpy as np
import bpinference_b200 as bp
from bpinference_b200 import synthetic as syn
name = sys.argv[1] if len(sys.argv) > 1 else "cfg5U"
chains = int(sys.argv[2]) if len(sys.argv) > 2 else 19
nobs = int(sys.argv[3]) if len(sys.argv) > 3 else 3000
cfg = syn.make_config(name, chains=chains, nobs=nobs)
mod = bp.bp_model(cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"])
eng = bp.stan_model(mod, cfg["priors"], simple_bd=cfg["kind"] == "simple_bd", noise_model=cfg["kind"] == "multitype_noise")
d = eng.data(cfg["data"], grouping=1) if name != "cfg5U" else eng.data(cfg["data"])
u = cfg["upars"].copy()
u[3] += 2.2   # one chain on the ring fallback
lp, g, st = eng.log_prob_grad(d, u)
print(d.path(chains)[0], "ok", float(lp[0]), int((st != 0).sum()))
