"""small sampler run for compute-sanitizer (memcheck / racecheck): bp_grouped passes + bp_nuts_advance + the pooled-window kernels
(bp_pool_stats / bp_pool_metric / bp_pool_apply / bp_pseudo_rescue), the L-BFGS ascent and the Hessian phase.
usage: python tools/sanitize_sampler.py [chains] [iter]"""
import sys
import numpy as np
import bpinference_b200 as bp
from bpinference_b200 import synthetic as syn
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 16
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cfg = syn.make_config("cfg2", chains=chains, nobs=60)
mod = bp.bp_model(cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"])
eng = bp.stan_model(mod, cfg["priors"])
init = bp.uniform_initialize(np.tile([0.0, 1.0], (6, 1)), chains, rng=np.random.default_rng(1))
fit = eng.sampling(cfg["data"], chains=chains, iter=iters, init=init, control=dict(metric="dense_e", max_treedepth=6), seed=2, pooled_adaptation=True,
                   init_opt_steps=40)
print("sampler ok", fit.info, float(np.nanmax(fit.rhat)))
fit2 = eng.sampling(cfg["data"], chains=chains, iter=iters, control=dict(max_treedepth=6), seed=2)
print("diag per-chain ok", fit2.info["passes"])
