"""short sampling run for kernel-level timing of a sampler pass (ncu launch list)"""
import os
os.environ["BPC_NO_GRAPH"] = "1"
import numpy as np
import bpinference_b200 as bp
from bpinference_b200 import synthetic as syn
cfg = syn.make_config("cfg2", seed=1, chains=1024)
mod = bp.bp_model(cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"])
eng = bp.stan_model(mod, cfg["priors"])
import sys
it = int(sys.argv[1]) if len(sys.argv) > 1 else 20
fit = eng.sampling(cfg["data"], chains=1024, iter=it, warmup=it // 2, seed=1)
print(fit.info)
