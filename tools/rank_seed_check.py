"""The bench's sampling block draws the starting values of rank r from default_rng(100 + r).  Runs the cfg5U and cfg2 sampling of
bench.py on ONE GPU with the starting values ranks 4 .. 7 of an 8-GPU run would draw (every rank adapting on its own chains only:
a harder setting than the pooled one), to see that no rank's starting values stall the run.   python tools/rank_seed_check.py"""
import sys
import time

import numpy as np

import bpinference_b200 as bp
from bpinference_b200 import synthetic as syn


def run(workload, chains, iters, warm, opt_steps, dense, subsample, r):
    cfg = syn.make_config(workload, seed=1, chains=chains, nobs=None)
    mod = bp.bp_model(cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"])
    eng = bp.stan_model(mod, cfg["priors"], simple_bd=cfg["kind"] == "simple_bd", noise_model=cfg["kind"] == "multitype_noise")
    d = eng.data(cfg["data"], device=0)
    P = cfg["nparams"]
    init = bp.uniform_initialize(np.tile([0.0, 1.0], (P, 1)), chains, rng=np.random.default_rng(100 + r))
    t0 = time.perf_counter()
    fit = eng.sampling(d, chains=chains, iter=iters, warmup=warm, init=init, control=dict(adapt_delta=0.8, max_treedepth=10, metric="dense_e" if dense else "diag_e"),
                       seed=1, pooled_adaptation=True, init_opt_steps=opt_steps, init_opt_subsample=subsample)
    dt = time.perf_counter() - t0
    print("%s rank %d: %.2f s  max rhat %.4f  divergent %d  min ESS %.0f  passes %d" % (workload, r, dt, float(np.nanmax(fit.rhat)), int(fit.divergent__.sum()),
                                                                                float(np.min(fit.ess_bulk())), fit.info["passes"]), flush=True)


if __name__ == "__main__":
    ranks = [int(a) for a in sys.argv[1:]] or [4, 5, 6, 7]
    for r in ranks:
        run("cfg5U", 512, 200, 100, 300, True, 500, r)
    for r in ranks:
        run("cfg2", 1024, 300, 150, 200, False, 0, r)
