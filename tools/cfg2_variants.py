"""cfg2 (two-type example shape, 1024 chains): ESS/s of a few sampler configurations.  usage: python tools/cfg2_variants.py"""
import time
import numpy as np
import bpinference_b200 as bp
from bpinference_b200 import synthetic as syn
cfg = syn.make_config("cfg2", seed=1, chains=1024)
mod = bp.bp_model(cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"])
eng = bp.stan_model(mod, cfg["priors"])
d = eng.data(cfg["data"])
init = bp.uniform_initialize(np.tile([0.0, 1.0], (6, 1)), 1024, rng=np.random.default_rng(100))
variants = [("A diag pooled, random init (r01)", dict(iter=300, warmup=150, init="random", control=dict(adapt_delta=0.8), pooled_adaptation=True)),
            ("B diag pooled + ascent", dict(iter=300, warmup=150, init=init, control=dict(adapt_delta=0.8), pooled_adaptation=True, init_opt_steps=200)),
            ("C dense pooled + ascent", dict(iter=300, warmup=150, init=init, control=dict(adapt_delta=0.8, metric="dense_e"), pooled_adaptation=True, init_opt_steps=200)),
            ("D dense pooled + ascent, warmup 300", dict(iter=450, warmup=300, init=init, control=dict(adapt_delta=0.8, metric="dense_e"), pooled_adaptation=True, init_opt_steps=200)),
            ("E dense pooled + ascent, delta .9", dict(iter=300, warmup=150, init=init, control=dict(adapt_delta=0.9, metric="dense_e"), pooled_adaptation=True, init_opt_steps=200)),
            ("F rstan defaults (per-chain diag)", dict(iter=300, warmup=150, init=init, control=dict(adapt_delta=0.8)))]
for name, kw in variants:
    t0 = time.perf_counter()
    fit = eng.sampling(d, chains=1024, seed=1, **kw)
    dt = time.perf_counter() - t0
    ess = fit.ess_bulk()
    flat = fit.draws.reshape(-1, 6)
    print("%-40s %.2f s  min ESS %.0f  ESS/s %.0f  rhat %.3f  div %d  depth %.2f  nleap %.1f  eps med %.4f  passes %d  mean %s" % (
        name, dt, ess.min(), ess.min() / dt, np.nanmax(fit.rhat), int(fit.divergent__.sum()), fit.treedepth__.mean(), fit.n_leapfrog__.mean(),
        np.median(fit.stepsize__[:, -1]), fit.info["passes"], flat.mean(0).round(3)), flush=True)
