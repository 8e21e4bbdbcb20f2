"""cfg2 sampling, 1024 chains x 300 iterations: per-chain against pooled warm-up adaptation (ESS/s, divergences, step sizes).
BPC_SAMPLER_TRACE=1 prints the wall time per block of passes.  usage: python tools/pooled_vs_perchain.py [pooled-only]"""
import sys, time, numpy as np
import bpinference_b200 as bp
from bpinference_b200 import synthetic as syn
cfg = syn.make_config("cfg2", seed=1, chains=1024)
mod = bp.bp_model(cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"])
eng = bp.stan_model(mod, cfg["priors"])
d = eng.data(cfg["data"])
for pooled in ((True,) if len(sys.argv) > 1 else (False, True)):
    t0 = time.perf_counter()
    fit = eng.sampling(d, chains=1024, iter=300, warmup=150, control=dict(adapt_delta=0.8, max_treedepth=10), seed=1, pooled_adaptation=pooled)
    dt = time.perf_counter() - t0
    ess = fit.ess_bulk()
    print("pooled", pooled, "s %.2f" % dt, "min ess/s %.0f" % (ess.min() / dt), "min ess %.0f" % ess.min(), fit.info, "rhat %.4f" % fit.rhat.max(), "div", int(fit.divergent__.sum()),
          "depth mean %.2f" % fit.treedepth__.mean(), "eps quantiles", np.quantile(fit.stepsize__[:, -1], [0, .1, .5, .9, 1]).round(3), flush=True)
    flat = fit.draws.reshape(-1, fit.draws.shape[2]); print("  mean", flat.mean(0).round(4), "truth", cfg["theta_true"])
