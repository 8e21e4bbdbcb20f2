import time, numpy as np
import bpinference_b200 as bp
from bpinference_b200 import synthetic as syn
cfg = syn.make_config("cfg2", seed=1, chains=1024)
mod = bp.bp_model(cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"])
eng = bp.stan_model(mod, cfg["priors"])
d = eng.data(cfg["data"])
for pooled in (False, True):
    t0 = time.perf_counter()
    fit = eng.sampling(d, chains=1024, iter=300, warmup=150, control=dict(adapt_delta=0.8, max_treedepth=10), seed=1, pooled_adaptation=pooled)
    dt = time.perf_counter() - t0
    ess = fit.ess_bulk()
    nl = fit.n_leapfrog__ if hasattr(fit, "n_leapfrog__") else None
    print("pooled", pooled, "s %.2f" % dt, "min ess/s %.0f" % (ess.min() / dt), "min ess %.0f" % ess.min(), fit.info, "rhat %.4f" % fit.rhat.max(), "div", int(fit.divergent__.sum()),
          "depth mean %.2f" % fit.treedepth__.mean(), "eps quantiles", np.quantile(fit.stepsize__[:, -1], [0, .1, .5, .9, 1]).round(3), flush=True)
    flat = fit.draws.reshape(-1, fit.draws.shape[2]); print("  mean", flat.mean(0).round(4), "truth", cfg["theta_true"])
