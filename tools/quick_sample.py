"""short sampling runs of the small configs (cfg1, cfg2, cfg4, cfg3): posterior means against the truth, R-hat, ESS, step sizes."""
import time, numpy as np
import bpinference_b200 as bp
from bpinference_b200 import synthetic as syn
np.set_printoptions(precision=4, linewidth=200)
for name, chains, it in [("cfg1", 4, 2000), ("cfg2", 64, 1000), ("cfg4", 8, 1000), ("cfg3", 4, 1000)]:
    cfg = syn.make_config(name, chains=chains)
    mod = bp.bp_model(cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"])
    eng = bp.stan_model(mod, cfg["priors"], simple_bd=cfg["kind"]=="simple_bd", noise_model=cfg["kind"]=="multitype_noise")
    t=time.time()
    fit = eng.sampling(cfg["data"], chains=chains, iter=it, control=dict(adapt_delta=0.9), seed=3)
    dt=time.time()-t
    flat = fit.draws.reshape(-1, fit.draws.shape[2])
    print(name, "s %.2f"%dt, fit.info, "div", bp.check_div(fit), bp.check_exceeded_treedepth(fit))
    print("  truth", cfg["theta_true"], cfg.get("sigma_obs_true"))
    print("  mean ", flat.mean(0)); print("  sd   ", flat.std(0)); print("  rhat ", fit.rhat); print("  ess  ", fit.ess_bulk())
    print("  eps", fit.stepsize__[:, -1][:4], "depth mean", fit.treedepth__.mean(), "acc", fit.accept_stat__.mean())
