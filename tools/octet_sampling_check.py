"""sampling on the octet path (ungrouped data, >= 16 chains): the CUDA-graph replay of the passes against eager launches
(BPC_NO_GRAPH=1) -- same draws -- and the posterior mean against the truth.  usage: python tools/octet_sampling_check.py"""
import os, time, numpy as np
import bpinference_b200 as bp
from bpinference_b200 import synthetic as syn
cfg = syn.make_config("cfg5U", seed=1, chains=16, nobs=20000)
mod = bp.bp_model(cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"])
out = {}
for label, env in (("graph", None), ("eager", "1")):
    if env: os.environ["BPC_NO_GRAPH"] = env
    eng = bp.stan_model(mod, cfg["priors"])
    d = eng.data(cfg["data"])
    assert "bp_fusedo" in d.path(16)[0], d.path(16)
    t0 = time.perf_counter()
    fit = eng.sampling(d, chains=16, iter=40, warmup=20, seed=3)
    out[label] = fit.draws.copy()
    print(label, "s %.2f" % (time.perf_counter() - t0), fit.info, "div", int(fit.divergent__.sum()), "mean", fit.draws.reshape(-1, fit.draws.shape[2]).mean(0).round(3), flush=True)
    os.environ.pop("BPC_NO_GRAPH", None)
print("truth", cfg["theta_true"])
print("graph == eager:", bool((out["graph"] == out["eager"]).all()))
