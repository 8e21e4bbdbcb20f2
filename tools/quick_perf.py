"""kernel time and evaluations per second of every synthetic config on one GPU (quick look, not the bench)."""
import time, numpy as np, json
import bpinference_b200 as bp
from bpinference_b200 import synthetic as syn
print("fp64 peak", bp.fp64_peak(0, 300.0))
for name, chains, nobs in [("cfg5U", 512, 100000), ("cfg5K", 512, 100000), ("cfg5G", 512, 100000), ("cfg2", 1024, None), ("cfg4", 4, None), ("cfg1", 4, None)]:
    cfg = syn.make_config(name, chains=chains, nobs=nobs)
    mod = bp.bp_model(cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"])
    eng = bp.stan_model(mod, cfg["priors"], simple_bd=cfg["kind"]=="simple_bd", noise_model=cfg["kind"]=="multitype_noise")
    d = eng.data(cfg["data"])
    print(name, "grouped", d.grouped, "groups", d.ngroups)
    for i in range(3):
        t=time.time(); lp, g, st = eng.log_prob_grad(d, cfg["upars"]); dt=time.time()-t
    N = cfg["data"]["ndatapts"]
    fl = eng.flops_per_chain(d, cfg["upars"][:2])
    print(name, "evals/s %.3e" % (chains*N/dt), "ms %.2f" % (dt*1e3), "bad", int((st!=0).sum()), "flops/eval", fl[0]/N, "TF %.2f" % (fl[0]*chains/dt/1e12))
