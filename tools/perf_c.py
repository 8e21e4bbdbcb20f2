"""cfg5U: centred W path vs bp_fusedw, wall time of bpc_log_prob (host buffers) — quick A/B, not the bench."""
import os, sys, time, numpy as np
import bpinference_b200 as bp
from bpinference_b200 import synthetic as syn
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 512
cfg = syn.make_config("cfg5U", chains=chains, nobs=100000)
mod = bp.bp_model(cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"])
eng = bp.stan_model(mod, cfg["priors"])
N = cfg["data"]["ndatapts"]
res = {}
for label, env in (("centred", {}), ("fusedw", {"BPC_NO_CPATH": "1"})):
    os.environ.update(env)
    d = eng.data(cfg["data"])
    for k in env: del os.environ[k]
    best = 1e9
    for i in range(6):
        t = time.time(); lp, g, st = eng.log_prob_grad(d, cfg["upars"]); dt = time.time() - t
        best = min(best, dt)
    fl = eng.flops_per_chain(d, cfg["upars"][:2])
    res[label] = (lp, g)
    print(label, d.path(chains), "ms %.3f" % (best * 1e3), "evals/s %.3e" % (chains * N / best), "bad", int((st != 0).sum()), "flops/eval %.0f" % (fl[0] / N))
a, b = res["centred"], res["fusedw"]
print("max rel lp diff %.2e" % np.max(np.abs(a[0] - b[0]) / np.abs(b[0])), "max grad diff / max grad %.2e" % (np.max(np.abs(a[1] - b[1])) / np.max(np.abs(b[1]))))
