"""cfg5U: variants of the W path, kernel time of bpc_log_prob and agreement of lp / gradient — quick A/B, not the bench.
usage: python tools/perf_c.py [chains] [label=ENV1:V1,ENV2:V2 ...]"""
import os, sys, numpy as np
import bpinference_b200 as bp
from bpinference_b200 import synthetic as syn, _ffi
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 512
variants = [("x", {}), ("c", {"BPC_CX": "0"}), ("fusedw", {"BPC_NO_CPATH": "1"})]
if len(sys.argv) > 2:
    variants = []
    for a in sys.argv[2:]:
        label, _, envs = a.partition("=")
        variants.append((label, dict(e.split(":") for e in envs.split(",") if e)))
cfg = syn.make_config("cfg5U", chains=chains, nobs=100000)
mod = bp.bp_model(cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"])
N = cfg["data"]["ndatapts"]
res = {}
for label, env in variants:
    os.environ.update(env)
    try:
        eng = bp.stan_model(mod, cfg["priors"])
        d = eng.data(cfg["data"])
        _ffi.lib().bpc_kernel_timing(d.handle, 1)
        for i in range(5):
            lp, g, st = eng.log_prob_grad(d, cfg["upars"])
        ms, n = _ffi.C.c_double(), _ffi.C.c_longlong()
        _ffi.lib().bpc_kernel_time_ms(d.handle, 1, _ffi.C.byref(ms), _ffi.C.byref(n))
        t = ms.value / n.value
        fl = eng.flops_per_chain(d, cfg["upars"][:2])
        res[label] = (lp, g)
        print(label, env, d.path(chains), "ms %.3f" % t, "evals/s %.3e" % (chains * N / (t * 1e-3)), "bad", int((st != 0).sum()), "flops/eval %.0f" % (fl[0] / N), flush=True)
        d.close(); eng.close()
    except Exception as e:
        print(label, env, "FAILED", str(e)[:400], flush=True)
    for k in env: del os.environ[k]
labels = list(res)
for l in labels[:-1]:
    a, b = res[l], res[labels[-1]]
    print(l, "vs", labels[-1], "max rel lp diff %.2e" % np.max(np.abs(a[0] - b[0]) / np.abs(b[0])), "max grad diff / max grad %.2e" % (np.max(np.abs(a[1] - b[1])) / np.max(np.abs(b[1]))))
