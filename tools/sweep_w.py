"""tuning sweep of the W-path kernel (bp_fusedw): observations per thread, CTAs per SM, CTA size"""
import os, sys
import numpy as np
import bpinference_b200 as bp
from bpinference_b200 import synthetic as syn, _ffi
cfg = syn.make_config("cfg5U", chains=512, nobs=100000)
mod = bp.bp_model(cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"])
for thr, wmin, opt in [(128, 3, 32), (128, 3, 64), (128, 3, 128), (128, 2, 32), (128, 4, 32), (64, 6, 64), (64, 4, 64), (256, 1, 32), (256, 2, 32), (96, 4, 64)]:
    os.environ["BPC_THREADS"] = str(thr); os.environ["BPC_OBS_PER_THREAD"] = str(opt); os.environ["BPC_WMINBLOCKS"] = str(wmin)
    try:
        eng = bp.stan_model(mod, cfg["priors"])
        d = eng.data(cfg["data"])
        _ffi.lib().bpc_kernel_timing(d.handle, 1)
        for i in range(4):
            lp, g, st = eng.log_prob_grad(d, cfg["upars"])
        ms, n = _ffi.C.c_double(), _ffi.C.c_longlong()
        _ffi.lib().bpc_kernel_time_ms(d.handle, 1, _ffi.C.byref(ms), _ffi.C.byref(n))
        print("threads %3d wminblocks %d obs/thread %3d : %.3f ms  evals/s %.3e  lp0 %.6f bad %d" % (thr, wmin, opt, ms.value / n.value, 512 * 100000 / (ms.value / n.value * 1e-3), lp[0], int((st != 0).sum())), flush=True)
        d.close(); eng.close()
    except Exception as e:
        print("threads", thr, "wmin", wmin, "opt", opt, "FAILED", str(e)[:300], flush=True)
