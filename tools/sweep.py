"""tuning sweep of the fused kernel: ring size, CTA size, CTAs per SM, observations per thread (env knobs of codegen)."""
import os, sys, time, itertools
import numpy as np
import bpinference_b200 as bp
from bpinference_b200 import synthetic as syn, _ffi
name = sys.argv[1] if len(sys.argv) > 1 else "cfg5U"
cfg = syn.make_config(name, chains=512, nobs=100000)
mod = bp.bp_model(cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"])
combos = [(12, 128, 0, 8), (12, 128, 0, 32), (8, 128, 3, 8), (8, 128, 3, 32), (8, 128, 2, 8), (6, 128, 4, 16), (6, 128, 3, 16), (8, 96, 4, 16), (8, 64, 6, 16), (12, 64, 4, 16), (8, 192, 2, 16), (6, 256, 2, 16), (24, 128, 1, 16)]
for mslot, thr, minb, opt in combos:
    os.environ["BPC_MSLOT"] = str(mslot); os.environ["BPC_THREADS"] = str(thr); os.environ["BPC_OBS_PER_THREAD"] = str(opt)
    if minb: os.environ["BPC_MINBLOCKS"] = str(minb)
    else: os.environ.pop("BPC_MINBLOCKS", None)
    try:
        eng = bp.stan_model(mod, cfg["priors"])
        d = eng.data(cfg["data"])
        _ffi.lib().bpc_kernel_timing(d.handle, 1)
        for i in range(4):
            lp, g, st = eng.log_prob_grad(d, cfg["upars"])
        ms, n = _ffi.C.c_double(), _ffi.C.c_longlong()
        _ffi.lib().bpc_kernel_time_ms(d.handle, 1, _ffi.C.byref(ms), _ffi.C.byref(n))
        src = eng.cuda_source
        mb = [l for l in src.splitlines() if "BP_MINBLOCKS" in l][0].split()[-1]
        print("mslot %2d threads %3d minblocks %s obs/thread %2d : kernel %.3f ms  evals/s %.3e  lp0 %.6f bad %d" % (mslot, thr, mb, opt, ms.value / n.value, 512 * 100000 / (ms.value / n.value * 1e-3), lp[0], int((st != 0).sum())), flush=True)
        d.close(); eng.close()
    except Exception as e:
        print("mslot", mslot, "threads", thr, "minblocks", minb, "FAILED", str(e)[:200], flush=True)
