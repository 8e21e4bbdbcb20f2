"""tuning sweep of the centred W path (bp_fusedc): chunks per group, groups per CTA, CTAs per SM, CTA size"""
import os, sys
import numpy as np
import bpinference_b200 as bp
from bpinference_b200 import synthetic as syn, _ffi
cfg = syn.make_config("cfg5U", chains=512, nobs=100000)
mod = bp.bp_model(cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"])
combos = [(128, 3, 8, 16), (128, 3, 4, 16), (128, 3, 16, 16), (128, 3, 8, 8), (128, 3, 8, 32), (128, 2, 8, 16), (128, 4, 8, 16),
          (64, 6, 8, 8), (64, 4, 8, 8), (256, 1, 8, 32), (256, 2, 8, 32), (128, 3, 6, 16), (128, 3, 12, 16)]
if len(sys.argv) > 1:
    combos = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
for combo in combos:
    thr, cmin, chunks, gpc = combo[:4]
    os.environ["BPC_CSPLIT"] = str(combo[4]) if len(combo) > 4 else "1"
    os.environ["BPC_THREADS"] = str(thr); os.environ["BPC_CMINBLOCKS"] = str(cmin)
    os.environ["BPC_CGROUP_CHUNKS"] = str(chunks); os.environ["BPC_CGROUPS_PER_CTA"] = str(gpc)
    try:
        eng = bp.stan_model(mod, cfg["priors"])
        d = eng.data(cfg["data"])
        _ffi.lib().bpc_kernel_timing(d.handle, 1)
        for i in range(4):
            lp, g, st = eng.log_prob_grad(d, cfg["upars"])
        ms, n = _ffi.C.c_double(), _ffi.C.c_longlong()
        _ffi.lib().bpc_kernel_time_ms(d.handle, 1, _ffi.C.byref(ms), _ffi.C.byref(n))
        print("split %s " % os.environ["BPC_CSPLIT"] + "threads %3d cminblocks %d chunks/group %2d groups/CTA %2d (%s tiles): %.3f ms  evals/s %.3e  lp0 %.6f bad %d" % (thr, cmin, chunks, gpc, d.path(512)[1], ms.value / n.value, 512 * 100000 / (ms.value / n.value * 1e-3), lp[0], int((st != 0).sum())), flush=True)
        d.close(); eng.close()
    except Exception as e:
        print("threads", thr, "cmin", cmin, "chunks", chunks, "gpc", gpc, "FAILED", str(e)[:300], flush=True)
