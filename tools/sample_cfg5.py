"""cfg5U / cfg5G sampling experiment: NUTS from uniform_initialize(0, 1) starting values (what the reference's example scripts do),
progress every few seconds (iterations done, step size, lp, phase), then posterior mean against the simulation parameters.
usage: python tools/sample_cfg5.py [workload] [nobs] [chains] [iter] [budget_s] [pooled 0/1] [init_opt_steps]"""
import ctypes as C, sys, time
import numpy as np
import bpinference_b200 as bp
from bpinference_b200 import _ffi, synthetic as syn

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg5U"
nobs = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
chains = int(sys.argv[3]) if len(sys.argv) > 3 else 64
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 300
budget = float(sys.argv[5]) if len(sys.argv) > 5 else 120.0
pooled = int(sys.argv[6]) if len(sys.argv) > 6 else 1
opt_steps = int(sys.argv[7]) if len(sys.argv) > 7 else 0
dense = int(sys.argv[8]) if len(sys.argv) > 8 else 0
nsub = int(sys.argv[9]) if len(sys.argv) > 9 else 0
cfg = syn.make_config(wl, seed=1, chains=chains, nobs=nobs)
mod = bp.bp_model(cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"])
eng = bp.stan_model(mod, cfg["priors"])
d = eng.data(cfg["data"])
P = cfg["nparams"]
init = bp.uniform_initialize(np.tile([0.0, 1.0], (P, 1)), chains, rng=np.random.default_rng(3))
inits = np.ascontiguousarray([[ch["Theta%d" % (k + 1)] for k in range(P)] for ch in init])
L = _ffi.lib()
d_sub = None
if nsub:
    from bpinference_b200.fit import subsample_data
    d_sub = eng.data(subsample_data(cfg["data"], nsub, 1, False))
args = _ffi.SampleArgs(chains, iters, iters // 2, 0.8, 10, 1, 0, _ffi.dptr(inits), pooled, 0.0, opt_steps, dense, d_sub.handle if d_sub is not None else None)
h = C.c_void_p()
_ffi.check(L.bpc_sampler_create(eng.handle, d.handle, C.byref(args), C.byref(h)))
state = C.c_int(0)
t0 = time.perf_counter()
it, eps, lp, ph = np.zeros(chains, np.int32), np.zeros(chains), np.zeros(chains), np.zeros(chains, np.int32)
last = 0.0
while state.value != 1 and time.perf_counter() - t0 < budget:
    _ffi.check(L.bpc_sampler_run(h, 256, C.byref(state)))
    if state.value == 2:
        _ffi.check(L.bpc_sampler_pool_stats(h, None, 0))
        _ffi.check(L.bpc_sampler_pool_apply(h, None))
    now = time.perf_counter() - t0
    if now - last > 3.0 or state.value == 1:
        last = now
        _ffi.check(L.bpc_sampler_progress(h, _ffi.iptr(it), _ffi.dptr(eps), _ffi.dptr(lp), _ffi.iptr(ph)))
        print("t %6.1f s  passes %7d  iter min/med/max %d/%d/%d  eps med %.3g  lp med %.6g (min %.6g)  phases %s" % (
            now, L.bpc_sampler_passes(h), it.min(), np.median(it), it.max(), np.median(eps), np.median(lp), lp.min(),
            np.bincount(ph, minlength=8).tolist()), flush=True)
dt = time.perf_counter() - t0
print("state", state.value, "seconds %.1f" % dt, "passes", L.bpc_sampler_passes(h), "leapfrogs", L.bpc_sampler_total_leapfrogs(h))
if state.value == 1:
    ns = iters - iters // 2
    dim = eng.dim
    draws = np.empty((chains, ns, dim)); lpd = np.empty((chains, ns)); acc = np.empty((chains, ns)); ep = np.empty((chains, ns))
    div, td, nl = (np.empty((chains, ns), dtype=np.int32) for _ in range(3))
    _ffi.check(L.bpc_sampler_draws(h, _ffi.dptr(draws), _ffi.dptr(lpd), _ffi.iptr(div), _ffi.iptr(td), _ffi.iptr(nl), _ffi.dptr(acc), _ffi.dptr(ep)))
    flat = draws.reshape(-1, dim)
    print("mean ", flat.mean(0).round(5)); print("truth", cfg["theta_true"]); print("sd   ", flat.std(0).round(5))
    print("z    ", ((flat.mean(0) - cfg["theta_true"]) / flat.std(0)).round(2))
    cm = draws.mean(1); W = draws.var(1, ddof=1).mean(0); B = cm.var(0, ddof=1)
    print("rhat ", np.sqrt((W * (ns - 1) / ns + B) / W).round(3), "div", int(div.sum()), "treedepth mean %.2f" % td.mean(), "nleap mean %.1f" % nl.mean(),
          "accept %.3f" % acc.mean(), "eps q", np.quantile(ep[:, -1], [0, .5, 1]).round(5))
L.bpc_sampler_free(h)
