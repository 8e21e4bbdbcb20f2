"""bp_grouped on cfg2 with 1024 chains: evaluation time when all, one or no chain needs sub-steps (the cooperative multi-step
route against the rates near the truth): what a sampler pass costs during warm-up."""
import time, torch, numpy as np
import bpinference_b200 as bp
from bpinference_b200 import synthetic as syn, _ffi
L = _ffi.lib()
chains = 1024
cfg = syn.make_config("cfg2", chains=chains)
mod = bp.bp_model(cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"])
eng = bp.stan_model(mod, cfg["priors"])
print("priors", cfg["priors"][0])
d = eng.data(cfg["data"])
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
def run(upars, label):
    u = torch.from_numpy(np.ascontiguousarray(upars)).to(dev); lp = torch.empty(chains, dtype=torch.float64, device=dev)
    g = torch.empty(chains, eng.dim, dtype=torch.float64, device=dev); st = torch.empty(chains, dtype=torch.int32, device=dev)
    s = torch.cuda.Stream(); torch.cuda.set_stream(s)
    def go(n):
        for _ in range(n):
            _ffi.check(L.bpc_log_prob_dev(eng.handle, d.handle, u.data_ptr(), chains, 1, lp.data_ptr(), g.data_ptr(), st.data_ptr(), s.cuda_stream))
    go(20); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s); go(500); e1.record(s); torch.cuda.synchronize()
    th = eng.constrain_pars(upars)
    print("%-28s %.1f us per evaluation; rates max %.2f median %.2f" % (label, e0.elapsed_time(e1) / 500 * 1e3, th.max(), np.median(th)), flush=True)
base = cfg["upars"]
run(base, "near the truth")
run(rng.uniform(-2, 2, base.shape), "inits uniform(-2, 2)")
u1 = base.copy(); u1[0] = rng.uniform(-2, 2, base.shape[1]); run(u1, "one chain at an init point")
u2 = base.copy(); u2[0] += 1.5; run(u2, "one chain with rates x ~3")
u3 = base.copy(); u3[0] += 2.5; run(u3, "one chain with rates x ~6")
