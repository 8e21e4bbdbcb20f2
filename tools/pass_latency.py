"""latency of one log-density evaluation (device buffers, no host sync in the loop) on small problems: one launch on the grouped
and single-tile one-type paths, prologue + kernel + epilogue otherwise"""
import time, torch, numpy as np
import bpinference_b200 as bp
from bpinference_b200 import synthetic as syn, _ffi
L = _ffi.lib()
for name, chains in [("cfg2", 64), ("cfg2", 1024), ("cfg4", 8), ("cfg1", 4)]:
    cfg = syn.make_config(name, chains=chains)
    mod = bp.bp_model(cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"])
    eng = bp.stan_model(mod, cfg["priors"], simple_bd=cfg["kind"] == "simple_bd", noise_model=cfg["kind"] == "multitype_noise")
    for grouping in ((0,) if cfg["kind"] == "simple_bd" else (2, 1)):
        d = eng.data(cfg["data"], grouping=grouping)
        dev = torch.device("cuda", 0)
        u = torch.from_numpy(cfg["upars"]).to(dev); lp = torch.empty(chains, dtype=torch.float64, device=dev)
        g = torch.empty(chains, eng.dim, dtype=torch.float64, device=dev); st = torch.empty(chains, dtype=torch.int32, device=dev)
        s = torch.cuda.Stream(); torch.cuda.set_stream(s)
        def run(n):
            for _ in range(n):
                _ffi.check(L.bpc_log_prob_dev(eng.handle, d.handle, u.data_ptr(), chains, 1, lp.data_ptr(), g.data_ptr(), st.data_ptr(), s.cuda_stream))
        run(50); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record(s); run(2000); e1.record(s); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print("%s chains %4d grouping %d: %.1f us per evaluation on the device, %.1f us wall (host launch rate)" % (name, chains, grouping, e0.elapsed_time(e1) / 2000 * 1e3, dt / 2000 * 1e6), flush=True)
