"""multi-rank sampling check (torchrun): chains sharded over ranks, pooled warm-up windows and R-hat over NCCL."""
import os, time
import numpy as np
import torch, torch.distributed as dist
import bpinference_b200 as bp
from bpinference_b200 import synthetic as syn
rank, lr = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
cfg = syn.make_config("cfg2", chains=64, nobs=120)
mod = bp.bp_model(cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"])
eng = bp.stan_model(mod, cfg["priors"])
d = eng.data(cfg["data"], device=lr)
for pooled in (False, True):
    t = time.time()
    fit = eng.sampling(d, chains=64, iter=400, control=dict(adapt_delta=0.85), seed=4, pooled_adaptation=pooled)
    dt = time.time() - t
    m = torch.tensor(fit.draws.reshape(-1, 6).mean(0), device="cuda")
    allm = [torch.zeros_like(m) for _ in range(dist.get_world_size())]
    dist.all_gather(allm, m)
    if rank == 0:
        print("pooled", pooled, "s %.1f" % dt, "total_chains", fit.info["total_chains"], "rhat max %.3f" % fit.rhat.max(), "passes", fit.info["passes"])
        print("  rank means", [np.round(x.cpu().numpy(), 3) for x in allm])
dist.barrier()
dist.destroy_process_group()
