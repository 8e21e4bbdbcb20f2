"""N > 1 host logic on CPU: two gloo ranks exchange what the GPU ranks exchange over NCCL — the chain moments
behind R-hat and the pooled warm-up statistics — and the reference arm of bench.py under a 2-rank launch."""
import json
import os
import subprocess
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bpinference_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _chain_moments(draws):
    """numpy twin of bp_chain_moments (sampler.cu): per dimension (count, sum mean, sum mean^2, sum var)."""
    m, v = draws.mean(axis=1), draws.var(axis=1, ddof=1)
    return np.stack([np.full(draws.shape[2], draws.shape[0], dtype=float), m.sum(0), (m * m).sum(0), v.sum(0)], axis=1).ravel()


def _pool_stats(q):
    """numpy twin of bp_pool_stats: (count, sum q, sum q^2) over all chains and window draws of one rank."""
    flat = q.reshape(-1, q.shape[2])
    return np.concatenate([[flat.shape[0]], flat.sum(0), (flat * flat).sum(0)])


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    all_draws = rng.normal(size=(8, 200, 3)) * [1.0, 2.0, 0.5] + rng.normal(size=(8, 1, 3)) * 0.2      # 8 chains in total
    mine = all_draws[rank * 4:(rank + 1) * 4]                                                          # sharded: 4 per rank
    mom = torch.from_numpy(_chain_moments(mine))
    dist.all_reduce(mom)
    rhat = np.empty(3)
    _ffi.check(_ffi.lib().bpc_rhat(_ffi.dptr(np.ascontiguousarray(mom.numpy())), 3, 200, _ffi.dptr(rhat)))
    pool = torch.from_numpy(_pool_stats(mine))
    dist.all_reduce(pool)
    n = pool[0].item()
    var = ((pool[4:7] - pool[1:4] ** 2 / n) / (n - 1)).numpy()
    if rank == 0:
        q.put((rhat, var, all_draws))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_rhat_and_pooled_variance():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    rhat, var, draws = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process statistics over all 8 chains
    n = draws.shape[1]
    W = draws.var(axis=1, ddof=1).mean(axis=0)
    B_n = draws.mean(axis=1).var(axis=0, ddof=1)
    np.testing.assert_allclose(rhat, np.sqrt((W * (n - 1) / n + B_n) / W), rtol=1e-12)
    np.testing.assert_allclose(var, draws.reshape(-1, 3).var(axis=0, ddof=1), rtol=1e-10)


def test_reference_arm_under_two_ranks():
    """bench.py --impl reference launched the way the driver launches N > 1: rank 0 alone runs and prints one JSON line."""
    port = 31000 + os.getpid() % 2000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "1", "--warmup", "0", "--impl", "reference", "--workload", "cfg1", "--cpu-seconds", "2"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["value"] > 0 and j["cpu_baseline"]["kind"] == "port" and j["e2e"]["h2d_bytes_per_step"] == 0
