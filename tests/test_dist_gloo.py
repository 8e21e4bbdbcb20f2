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
    """numpy twin of bp_chain_moments (sampler.cu): per dimension (count, sum mean, sum mean^2, sum var) over the HALF chains."""
    h = draws.shape[1] // 2
    halves = np.concatenate([draws[:, :h], draws[:, draws.shape[1] - h:]], axis=0)
    m, v = halves.mean(axis=1), halves.var(axis=1, ddof=1)
    return np.stack([np.full(draws.shape[2], halves.shape[0], dtype=float), m.sum(0), (m * m).sum(0), v.sum(0)], axis=1).ravel()


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
    # single-process split R-hat over all 8 chains (16 half chains)
    h = draws.shape[1] // 2
    draws_h = np.concatenate([draws[:, :h], draws[:, draws.shape[1] - h:]], axis=0)
    n = h
    W = draws_h.var(axis=1, ddof=1).mean(axis=0)
    B_n = draws_h.mean(axis=1).var(axis=0, ddof=1)
    np.testing.assert_allclose(rhat, np.sqrt((W * (n - 1) / n + B_n) / W), rtol=1e-12)
    np.testing.assert_allclose(var, draws.reshape(-1, 3).var(axis=0, ddof=1), rtol=1e-10)


def _pool_worker(rank, world, port, q):
    """the pooled warm-up protocol of fit.pooled_loop with ranks that stop at different times: rank 1's chains all "fail" at once
    (state 1 from the first call), rank 0 goes through three windows and then finishes."""
    from bpinference_b200.fit import pooled_loop
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    script = [2, 2, 2, 1] if rank == 0 else [1]
    calls = {"run": 0, "apply": 0}
    buf = torch.zeros(4, dtype=torch.float64)

    def run():
        st = script[min(calls["run"], len(script) - 1)]
        calls["run"] += 1
        return st

    def exchange():
        st = script[min(calls["run"] - 1, len(script) - 1)]
        buf[:] = torch.tensor([10.0 if st == 2 else 0.0, 1.0 + rank if st == 2 else 0.0, 2.0 if st == 2 else 0.0, 1.0 if st == 1 else 0.0], dtype=torch.float64)
        dist.all_reduce(buf)
        return float(buf[3])

    def apply():
        calls["apply"] += 1
        assert buf[0] == 10.0        # only rank 0's window contributes: rank 1 sends zeros

    n = pooled_loop(run, exchange, apply, world)
    q.put((rank, n, calls["run"], calls["apply"]))
    dist.barrier()
    dist.destroy_process_group()


def test_pooled_protocol_does_not_deadlock_when_a_rank_finishes_early():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33000 + os.getpid() % 2000
    procs = [ctx.Process(target=_pool_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0] == (0, 4, 4, 3) and res[1] == (1, 4, 4, 0)     # both ranks made the same 4 exchanges; only rank 0 applied windows


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
    # the thread count does not follow the OMP_NUM_THREADS=1 torchrun exports to its workers (the N > 1 ratios were void)
    assert j["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    import bench
    assert j["config"] == bench.bench_config("cfg1", 4, 475)
