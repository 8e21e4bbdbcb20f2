#!/usr/bin/env python
"""bench.py — fused log-density + gradient throughput of libbpcuda on B200, per the driver's contract.

A "step" is ONE evaluation of the posterior log density and its gradient for every chain of this rank over
every observation (what one leapfrog step of every chain costs).  Unit of the metric: (chain x observation)
evaluations per second, whole job.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg5U|cfg5K|cfg5G|cfg2|cfg1|cfg3|cfg4] [--impl reference]

N > 1: launched by torchrun, one rank per GPU; chains are sharded (weak scaling: --chains-per-gpu chains on
each rank), observations and the compiled model are replicated, no collective on the data path; NCCL is only
used for the barrier and the max-over-ranks of the device time.

--impl reference: the reference arm.  jproney/bpinference is R + Stan and neither R nor rstan exists in this
image, so the arm times the CPU restatement of the same path (oracle/, kind "port") with all host threads on
a bounded sample of the same workload, rank 0 only.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fused logp+grad (chain x observation) evaluations per second"
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from one `ncu --set full` capture of the default
# workload (profiles/r02_ncu_bp_fusedo_cfg5U.txt: 19.6 MB read + 54.7 MB written — the per-(octet, group) S' matrices, 102 MB of
# scratch per launch, are written once and streamed back by bp_toep, partly still from L2; the observation table is 5.6 MB).
# Algorithmic bytes are 5.68 MB (table + per-chain input / output): the traffic is ~13 x that, all of it the kernel's own scratch,
# at 0.5 % of the HBM bandwidth.
NCU_TRAFFIC_BYTES = {("cfg5U", 512, 100000): 19.596e6 + 54.739e6}
UNIT = "evals/s"
DEFAULT_CHAINS = {"cfg5U": 512, "cfg5K": 512, "cfg5G": 512, "cfg2": 1024, "cfg1": 4, "cfg3": 4, "cfg4": 4}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg5U")
    ap.add_argument("--chains-per-gpu", type=int, default=0)
    ap.add_argument("--nobs", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU work of the cpu_baseline / reference sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sampling", action="store_true", help="skip the NUTS runs that report ESS/s (all ranks, pooled warm-up over NCCL)")
    ap.add_argument("--sampling-iter", type=int, default=200, help="iterations (half warm-up) of the NUTS run on the main workload")
    ap.add_argument("--no-other-workloads", action="store_true", help="skip the other shapes of SURVEY.md §8(d)")
    ap.add_argument("--no-parity", action="store_true", help="skip the lp / gradient check of 8 chains against the CPU port")
    return ap.parse_args()


def workload_description(name, chains, N):
    return {"cfg5U": "synthetic 3-type process (9 events, 9 rates), %d observations each with its own duration, %d chains per GPU" % (N, chains),
            "cfg5K": "cfg5U plus one dependent variable per observation (rate 1 = c[1]*exp(-c[10]*x[1])), %d obs, %d chains per GPU" % (N, chains),
            "cfg5G": "synthetic 3-type process, %d observations in 6 (level, duration) groups, %d chains per GPU" % (N, chains),
            "cfg2": "two-type birth-death-transition (examples/two_type.R), %d observations, %d chains per GPU" % (N, chains),
            "cfg1": "one-type simple birth-death (examples/one_type.R), %d observations, %d chains" % (N, chains),
            "cfg3": "one-type logistic birth vs dose (examples/one_type_logistic_birth.R), %d observations, %d chains" % (N, chains),
            "cfg4": "apoptosis with observation noise (examples/apoptosis_with_noise.R), %d observations, %d chains" % (N, chains)}[name]


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe): NVML polled every 5 ms from a
    thread (a timed region of tens of milliseconds sees many samples); `nvidia-smi -lms 100` when NVML cannot be loaded."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nvml, self.stopflag = index, [], None, None, False

    def _poll_nvml(self):
        n, h = self.nvml
        bits = [("hw_slowdown", n.nvmlClocksThrottleReasonHwSlowdown), ("hw_thermal_slowdown", n.nvmlClocksThrottleReasonHwThermalSlowdown),
                ("sw_thermal_slowdown", n.nvmlClocksThrottleReasonSwThermalSlowdown), ("sw_power_cap", n.nvmlClocksThrottleReasonSwPowerCap)]
        mx = n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)
        while not self.stopflag:
            try:
                r = n.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append("%d, %d, %.2f, %s" % (n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM), mx, n.nvmlDeviceGetPowerUsage(h) / 1000.0,
                                                       ", ".join("Active" if r & b else "Not Active" for _, b in bits)))
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        try:
            import pynvml as n
            n.nvmlInit()
            self.nvml = (n, n.nvmlDeviceGetHandleByIndex(self.index))
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            self.proc = True
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        if self.nvml:
            self.stopflag = True
            self.thread.join(timeout=1)
        else:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def bench_config(workload, chains, N):
    """the `config` object of the JSON line: identical in both arms (the driver compares them)."""
    return {"workload": workload + ": " + workload_description(workload, chains, N), "chains_per_gpu": chains, "observations": N,
            "l2": "256 MiB buffer rewritten between the timed steps (L2 flush, outside each step's event pair); the 5.6 MB "
                  "observation table is re-read by every chain of a step and is L2-resident by design",
            "reference_arm": "jproney/bpinference is R + Stan (rstan); neither exists in this image, so --impl reference times the CPU "
                             "restatement of the same path (oracle/, kind \"port\": closed-form moments, strictly less work than rstan's "
                             "RK45 forward-sensitivity solve) with every host core on a bounded sample of this workload"}


def cpu_port_rate(cfg, seconds, nthreads=0):
    """times the CPU restatement (oracle/) on a bounded sample of the workload: returns (evals/s, cores, sample text)."""
    from oracle import bp_oracle as bo
    om = bo.OracleModel(cfg["kind"], cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"], cfg["priors"])
    # every host core this process may run on, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1 to its workers)
    cores = host_cores() if nthreads == 0 else nthreads
    d = cfg["data"]
    N = int(d["ndatapts"])

    def sub(nobs):
        if nobs >= N:
            return d
        s = dict(d)
        s["ndatapts"] = nobs
        for k in ("pop_vec", "init_pop"):
            a = np.asarray(d[k])
            s[k] = a[:nobs]
        s["var_idx"] = np.asarray(d["var_idx"])[:nobs]
        if "times_idx" in d:
            s["times_idx"] = np.asarray(d["times_idx"])[:nobs]
        else:
            s["times"] = np.asarray(d["times"])[:nobs]
        return s
    chains = max(cores, 1)
    u = np.resize(cfg["upars"], (chains, cfg["upars"].shape[1]))
    probe_n = min(N, 2000)
    t0 = time.perf_counter()
    bo.log_prob(om, sub(probe_n), u, grouped=False if cfg["name"] in ("cfg5U", "cfg5K") else None, nthreads=cores)
    bo.log_prob(om, sub(probe_n), u, grouped=False if cfg["name"] in ("cfg5U", "cfg5K") else None, nthreads=cores)
    rate = 2 * chains * probe_n / (time.perf_counter() - t0)
    nobs = int(min(N, max(probe_n, rate * seconds / chains)))
    reps = max(1, int(round(rate * seconds / (chains * nobs))))
    t0 = time.perf_counter()
    for _ in range(reps):
        bo.log_prob(om, sub(nobs), u, grouped=False if cfg["name"] in ("cfg5U", "cfg5K") else None, nthreads=cores)
    dt = time.perf_counter() - t0
    return chains * nobs * reps / dt, cores, "%d chains x first %d of %d observations x %d passes, %d OpenMP threads, %.1f s" % (chains, nobs, N, reps, cores, dt)


def sampling_report(bp, syn, torch, dist, rank, world, local_rank, workload, chains, nobs, iters, warm, opt_steps, dense, subsample=0, adapt_delta=0.8):
    """the second half of BASELINE.json's metric: ESS/s of the device-resident NUTS sampler.  Chains are sharded over the ranks
    (`chains` per GPU); the diagonal metric of every warm-up window is estimated from the draws of ALL chains of ALL ranks — the one
    collective of the path: an all-reduce of 2 + 2 dim doubles over NCCL at every window boundary, and 4 dim doubles for split R-hat at
    the end.  Starting values as the reference's example scripts draw them: uniform_initialize(0, 1) (examples/two_type.R:22-23)."""
    cfg = syn.make_config(workload, seed=1, chains=chains, nobs=nobs)
    mod = bp.bp_model(cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"])
    eng = bp.stan_model(mod, cfg["priors"], simple_bd=cfg["kind"] == "simple_bd", noise_model=cfg["kind"] == "multitype_noise")
    d = eng.data(cfg["data"], device=local_rank)
    P = cfg["nparams"]
    init = bp.uniform_initialize(np.tile([0.0, 1.0], (P, 1)), chains, rng=np.random.default_rng(100 + rank))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fit = eng.sampling(d, chains=chains, iter=iters, warmup=warm, init=init, control=dict(adapt_delta=adapt_delta, max_treedepth=10, metric="dense_e" if dense else "diag_e"),
                       seed=1, pooled_adaptation=True, init_opt_steps=opt_steps, init_opt_subsample=subsample)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ess = fit.ess_bulk()
    flat = fit.draws.reshape(-1, fit.draws.shape[2])
    dev = torch.device("cuda", local_rank)
    # [seconds (max), ESS per parameter (sum over ranks: chains are independent), leapfrogs, divergent, sum of draws, sum of squares, count]
    v_max = torch.tensor([dt], dtype=torch.float64, device=dev)
    v_sum = torch.tensor(np.concatenate([ess, [fit.info["leapfrogs"], float(fit.divergent__.sum()), flat.shape[0]], flat.sum(0), (flat * flat).sum(0)]),
                         dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(v_max, op=dist.ReduceOp.MAX)
        dist.all_reduce(v_sum)
    dt = float(v_max.item())
    vs = v_sum.cpu().numpy()
    dim = len(ess)
    ess_all, leap, ndiv, cnt = vs[:dim], vs[dim], vs[dim + 1], vs[dim + 2]
    mean = vs[dim + 3:2 * dim + 3] / cnt
    sd = np.sqrt(np.maximum(vs[2 * dim + 3:3 * dim + 3] / cnt - mean * mean, 0.0))
    truth = np.asarray(cfg["theta_true"], dtype=float)
    z = (mean[:P] - truth) / sd[:P]
    d.close()
    try:
        have_rstan = subprocess.run(["Rscript", "-e", "library(rstan)"], capture_output=True, timeout=60).returncode == 0
    except Exception:
        have_rstan = False
    desc = "%s: %s; NUTS, %d iterations (%d warm-up), adapt_delta %.2f, starting values uniform(0, 1) followed by at most %d L-BFGS evaluations per chain" % (
        workload, workload_description(workload, chains, int(cfg["data"]["ndatapts"])), iters, warm, adapt_delta, opt_steps)
    if subsample:
        desc += " on a subsample of %d observations and as many on all of them" % subsample
    desc += "; %s metric: Laplace covariance at the optimum first, then the %s of the warm-up draws of all chains of all ranks at every window " \
            "(NCCL all-reduce of %d doubles), pooled step size" % ("dense (dense_e)" if dense else "diagonal (diag_e)", "covariance" if dense else "variances",
                                                                    2 + (eng.dim + eng.dim * (eng.dim + 1) // 2 if dense else 2 * eng.dim))
    return {"workload": desc,
            "treedepth_mean": float(fit.treedepth__.mean()), "n_leapfrog_mean": float(fit.n_leapfrog__.mean()), "stepsize": float(fit.stepsize__[0, -1]),
            "chains_total": int(chains * world), "seconds": dt, "min_ess_per_s": float(ess_all.min() / dt), "min_ess": float(ess_all.min()),
            "leapfrogs_per_s": leap / dt, "leapfrogs": int(leap), "passes": fit.info["passes"], "max_rhat": float(np.nanmax(fit.rhat)),
            "divergent": int(ndiv), "pool_exchanges": fit.info["pool_exchanges"], "collective": "torch.distributed all_reduce (NCCL)" if world > 1 else "none (one rank)",
            "posterior_mean": [float(x) for x in mean], "posterior_sd": [float(x) for x in sd], "max_abs_z_vs_truth": float(np.abs(z).max()),
            "ess_note": "Geyer ESS (rstan 2.19's n_eff) summed over ranks; R-hat = split R-hat over all chains",
            "rstan": "installed" if have_rstan else "rstan: not installed (no baseline ESS/s can be measured in this image)"}


def time_workload(bp, syn, _ffi, torch, dist, world, local_rank, rank, name, chains, steps, peaks, flush):
    """one more shape of SURVEY.md §8(d), device-timed like the main workload (events on the launching stream, L2 flush between
    steps): returns the entry of `other_workloads`."""
    dev = torch.device("cuda", local_rank)
    cfg = syn.make_config(name, seed=1, chains=chains)
    rng = np.random.default_rng(1000 + rank)
    upars = cfg["u_true"][None, :] + 0.1 * rng.standard_normal((chains, len(cfg["u_true"])))
    N = int(cfg["data"]["ndatapts"])
    mod = bp.bp_model(cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"])
    eng = bp.stan_model(mod, cfg["priors"], simple_bd=cfg["kind"] == "simple_bd", noise_model=cfg["kind"] == "multitype_noise")
    data = eng.data(cfg["data"], device=local_rank)
    L = _ffi.lib()
    dim = eng.dim
    u_dev = torch.from_numpy(upars).to(dev)
    lp_dev = torch.empty(chains, dtype=torch.float64, device=dev)
    g_dev = torch.empty(chains, dim, dtype=torch.float64, device=dev)
    st_dev = torch.empty(chains, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()

    def step(ev=None):
        flush.zero_()
        if ev is not None:
            ev[0].record(stream)
        _ffi.check(L.bpc_log_prob_dev(eng.handle, data.handle, u_dev.data_ptr(), chains, 1, lp_dev.data_ptr(), g_dev.data_ptr(),
                                      st_dev.data_ptr(), stream.cuda_stream))
        if ev is not None:
            ev[1].record(stream)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    flops = float(eng.flops_per_chain(data, upars).sum())
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for ev in evs:
        step(ev)
    torch.cuda.synchronize()
    ms = sum(a_.elapsed_time(b_) for a_, b_ in evs)
    bad = int((st_dev != 0).sum().item())
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms.item())
    kernel_name, ntiles, _ = data.path(chains)
    tensor = "bp_fusedo" in kernel_name or "bp_fusedc" in kernel_name or "bp_fusedw" in kernel_name
    peak = peaks["dmma"] if tensor else peaks["dfma"]
    per_step = ms / steps
    out = {"workload": name + ": " + workload_description(name, chains, N), "kernel": kernel_name, "ms_per_step": per_step,
           "value": float(world) * chains * N / (per_step * 1e-3), "unit": UNIT, "us_per_evaluation_of_all_chains": per_step * 1e3,
           "rejected_chains": bad,
           "roofline": {"bound": "fp64", "achieved": flops / (per_step * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                        "frac": flops / (per_step * 1e-3) / 1e12 / peak, "peak_kind": "DMMA loop" if tensor else "DFMA loop",
                        "flops_per_eval": flops / (chains * N),
                        "note": "whole step (all launches of one evaluation) against the measured FP64 peak; shapes with a few hundred "
                                "observations per chain are launch / latency bound, which is what the small fraction says"}}
    data.close()
    return out


def parity_report(eng, cfg, data, upars, nchains=8):
    """max relative error of the device's lp / gradient against the CPU port (oracle/) for the first chains of rank 0 on ALL
    observations of the workload, outside the timed region."""
    from oracle import bp_oracle as bo
    om = bo.OracleModel(cfg["kind"], cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"], cfg["priors"])
    u = np.ascontiguousarray(upars[:nchains])
    lp, g, st = eng.log_prob_grad(data, u)
    lp_o, g_o, st_o = bo.log_prob(om, cfg["data"], u, grouped=False if cfg["name"] in ("cfg5U", "cfg5K") else None, nthreads=host_cores())
    return {"chains": int(len(u)), "observations": int(cfg["data"]["ndatapts"]), "status_equal": bool((st == st_o).all()),
            "lp_max_rel_err": float(np.max(np.abs(lp - lp_o) / np.abs(lp_o))),
            "grad_max_rel_err": float(np.max(np.abs(g - g_o) / np.abs(g_o).max(axis=1, keepdims=True))),
            "checker": "oracle/ (CPU port, test infrastructure); tolerance of the tests: 1e-11 / 1e-9"}


def run_reference(args, rank):
    if rank != 0:
        return
    from bpinference_b200 import synthetic as syn
    chains = args.chains_per_gpu or DEFAULT_CHAINS[args.workload]
    cfg = syn.make_config(args.workload, chains=min(chains, 64), nobs=args.nobs or None)
    N = int(cfg["data"]["ndatapts"])
    per_step = max(2.0, args.cpu_seconds / max(1, args.steps + args.warmup))
    rates, info = [], None
    for i in range(args.warmup + args.steps):
        r, cores, sample = cpu_port_rate(cfg, per_step)
        if i >= args.warmup:
            rates.append(r)
            info = (cores, sample)
    v = float(np.mean(rates))
    _RESULT_LINE.append(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(args.workload, chains, N),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": info[0], "kind": "port", "sample": info[1] + " per step",
                         "note": "the CPU rate is linear in chains and does not depend on the GPU count: one host, all its cores"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    args = parse()
    # stdout carries exactly ONE line, the JSON; whatever libraries write to fd 1 meanwhile (e.g. "NCCL version ..." under NCCL_DEBUG=VERSION)
    # goes to stderr: fd 1 points at fd 2 until the line is printed
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        _main(args)
    finally:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    if _RESULT_LINE:
        print(_RESULT_LINE[0], flush=True)


_RESULT_LINE = []


def _main(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    import torch
    import torch.distributed as dist
    import bpinference_b200 as bp
    from bpinference_b200 import _ffi, synthetic as syn

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libbpcuda has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    chains = args.chains_per_gpu or DEFAULT_CHAINS[args.workload]
    # same observations on every rank (replicated data), different chains per rank
    cfg = syn.make_config(args.workload, seed=1, chains=chains, nobs=args.nobs or None)
    rng = np.random.default_rng(1000 + rank)
    upars = cfg["u_true"][None, :] + 0.1 * rng.standard_normal((chains, len(cfg["u_true"])))
    N = int(cfg["data"]["ndatapts"])
    mod = bp.bp_model(cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"])
    eng = bp.stan_model(mod, cfg["priors"], simple_bd=cfg["kind"] == "simple_bd", noise_model=cfg["kind"] == "multitype_noise")
    data = eng.data(cfg["data"], device=local_rank)
    L = _ffi.lib()
    dim = eng.dim
    u_dev = torch.from_numpy(upars).to(dev)
    lp_dev = torch.empty(chains, dtype=torch.float64, device=dev)
    g_dev = torch.empty(chains, dim, dtype=torch.float64, device=dev)
    st_dev = torch.empty(chains, dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    stream = torch.cuda.Stream(device=dev)      # the stream every kernel of the timed region is launched on
    torch.cuda.set_stream(stream)

    def step_dev(ev=None):
        flush.zero_()                                                    # L2 flush between timed iterations (not part of a step)
        if ev is not None:
            ev[0].record(stream)
        _ffi.check(L.bpc_log_prob_dev(eng.handle, data.handle, u_dev.data_ptr(), chains, 1, lp_dev.data_ptr(), g_dev.data_ptr(),
                                      st_dev.data_ptr(), stream.cuda_stream))
        if ev is not None:
            ev[1].record(stream)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- kernel-resident throughput (`value`) -------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step_dev()
    sync_all()
    flops_chain = eng.flops_per_chain(data, upars)        # exact algorithmic flops of every chain of this rank (one extra pass)
    L.bpc_launch_count(1)
    _ffi.check(L.bpc_kernel_timing(data.handle, 1))
    clocks = ClockSampler(local_rank)
    clocks.start()
    sync_all()
    # every step between its own pair of events on the launching stream; the L2 flush runs between the steps, outside the pairs
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for ev in evs:
        step_dev(ev)
    sync_all()
    ms = sum(a_.elapsed_time(b_) for a_, b_ in evs)
    clk = clocks.stop()
    launches = int(L.bpc_launch_count(0))
    kms, kl = _ffi.C.c_double(), _ffi.C.c_longlong()
    _ffi.check(L.bpc_kernel_time_ms(data.handle, 1, _ffi.C.byref(kms), _ffi.C.byref(kl)))
    _ffi.check(L.bpc_kernel_timing(data.handle, 0))
    bad = int((st_dev != 0).sum().item())
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())

    # ---- end to end through the public API with host buffers (`e2e`) --------------------------
    u_host = torch.from_numpy(upars.copy()).pin_memory().numpy()
    for _ in range(2):
        eng.log_prob_grad(data, u_host)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        lp_h, g_h, st_h = eng.log_prob_grad(data, u_host)                # H2D of upars, the step's launches, D2H of lp/grad/status, sync
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t_e2e = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_s = float(t_e2e.item())

    # ---- measured FP64 peaks of this device (same run): DFMA loop and DMMA loop -----------------
    peak_tf, clock_mhz = bp.fp64_peak(local_rank, 300.0)
    peak_dmma = bp.fp64_tensor_peak(local_rank, 300.0)
    peaks = {"dfma": peak_tf, "dmma": peak_dmma}

    # ---- the other shapes of SURVEY.md §8(d), every rank its own chains (weak scaling, no collective) ----------------------
    def side(fn, *a, **k):
        """the side blocks of the line (other shapes, sampling) must not cost the headline: on ONE rank a failure becomes an `error` entry.
        With several ranks the call stays unguarded -- a rank that swallowed an exception would leave the others waiting in a collective."""
        if world > 1:
            return fn(*a, **k)
        try:
            return fn(*a, **k)
        except Exception as e:      # noqa: BLE001 (reported in the JSON line)
            return {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}

    others = []
    if not args.no_other_workloads and args.workload == "cfg5U" and not args.nobs:
        for name in ("cfg5G", "cfg5K", "cfg2", "cfg1", "cfg3", "cfg4"):
            others.append(side(time_workload, bp, syn, _ffi, torch, dist, world, local_rank, rank, name, DEFAULT_CHAINS[name], 5 if name == "cfg5K" else 10,
                               peaks, flush))

    # ---- sampling: every rank its chains, warm-up statistics pooled over all ranks ----------------------------------------
    sampling = None
    if not args.no_sampling:
        # cfg2 (250 observations at ONE duration: a thin curved posterior no linear metric straightens, trees of depth 7-8 whatever the metric):
        # pooled diagonal metric, the best of tools/cfg2_variants.py; cfg5U (100 000 observations: a near-Gaussian posterior with a strongly
        # rotated ridge): pooled dense metric
        sampling = {"cfg2": side(sampling_report, bp, syn, torch, dist, rank, world, local_rank, "cfg2", 1024, None, 300, 150, 200, False)}
        if args.workload == "cfg5U" and not args.nobs:
            sampling["cfg5U"] = side(sampling_report, bp, syn, torch, dist, rank, world, local_rank, "cfg5U", chains, None, args.sampling_iter,
                                     args.sampling_iter // 2, 300, True, subsample=500)

    if rank == 0:
        total_evals = float(world) * chains * N * args.steps
        value = total_evals / (ms_max * 1e-3)
        k_avg_ms = kms.value / max(1, kl.value)
        flops_launch = float(flops_chain.sum())
        achieved = flops_launch / (k_avg_ms * 1e-3) / 1e12
        try:
            peaks_file = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            peaks_file = {}
        hbm_peak = peaks_file.get("hbm_gbs", 6650.0)
        kernel_name, ntiles, scratch_bytes = data.path(chains)
        # algorithmic HBM bytes of one launch: the observation table once + per-chain input / output.  The
        # kernels' own scratch (per-CTA partial sums, the octet path's S' / N_0 / W) is NOT algorithmic: it is reported as `scratch_bytes`
        # and shows up in `traffic`.
        obs_bytes = N * 8 * ((2 * cfg["data"].get("ntypes", 1)) + 1 + (cfg["ndep"] or 0))
        alg_bytes = obs_bytes + chains * (2 * dim + 1) * 8 + chains * 4
        traffic = NCU_TRAFFIC_BYTES.get((args.workload, chains, N))
        tensor = "bp_fusedo" in kernel_name or "bp_fusedc" in kernel_name or "bp_fusedw" in kernel_name
        peak = peak_dmma if tensor else peak_tf
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": bench_config(args.workload, chains, N),
            "rejected_chains": bad,
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/)",
                         "kernel": kernel_name, "ctas_per_chain": ntiles,
                         "kernel_ms": k_avg_ms, "flops_per_launch": flops_launch, "flops_per_eval": flops_launch / (chains * N),
                         "peak_dmma": peak_dmma, "peak_dfma": peak_tf, "frac_of_dfma_peak": achieved / peak_tf,
                         "peak_source": "measured in this run on this device: a register-resident DMMA (mma.sync m8n8k4 f64) loop — the kernel's flops "
                                        "are ~85 % DMMA — and a DFMA loop (bpc_fp64_tensor_peak / bpc_fp64_peak); MEASURED_PEAKS.json has no FP64 "
                                        "figure; nominal 148 SM x 64 lanes x 2 x 1.965 GHz = 37.2; DMMA and DFMA share one datapath",
                         "hbm": {"algorithmic_bytes": alg_bytes, "algorithmic_gbs": alg_bytes / (k_avg_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                                 "frac": alg_bytes / (k_avg_ms * 1e-3) / 1e9 / hbm_peak, "scratch_bytes": scratch_bytes,
                                 "traffic_over_algorithmic": (traffic / alg_bytes) if traffic else None,
                                 "peak_source": "MEASURED_PEAKS.json" if peaks_file else "fallback"}},
            "e2e": {"value": float(world) * chains * N * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": chains * dim * 8,
                    "d2h_bytes_per_step": chains * (dim + 1) * 8 + chains * 4,
                    "note": "bpc_log_prob with host buffers: upars H2D, lp/grad/status D2H every step; observations stay resident in the "
                            "data handle, as the data stays inside the stanfit in rstan::log_prob"},
            "gpu_launches": launches, "clocks": clk,
        }
        if not args.no_parity:
            out["parity"] = parity_report(eng, cfg, data, upars)
        if others:
            out["other_workloads"] = others
        if sampling:
            out["sampling"] = sampling
        if not args.no_cpu_baseline:
            v, cores, sample = cpu_port_rate(cfg, args.cpu_seconds)
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                                   "rstan": "not installed (no R in this image): the CPU restatement in oracle/ stands in; it does less work "
                                            "than rstan's RK45 forward-sensitivity solve"}
        _RESULT_LINE.append(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
