"""Sampling front end and fit object: the stand-in for `rstan::sampling` and the `stanfit` it returns, plus the
reference's post-hoc diagnostics check_div / check_exceeded_treedepth / check_rhat (R/stan_utils.R:493-533).

Multi-GPU: one process per GPU (torchrun); chains are sharded over ranks; torch.distributed (NCCL) carries the two
small exchanges of the path: the pooled warm-up statistics (optional) and the chain moments behind R-hat.  (The same
exchanges inside ONE process, one host thread per GPU over ncclCommInitAll, are bpc_sample_multi / `sampling_multi` below.)
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from . import _ffi


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist
    except Exception:
        pass
    return None


class BpFit:
    """what rstan::sampling returns, reduced to what the reference's scripts read: draws of Theta1..P (+ sigma_obs), the
    transformed parameters r_mat and theta (multitype_template.stan:105-117), lp__, and the sampler parameters divergent__,
    treedepth__, n_leapfrog__, accept_stat__, stepsize__.

    Column order of `as_matrix()` / `extract()` / `summary()` is rstan's: parameters in declaration order, transformed parameters
    (r_mat column-major, then theta with the first index fastest), generated quantities (log_lik, when sampled with loo=True), lp__ —
    so scripts that index columns positionally (vignettes/two-type-inference.Rmd:63-69 reads `samples[,1:6]`) keep working."""

    def __init__(self, names, draws, lp, divergent, treedepth, n_leapfrog, accept_stat, stepsize, rhat, info, tpars=None, log_lik=None):
        self.names = list(names)
        self.draws = draws                   # [chains, iter - warmup, dim] constrained
        self.lp__ = lp
        self.divergent__, self.treedepth__, self.n_leapfrog__ = divergent, treedepth, n_leapfrog
        self.accept_stat__, self.stepsize__ = accept_stat, stepsize
        self.rhat = rhat                     # split R-hat over ALL chains of all ranks (parameters only)
        self.info = info
        self._tpars = tpars                  # callable -> (names, values [chains, ns, ncols]); evaluated on first use
        self._tp_cache = None
        self.log_lik = log_lik               # [chains, ns, ndatapts] or None

    # -- transformed parameters (computed by the device's rate functions on first use) -----------
    def transformed_parameters(self):
        """(names, values [chains, iter - warmup, ncols]) of r_mat and theta."""
        if self._tp_cache is None:
            self._tp_cache = self._tpars() if self._tpars is not None else ([], np.empty(self.draws.shape[:2] + (0,)))
        return self._tp_cache

    @property
    def colnames(self):
        tn, _ = self.transformed_parameters()
        ll = [] if self.log_lik is None else ["log_lik[%d]" % (k + 1) for k in range(self.log_lik.shape[2])]
        return self.names + list(tn) + ll + ["lp__"]

    def as_matrix(self):
        """draws x columns (chains concatenated, as rstan's as.matrix(fit)), columns = `colnames`."""
        _, tv = self.transformed_parameters()
        parts = [self.draws, tv] + ([] if self.log_lik is None else [self.log_lik]) + [self.lp__[:, :, None]]
        return np.concatenate(parts, axis=2).reshape(-1, sum(p.shape[2] for p in parts))

    def get_sampler_params(self, inc_warmup: bool = False):
        """list with one dict per chain (rstan::get_sampler_params(fit, inc_warmup = FALSE))."""
        return [dict(accept_stat__=self.accept_stat__[c], stepsize__=self.stepsize__[c], treedepth__=self.treedepth__[c],
                     n_leapfrog__=self.n_leapfrog__[c], divergent__=self.divergent__[c]) for c in range(self.draws.shape[0])]

    def extract(self, pars=None):
        """dict name -> draws (chains concatenated).  Scalars Theta*, sigma_obs, lp__; `r_mat` [draws, nevents, ntypes]
        (one-type template: [draws, 2, ndep_levels]) and `theta` [draws, ndep_levels, 2*nevents*ntypes] as rstan::extract shapes them."""
        flat = self.draws.reshape(-1, self.draws.shape[2])
        out = {nm: flat[:, k] for k, nm in enumerate(self.names)}
        tn, tv = self.transformed_parameters()
        if len(tn):
            tvf = tv.reshape(-1, tv.shape[2])
            nr = sum(1 for nm in tn if nm.startswith("r_mat["))
            rows = max(int(nm[6:-1].split(",")[0]) for nm in tn[:nr])
            out["r_mat"] = tvf[:, :nr].reshape(-1, nr // rows, rows).transpose(0, 2, 1)          # column-major columns -> [e, j]
            if len(tn) > nr:
                lev = max(int(nm[6:-1].split(",")[0]) for nm in tn[nr:])
                out["theta"] = tvf[:, nr:].reshape(-1, (len(tn) - nr) // lev, lev).transpose(0, 2, 1)
        if self.log_lik is not None:
            out["log_lik"] = self.log_lik.reshape(-1, self.log_lik.shape[2])
        out["lp__"] = self.lp__.reshape(-1)
        return out if pars is None else {k: out[k] for k in ([pars] if isinstance(pars, str) else pars)}

    @staticmethod
    def _ess(x):
        """effective sample size of draws x [chains, n]: Geyer's initial positive sequence on the autocorrelation averaged over
        chains, as rstan 2.19's summary computes n_eff (no rank normalisation)."""
        Cn, n = x.shape
        xc = x - x.mean(axis=1, keepdims=True)
        nfft = 1 << int(np.ceil(np.log2(2 * n)))
        f = np.fft.rfft(xc, nfft, axis=1)
        acov = np.fft.irfft(f * np.conj(f), nfft, axis=1)[:, :n] / n
        W = acov[:, 0].mean() * n / (n - 1)
        if not W > 0:
            return float("nan")
        B_n = x.mean(axis=1).var(ddof=1) if Cn > 1 else 0.0
        var_plus = W * (n - 1) / n + B_n
        rho = 1.0 - (W - acov.mean(axis=0) * n / (n - 1)) / var_plus
        s, t = 0.0, 0
        while t + 1 < n:
            pair = rho[t] + rho[t + 1]
            if pair < 0:
                break
            s += pair
            t += 2
        tau = -1.0 + 2.0 * s
        return Cn * n / max(tau, 1.0 / np.log10(max(Cn * n, 10)))

    def ess_bulk(self):
        """effective sample size per parameter over THIS rank's chains (see _ess)."""
        return np.array([self._ess(self.draws[:, :, k]) for k in range(self.draws.shape[2])])

    @staticmethod
    def _split_rhat(x):
        Cn, n = x.shape
        if n >= 4:
            h = n // 2
            x = np.concatenate([x[:, :h], x[:, n - h:]], axis=0)
            n = h
        W = x.var(axis=1, ddof=1).mean()
        if not W > 0:
            return float("nan")
        return float(np.sqrt((W * (n - 1) / n + x.mean(axis=1).var(ddof=1)) / W))

    def summary(self, probs=(0.025, 0.25, 0.5, 0.75, 0.975), max_columns: int = 4096):
        """rows in `colnames` order; columns mean, sd, quantiles, n_eff, Rhat (rstan::summary(fit)$summary).  Parameters carry the
        all-rank split R-hat; transformed parameters and lp__ the statistics of this rank's chains.  Models whose transformed
        parameters run to more than `max_columns` columns (one theta row per dependent-variable level) list the parameters and lp__ only."""
        def row(x, rhat=None):
            f = x.reshape(-1)
            return dict(mean=f.mean(), sd=f.std(ddof=1), **{"%g%%" % (100 * p): np.quantile(f, p) for p in probs}, n_eff=self._ess(x),
                        Rhat=self._split_rhat(x) if rhat is None else rhat)
        rows = {nm: row(self.draws[:, :, k], self.rhat[k]) for k, nm in enumerate(self.names)}
        tn, tv = self.transformed_parameters()
        if len(tn) <= max_columns:
            for k, nm in enumerate(tn):
                rows[nm] = row(tv[:, :, k])
        rows["lp__"] = row(self.lp__)
        return rows


def check_div(fit) -> int:
    """R/stan_utils.R:493-498 — number of post-warm-up transitions that ended with a divergence."""
    return int(np.sum(np.vstack([p["divergent__"] for p in fit.get_sampler_params(inc_warmup=False)])))


def check_exceeded_treedepth(fit, max_depth: int = 10) -> str:
    """R/stan_utils.R:502-512."""
    td = np.concatenate([p["treedepth__"] for p in fit.get_sampler_params(inc_warmup=False)])
    n, N = int(np.sum(td == max_depth)), int(td.size)
    return "%s of %s iterations saturated the maximum tree depth of %s (%s%%)" % (n, N, max_depth, 100 * n / N)


def check_rhat(fit) -> str:
    """R/stan_utils.R:516-533."""
    ok = True
    for nm, r in zip(fit.names, fit.rhat):
        if (r > 1.1 or np.isinf(r)) and not np.isnan(r):
            print("Rhat for parameter %s is %s!" % (nm, r))
            ok = False
    return "Rhat looks reasonable for all parameters" if ok else "Rhat above 1.1 indicates that the chains very likely have not mixed"


def pooled_loop(run, exchange, apply, world: int) -> int:
    """the pooled warm-up protocol of a multi-rank run (the only collective of the sampling path), written so that it cannot deadlock:
    EVERY rank enters the exchange every time its own chains stop — at a window boundary (run() == 2) or because they have all finished
    or failed (run() == 1).  A rank that is done keeps contributing zeros and a raised "finished" flag until the flag sum says all
    `world` ranks are done, so no rank ever waits in a collective the others have left (a rank whose chains all failed
    initialisation used to skip the all-reduce).  run() -> state; exchange() -> sum over ranks of the finished flags, leaving the reduced
    statistics where apply() reads them.  Returns the number of exchanges.  bpc_sample_multi (csrc/multi.cpp) runs the same protocol."""
    n = 0
    while True:
        state = run()
        done = exchange()
        n += 1
        if done >= world:
            return n
        if state == 2:
            apply()


def _metric_code(control) -> int:
    m = control.get("metric", "diag_e")
    if m not in ("diag_e", "dense_e"):
        raise ValueError("control$metric must be \"diag_e\" or \"dense_e\"")
    return 1 if m == "dense_e" else 0


def subsample_data(dat: dict, n: int, seed: int, simple_bd: bool) -> dict:
    """a Stan data list with n of the observations of `dat`, drawn without replacement (the same draw on every rank: the seed is the run's)."""
    N = int(dat["ndatapts"])
    idx = np.sort(np.random.default_rng(10_000 + int(seed)).choice(N, size=n, replace=False))
    s = dict(dat)
    s["ndatapts"] = n
    for k in ("pop_vec", "init_pop"):
        a = np.asarray(dat[k])
        s[k] = a[idx] if a.ndim == 1 else a.reshape(N, -1)[idx]
    s["var_idx"] = np.asarray(dat["var_idx"])[idx]
    if simple_bd:
        s["times"] = np.asarray(dat["times"])[idx]
    else:
        s["times_idx"] = np.asarray(dat["times_idx"])[idx]
    return s


def _names(engine):
    return ["Theta%d" % (k + 1) for k in range(engine.nparams)] + (["sigma_obs"] if engine.kind == _ffi.MULTITYPE_NOISE else [])


def _tpars_fn(engine, d, draws):
    """closure computing the r_mat / theta columns of `draws` [chains, ns, dim] with the device's rate functions."""
    def run():
        L = _ffi.lib()
        Cn, ns, dim = draws.shape
        nlev = int(L.bpc_data_nlevels(d.handle))
        E, n = engine.nevents, engine.ntypes
        flat = np.ascontiguousarray(draws.reshape(-1, dim))
        rates = np.empty((flat.shape[0], nlev, E))
        _ffi.check(L.bpc_transformed_parameters(engine.handle, d.handle, _ffi.dptr(flat), flat.shape[0], _ffi.dptr(rates)))
        if engine.kind == _ffi.SIMPLE_BD:      # matrix[2, ndep_levels] r_mat, filled for every level (one_type_template.stan:19-27)
            names = ["r_mat[%d,%d]" % (e + 1, v + 1) for v in range(nlev) for e in range(2)]
            return names, rates.reshape(Cn, ns, nlev * 2)
        pz = np.asarray(engine.model["p_vec"], dtype=int) - 1
        emat = np.asarray(engine.model["e_mat"], dtype=float).reshape(E, n)
        # r_mat[e, p_vec[e]] = rate, zero elsewhere; rstan keeps the matrix of the LAST level (it is reset at the top of the loop)
        rm = np.zeros((flat.shape[0], nlev, n, E))                      # [draw, level, column j, row e]: column-major flattening
        for e in range(E):
            rm[:, :, pz[e], e] = rates[:, :, e]
        names = ["r_mat[%d,%d]" % (e + 1, j + 1) for j in range(n) for e in range(E)]
        cols = [rm[:, nlev - 1].reshape(flat.shape[0], n * E)]
        th = np.concatenate([np.broadcast_to(emat.ravel(order="F")[None, None, :], (flat.shape[0], nlev, E * n)),
                             rm.reshape(flat.shape[0], nlev, n * E)], axis=2)     # theta[i] = append_array(to_array_1d(e_mat), to_array_1d(r_mat))
        names += ["theta[%d,%d]" % (v + 1, q + 1) for q in range(2 * E * n) for v in range(nlev)]
        cols.append(th.transpose(0, 2, 1).reshape(flat.shape[0], 2 * E * n * nlev))
        vals = np.concatenate(cols, axis=1)
        return names, vals.reshape(Cn, ns, vals.shape[1])
    return run


def _default_device():
    """the GPU of this process: LOCAL_RANK under torchrun, else torch's current device when torch is loaded, else 0."""
    if "LOCAL_RANK" in os.environ:
        return int(os.environ["LOCAL_RANK"])
    try:
        import sys
        if "torch" in sys.modules:
            import torch
            if torch.cuda.is_available():
                return int(torch.cuda.current_device())
    except Exception:
        pass
    return 0


def sampling(engine, dat, chains: int = 4, iter: int = 2000, warmup: Optional[int] = None, init="random", control: Optional[dict] = None,
             seed: int = 1, pooled_adaptation: bool = False, device: Optional[int] = None, loo: bool = False, init_opt_steps: int = 0,
             init_opt_subsample: int = 0):
    # control$metric: "diag_e" (rstan's default) or "dense_e" (needs pooled_adaptation: the covariance comes from all chains)
    """rstan::sampling(stan_mod, data = dat, chains, iter, warmup, init, control = list(adapt_delta, max_treedepth)).

    `chains` is the number of chains of THIS process; under torchrun every rank runs `chains` chains with global
    chain ids rank*chains + c.  `init` is "random" or a list of dicts {"Theta1": ..} as uniform_initialize returns.
    loo: also return log_lik [chains, draws, ndatapts] (generate(..., loo = T), R/stan_utils.R:173-175,189-191).
    init_opt_steps > 0: every chain first climbs from its initial point with that many steps of a device-side optimiser (see
    bpc_sample_args.init_opt_steps) — an extension for data sets whose posterior is orders of magnitude narrower than the range the
    initial values are drawn from; 0 is rstan's behaviour.
    init_opt_subsample = n > 0 (with init_opt_steps and pooled_adaptation, `dat` a data list): the ascent first runs on n observations
    drawn at random from the list (far from the mode every evaluation costs many times one near it), then on all of them."""
    L = _ffi.lib()
    control = control or {}
    warmup = iter // 2 if warmup is None else warmup
    dist = _dist()
    rank = dist.get_rank() if dist else 0
    world = dist.get_world_size() if dist else 1
    own_data = not hasattr(dat, "handle")
    d = engine.data(dat, device=_default_device() if device is None else device) if own_data else dat
    inits = None
    if not isinstance(init, str):
        if len(init) != chains:
            raise ValueError("init must have one entry per chain")
        inits = np.ascontiguousarray([[float(ch["Theta%d" % (k + 1)]) for k in range(engine.nparams)] for ch in init], dtype=np.float64)
    d_sub = None
    if init_opt_subsample and init_opt_steps > 0 and pooled_adaptation and 0 < init_opt_subsample < d.ndatapts:
        if not isinstance(d.source, dict):
            raise ValueError("init_opt_subsample needs the Stan data list the handle was made from")
        d_sub = engine.data(subsample_data(d.source, int(init_opt_subsample), seed, engine.kind == _ffi.SIMPLE_BD), device=d.device)
    args = _ffi.SampleArgs(chains, iter, warmup, float(control.get("adapt_delta", 0.8)), int(control.get("max_treedepth", 10)), seed,
                           rank * chains, _ffi.dptr(inits) if inits is not None else None, int(bool(pooled_adaptation)),
                           float(control.get("stepsize", 0.0)), int(init_opt_steps), _metric_code(control),
                           d_sub.handle if d_sub is not None else None)
    h = C.c_void_p()
    try:
        _ffi.check(L.bpc_sampler_create(engine.handle, d.handle, C.byref(args), C.byref(h)))
        state = C.c_int(0)
        dim = engine.dim
        npool = int(L.bpc_sampler_pool_len(h))
        pool = mom = None
        if dist:
            import torch
            dev = torch.device("cuda", d.device)
            pool = torch.zeros(npool, dtype=torch.float64, device=dev)
            mom = torch.zeros(4 * dim, dtype=torch.float64, device=dev)
        def run():
            _ffi.check(L.bpc_sampler_run(h, 1 << 30, C.byref(state)))      # returns at 1 (all chains finished) or 2 (pooled window)
            return state.value

        if pooled_adaptation and dist:
            def exchange():
                _ffi.check(L.bpc_sampler_pool_stats(h, pool.data_ptr(), pool.numel()))
                dist.all_reduce(pool)
                torch.cuda.synchronize()
                return float(pool[npool - 1].item())
            exchanges = pooled_loop(run, exchange, lambda: _ffi.check(L.bpc_sampler_pool_apply(h, pool.data_ptr())), world)
        else:
            exchanges = 0
            while run() != 1:
                _ffi.check(L.bpc_sampler_pool_stats(h, None, 0))
                _ffi.check(L.bpc_sampler_pool_apply(h, None))
                exchanges += 1
        ns = iter - warmup
        draws = np.empty((chains, ns, dim))
        lp, acc, eps = np.empty((chains, ns)), np.empty((chains, ns)), np.empty((chains, ns))
        div, td, nl = (np.empty((chains, ns), dtype=np.int32) for _ in range(3))
        _ffi.check(L.bpc_sampler_draws(h, _ffi.dptr(draws), _ffi.dptr(lp), _ffi.iptr(div), _ffi.iptr(td), _ffi.iptr(nl), _ffi.dptr(acc), _ffi.dptr(eps)))
        red = np.empty(4 * dim)
        if dist:
            _ffi.check(L.bpc_sampler_chain_moments(h, mom.data_ptr(), mom.numel(), None))
            dist.all_reduce(mom)
            red = mom.cpu().numpy()
        else:
            _ffi.check(L.bpc_sampler_chain_moments(h, None, 0, _ffi.dptr(red)))
        rhat = np.empty(dim)
        _ffi.check(L.bpc_rhat(_ffi.dptr(np.ascontiguousarray(red)), dim, ns, _ffi.dptr(rhat)))
        half = 2 if ns >= 4 else 1
        info = dict(passes=int(L.bpc_sampler_passes(h)), leapfrogs=int(L.bpc_sampler_total_leapfrogs(h)), chains=chains, iter=iter, warmup=warmup,
                    total_chains=int(round(red[0] / half)), pool_exchanges=exchanges, world=world)
        log_lik = None
        if loo:
            # (C x draws x N doubles, as rstan stores them; evaluated in batches of at most 2^26 values on the device)
            u = engine.unconstrain_pars(draws.reshape(-1, dim))
            step = max(1, (1 << 26) // max(1, d.ndatapts))
            log_lik = np.concatenate([engine.log_lik(d, u[i:i + step]) for i in range(0, len(u), step)], axis=0).reshape(chains, ns, -1)
        if own_data:
            # the transformed parameters need the data handle: evaluate them now, before it is closed
            tp = _tpars_fn(engine, d, draws)()
            tpars = (lambda tp=tp: tp)
        else:
            tpars = _tpars_fn(engine, d, draws)
    finally:
        if h:
            L.bpc_sampler_free(h)
        if d_sub is not None:
            d_sub.close()
        if own_data:
            d.close()
    return BpFit(_names(engine), draws, lp, div, td, nl, acc, eps, rhat, info, tpars=tpars, log_lik=log_lik)


def sampling_multi(engine, dat: dict, devices, chains: int = 4, iter: int = 2000, warmup: Optional[int] = None, init="random",
                   control: Optional[dict] = None, seed: int = 1, pooled_adaptation: bool = False, init_opt_steps: int = 0):
    """the same run sharded over several GPUs of THIS process (bpc_sample_multi): `chains` chains on each device of `devices`, one
    host thread per device inside libbpcuda, one ncclCommInitAll communicator for the pooled warm-up statistics and split R-hat.
    No torch, no torchrun: this is what the R shim calls.  `dat` is the Stan data list (it is replicated on every device)."""
    L = _ffi.lib()
    control = control or {}
    warmup = iter // 2 if warmup is None else warmup
    devices = np.ascontiguousarray(np.asarray(devices, dtype=np.int32))
    ndev, total = len(devices), len(devices) * chains
    sd, keep = engine.stan_data_struct(dat)
    inits = None
    if not isinstance(init, str):
        if len(init) != total:
            raise ValueError("init must have one entry per chain (devices x chains)")
        inits = np.ascontiguousarray([[float(ch["Theta%d" % (k + 1)]) for k in range(engine.nparams)] for ch in init], dtype=np.float64)
    args = _ffi.SampleArgs(chains, iter, warmup, float(control.get("adapt_delta", 0.8)), int(control.get("max_treedepth", 10)), seed,
                           0, _ffi.dptr(inits) if inits is not None else None, int(bool(pooled_adaptation)),
                           float(control.get("stepsize", 0.0)), int(init_opt_steps), _metric_code(control), None)
    ns, dim = iter - warmup, engine.dim
    draws = np.empty((total, ns, dim))
    lp, acc, eps = np.empty((total, ns)), np.empty((total, ns)), np.empty((total, ns))
    div, td, nl = (np.empty((total, ns), dtype=np.int32) for _ in range(3))
    rhat = np.full(dim, np.nan)
    info = (C.c_longlong * 4)()
    _ffi.check(L.bpc_sample_multi(engine.handle, C.byref(sd), _ffi.iptr(devices), ndev, C.byref(args), _ffi.dptr(draws), _ffi.dptr(lp),
                                  _ffi.iptr(div), _ffi.iptr(td), _ffi.iptr(nl), _ffi.dptr(acc), _ffi.dptr(eps), _ffi.dptr(rhat), info))
    del keep
    meta = dict(passes=int(info[0]), leapfrogs=int(info[1]), chains=chains, iter=iter, warmup=warmup, total_chains=total,
                pool_exchanges=int(info[2]), world=ndev, nccl_version=int(info[3]), devices=devices.tolist())

    def tpars():
        d = engine.data(dat, device=int(devices[0]))
        try:
            return _tpars_fn(engine, d, draws)()
        finally:
            d.close()
    return BpFit(_names(engine), draws, lp, div, td, nl, acc, eps, rhat, meta, tpars=tpars)
