"""Sampling front end and fit object: the stand-in for `rstan::sampling` and the `stanfit` it returns, plus the
reference's post-hoc diagnostics check_div / check_exceeded_treedepth / check_rhat (R/stan_utils.R:493-533).

Multi-GPU: one process per GPU (torchrun); chains are sharded over ranks; torch.distributed (NCCL) carries the two
small exchanges of the path: the pooled warm-up statistics (optional) and the chain moments behind R-hat.
"""
from __future__ import annotations

import ctypes as C
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
    """what rstan::sampling returns, reduced to what the reference's scripts read: draws of Theta1..P (+ sigma_obs),
    lp__, and the sampler parameters divergent__, treedepth__, n_leapfrog__, accept_stat__, stepsize__."""

    def __init__(self, names, draws, lp, divergent, treedepth, n_leapfrog, accept_stat, stepsize, rhat, info):
        self.names = list(names)
        self.draws = draws                   # [chains, iter - warmup, dim] constrained
        self.lp__ = lp
        self.divergent__, self.treedepth__, self.n_leapfrog__ = divergent, treedepth, n_leapfrog
        self.accept_stat__, self.stepsize__ = accept_stat, stepsize
        self.rhat = rhat                     # over ALL chains of all ranks
        self.info = info

    def get_sampler_params(self, inc_warmup: bool = False):
        """list with one dict per chain (rstan::get_sampler_params(fit, inc_warmup = FALSE))."""
        return [dict(accept_stat__=self.accept_stat__[c], stepsize__=self.stepsize__[c], treedepth__=self.treedepth__[c],
                     n_leapfrog__=self.n_leapfrog__[c], divergent__=self.divergent__[c]) for c in range(self.draws.shape[0])]

    def extract(self):
        flat = self.draws.reshape(-1, self.draws.shape[2])
        out = {nm: flat[:, k] for k, nm in enumerate(self.names)}
        out["lp__"] = self.lp__.reshape(-1)
        return out

    def ess_bulk(self):
        """effective sample size per parameter (Geyer initial-positive-sequence on the pooled autocorrelation of
        this rank's chains)."""
        Cn, n, dim = self.draws.shape
        out = np.empty(dim)
        for k in range(dim):
            x = self.draws[:, :, k]
            xc = x - x.mean(axis=1, keepdims=True)
            nfft = 1 << int(np.ceil(np.log2(2 * n)))
            f = np.fft.rfft(xc, nfft, axis=1)
            acov = np.fft.irfft(f * np.conj(f), nfft, axis=1)[:, :n] / n
            W = acov[:, 0].mean() * n / (n - 1)
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
            out[k] = Cn * n / max(tau, 1.0 / np.log10(max(Cn * n, 10)))
        return out

    def summary(self, probs=(0.025, 0.25, 0.5, 0.75, 0.975)):
        """rows = parameters; columns mean, sd, quantiles, n_eff, Rhat (rstan::summary(fit)$summary)."""
        flat = self.draws.reshape(-1, self.draws.shape[2])
        ess = self.ess_bulk()
        rows = {}
        for k, nm in enumerate(self.names):
            rows[nm] = dict(mean=flat[:, k].mean(), sd=flat[:, k].std(ddof=1), **{"%g%%" % (100 * p): np.quantile(flat[:, k], p) for p in probs},
                            n_eff=ess[k], Rhat=self.rhat[k])
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


def sampling(engine, dat, chains: int = 4, iter: int = 2000, warmup: Optional[int] = None, init="random", control: Optional[dict] = None,
             seed: int = 1, pooled_adaptation: bool = False, device: Optional[int] = None):
    """rstan::sampling(stan_mod, data = dat, chains, iter, warmup, init, control = list(adapt_delta, max_treedepth)).

    `chains` is the number of chains of THIS process; under torchrun every rank runs `chains` chains with global
    chain ids rank*chains + c.  `init` is "random" or a list of dicts {"Theta1": ..} as uniform_initialize returns."""
    L = _ffi.lib()
    control = control or {}
    warmup = iter // 2 if warmup is None else warmup
    dist = _dist()
    rank = dist.get_rank() if dist else 0
    d = dat if hasattr(dat, "handle") else engine.data(dat, device=0 if device is None else device)
    inits = None
    if not isinstance(init, str):
        if len(init) != chains:
            raise ValueError("init must have one entry per chain")
        inits = np.ascontiguousarray([[float(ch["Theta%d" % (k + 1)]) for k in range(engine.nparams)] for ch in init], dtype=np.float64)
    args = _ffi.SampleArgs(chains, iter, warmup, float(control.get("adapt_delta", 0.8)), int(control.get("max_treedepth", 10)), seed,
                           rank * chains, _ffi.dptr(inits) if inits is not None else None, int(bool(pooled_adaptation)),
                           float(control.get("stepsize", 0.0)))
    h = C.c_void_p()
    _ffi.check(L.bpc_sampler_create(engine.handle, d.handle, C.byref(args), C.byref(h)))
    try:
        state = C.c_int(0)
        dim = engine.dim
        pool = mom = None
        if dist:
            import torch
            dev = torch.device("cuda", d.device)
            pool = torch.zeros(1 + 2 * dim, dtype=torch.float64, device=dev)
            mom = torch.zeros(4 * dim, dtype=torch.float64, device=dev)
        while state.value != 1:
            _ffi.check(L.bpc_sampler_run(h, 1 << 30, C.byref(state)))
            if state.value == 2:   # pooled warm-up window: the only collective of the sampling path
                if dist:
                    _ffi.check(L.bpc_sampler_pool_stats(h, pool.data_ptr(), pool.numel()))
                    dist.all_reduce(pool)
                    torch.cuda.synchronize()
                    _ffi.check(L.bpc_sampler_pool_apply(h, pool.data_ptr()))
                else:
                    _ffi.check(L.bpc_sampler_pool_stats(h, None, 0))
                    _ffi.check(L.bpc_sampler_pool_apply(h, None))
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
        info = dict(passes=int(L.bpc_sampler_passes(h)), leapfrogs=int(L.bpc_sampler_total_leapfrogs(h)), chains=chains, iter=iter, warmup=warmup,
                    total_chains=int(red[0]))
    finally:
        L.bpc_sampler_free(h)
    names = ["Theta%d" % (k + 1) for k in range(engine.nparams)] + (["sigma_obs"] if engine.kind == _ffi.MULTITYPE_NOISE else [])
    return BpFit(names, draws, lp, div, td, nl, acc, eps, rhat, info)
