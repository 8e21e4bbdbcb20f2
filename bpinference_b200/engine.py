"""The engine boundary: what a user of the reference hands to rstan goes to libbpcuda instead.

    reference (examples/two_type.R:18-26)                   here
    ------------------------------------------------------  ---------------------------------------------
    stan_code <- generate(mod, priors, noise_model=, ...)   eng = stan_model(mod, priors, noise_model=, ...)
    stan_mod  <- rstan::stan_model(model_code = stan_code)
    fit <- rstan::sampling(stan_mod, data = dat, ...)       fit = eng.sampling(dat, chains=, iter=, warmup=, init=, control=)
    rstan::log_prob(fit, upars, adjust_transform, gradient) eng.log_prob(dat, upars, adjust_transform, gradient)
    rstan::grad_log_prob(fit, upars)                        eng.grad_log_prob(dat, upars)

`stan_model` takes the same (model, priors, simple_bd, noise_model) as generate() (R/stan_utils.R:130).
All arithmetic happens in libbpcuda's CUDA kernels; this module only marshals arrays.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _ffi
from .priors import prior_fields


def _f64(a, order="C"):
    return np.require(np.asarray(a, dtype=np.float64), requirements=["C" if order == "C" else "F", "A"])


def _stan_data_struct(engine: "BpCudaModel", dat: dict):
    """the Stan data list as the bpc_stan_data struct of include/bpcuda.h (+ the arrays it points into, to be kept alive)."""
    simple = engine.kind == _ffi.SIMPLE_BD
    N = int(dat["ndatapts"])
    n = engine.ntypes
    pop = np.asfortranarray(_f64(dat["pop_vec"]).reshape(N, n) if not simple else _f64(dat["pop_vec"]).reshape(N, 1))
    init = np.asfortranarray(_f64(dat["init_pop"]).reshape(N, n) if not simple else _f64(dat["init_pop"]).reshape(N, 1))
    times = _f64(np.atleast_1d(dat["times"])).ravel()
    fv = np.asfortranarray(_f64(dat["function_var"]).reshape(int(dat["ndep_levels"]), int(dat["ndep"])))
    vi = np.ascontiguousarray(np.asarray(dat["var_idx"], dtype=np.int32))
    ti = None if simple else np.ascontiguousarray(np.asarray(dat["times_idx"], dtype=np.int32))
    if simple and len(times) != N:
        raise ValueError("times, initial populations, and final populations must all have same number of rows!")
    sd = _ffi.StanData(N, 0 if simple else len(times), int(dat["ndep_levels"]), int(dat["ndep"]),
                       _ffi.dptr(pop), _ffi.dptr(init), _ffi.dptr(times), None if simple else _ffi.iptr(ti), _ffi.dptr(fv), _ffi.iptr(vi))
    return sd, (pop, init, times, fv, vi, ti)


class StanData:
    """device copy of one Stan data list (bpc_data handle)."""

    def __init__(self, engine: "BpCudaModel", dat: dict, device: int = 0, grouping: int = 0):
        self.engine = engine
        L = _ffi.lib()
        N = int(dat["ndatapts"])
        sd, self._keep = _stan_data_struct(engine, dat)
        self.handle = C.c_void_p()
        _ffi.check(L.bpc_data_create(engine.handle, C.byref(sd), int(device), int(grouping), C.byref(self.handle)))
        self.ndatapts = N
        self.device = device
        self.source = dat          # the Stan data list (fit.sampling draws the subsample of a two-stage ascent from it)

    @property
    def ngroups(self) -> int:
        return _ffi.lib().bpc_data_ngroups(self.handle)

    @property
    def grouped(self) -> bool:
        return bool(_ffi.lib().bpc_data_is_grouped(self.handle))

    def path(self, nchains: int):
        """(kernel path name, CTAs per chain, partial-sum bytes per launch) of bpc_log_prob for `nchains` chains."""
        p, nt, sb = C.c_int(), C.c_int(), C.c_ulonglong()
        _ffi.check(_ffi.lib().bpc_data_path(self.handle, int(nchains), C.byref(p), C.byref(nt), C.byref(sb)))
        names = {0: "bp_fused", 1: "bp_grouped", 2: "bp_fusedw + bp_wpost (bp_fused for chains that need sub-steps)", 3: "bp_onetype",
                 4: "bp_ctable + bp_fusedc + bp_wpost (bp_fused for chains that need sub-steps)",
                 5: "bp_otable + bp_fusedo + bp_toep + bp_wpost (bp_fused for chains that need sub-steps)"}
        return names[p.value], nt.value, sb.value

    def close(self):
        if getattr(self, "handle", None):
            _ffi.lib().bpc_data_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BpCudaModel:
    """a compiled model (bpc_model handle): rate expressions emitted as CUDA and compiled by NVRTC for sm_100a."""

    def __init__(self, model, priors: Sequence, simple_bd: bool = False, noise_model: bool = False):
        if getattr(model, "r_class", None) != "bp_model":
            raise ValueError("model must be a bp_model object!")
        if len(priors) != model["nparams"]:
            raise ValueError("incorrect number of priors for model!")
        e_mat = np.asfortranarray(_f64(model["e_mat"]))
        if simple_bd and not (e_mat.shape == (2, 1) and e_mat[0, 0] == 2 and e_mat[1, 0] == 0):
            raise ValueError("model is not a simple birth-death process!")
        if noise_model and not simple_bd and e_mat.shape[1] == 1:
            raise ValueError("observational noise models are not available for one-type processes")
        self.kind = _ffi.SIMPLE_BD if simple_bd else (_ffi.MULTITYPE_NOISE if noise_model else _ffi.MULTITYPE)
        self.ntypes, self.nevents = e_mat.shape[1], e_mat.shape[0]
        self.nparams, self.ndep = int(model["nparams"]), int(model["ndep"])
        p_vec = np.ascontiguousarray(np.asarray(model["p_vec"], dtype=np.int32))
        fds = [s.encode() for s in model["func_deps"]]
        fd_arr = (C.c_char_p * len(fds))(*fds)
        pr_arr = (_ffi.Prior * len(priors))()
        self._names = []
        for k, pr in enumerate(priors):
            name, params, bounds = prior_fields(pr)
            self._names.append(name.encode())
            pr_arr[k].name = self._names[-1]
            pr_arr[k].nparams = len(params)
            for i, v in enumerate(params):
                pr_arr[k].params[i] = v
            pr_arr[k].has_bounds = 0 if bounds is None else 1
            pr_arr[k].lower, pr_arr[k].upper = (0.0, 0.0) if bounds is None else (bounds[0], bounds[1])
        desc = _ffi.ModelDesc(self.ntypes, self.nevents, self.nparams, self.ndep, _ffi.dptr(e_mat), _ffi.iptr(p_vec), fd_arr, pr_arr,
                              self.kind)
        self.handle = C.c_void_p()
        rc = _ffi.lib().bpc_model_create(C.byref(desc), C.byref(self.handle))
        if rc in (_ffi.E_INVALID, _ffi.E_EXPR, _ffi.E_PRIOR):
            raise ValueError(_ffi.lib().bpc_last_error().decode())
        _ffi.check(rc)
        self.dim = _ffi.lib().bpc_model_dim(self.handle)
        self.model, self.priors = model, list(priors)

    # -- inspection ---------------------------------------------------------------------------
    @property
    def cuda_source(self) -> str:
        return _ffi.lib().bpc_model_cuda_source(self.handle).decode()

    @property
    def warnings(self) -> str:
        return _ffi.lib().bpc_model_warnings(self.handle).decode()

    def cubin(self) -> bytes:
        p, n = C.c_void_p(), C.c_ulonglong()
        _ffi.check(_ffi.lib().bpc_model_cubin(self.handle, C.byref(p), C.byref(n)))
        return C.string_at(p, n.value)

    # -- data ---------------------------------------------------------------------------------
    def data(self, dat: dict, device: int = 0, grouping: int = 0) -> StanData:
        return StanData(self, dat, device, grouping)

    def stan_data_struct(self, dat: dict):
        return _stan_data_struct(self, dat)

    def _data(self, dat) -> StanData:
        return dat if isinstance(dat, StanData) else StanData(self, dat)

    # -- evaluate -----------------------------------------------------------------------------
    def log_prob_grad(self, dat, upars, adjust_transform: bool = True, gradient: bool = True):
        """(lp [C], grad [C, dim] or None, status [C]) at unconstrained vectors upars [C, dim]."""
        d = self._data(dat)
        u = np.ascontiguousarray(np.atleast_2d(_f64(upars)))
        if u.shape[1] != self.dim:
            raise ValueError("upars must have %d columns" % self.dim)
        Cn = u.shape[0]
        lp = np.empty(Cn)
        grad = np.empty((Cn, self.dim)) if gradient else None
        status = np.empty(Cn, dtype=np.int32)
        _ffi.check(_ffi.lib().bpc_log_prob(self.handle, d.handle, _ffi.dptr(u), Cn, int(bool(adjust_transform)), _ffi.dptr(lp),
                                           _ffi.dptr(grad) if gradient else None, _ffi.iptr(status)))
        return lp, grad, status

    def log_prob(self, dat, upars, adjust_transform: bool = True, gradient: bool = False):
        """rstan::log_prob: value(s); with gradient=True returns (lp, grad)."""
        lp, g, _ = self.log_prob_grad(dat, upars, adjust_transform, gradient)
        one = np.ndim(upars) == 1
        if gradient:
            return (lp[0], g[0]) if one else (lp, g)
        return lp[0] if one else lp

    def grad_log_prob(self, dat, upars, adjust_transform: bool = True):
        """rstan::grad_log_prob."""
        _, g, _ = self.log_prob_grad(dat, upars, adjust_transform, True)
        return g[0] if np.ndim(upars) == 1 else g

    def log_lik(self, dat, upars):
        """per-observation log_lik [C, ndatapts] of the LOO blocks (generate(..., loo = T))."""
        d = self._data(dat)
        u = np.ascontiguousarray(np.atleast_2d(_f64(upars)))
        out = np.empty((u.shape[0], d.ndatapts))
        _ffi.check(_ffi.lib().bpc_log_lik(self.handle, d.handle, _ffi.dptr(u), u.shape[0], _ffi.dptr(out)))
        return out

    def moments(self, dat, upars):
        """(mu [C, N, n], Sigma [C, N, n, n]) of the final population of every observation (bpc_moments)."""
        d = self._data(dat)
        u = np.ascontiguousarray(np.atleast_2d(_f64(upars)))
        n = self.ntypes
        mu = np.empty((u.shape[0], d.ndatapts, n))
        sig = np.empty((u.shape[0], d.ndatapts, n, n))
        _ffi.check(_ffi.lib().bpc_moments(self.handle, d.handle, _ffi.dptr(u), u.shape[0], _ffi.dptr(mu), _ffi.dptr(sig)))
        return mu, sig

    def flop_model(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        _ffi.check(_ffi.lib().bpc_flop_model(self.handle, C.byref(a), C.byref(b), C.byref(c)))
        out = dict(per_app_fwd=a.value, per_app_bwd=b.value, per_obs=c.value)
        _ffi.check(_ffi.lib().bpc_flop_model_centred(self.handle, C.byref(a), C.byref(b), C.byref(c)))
        out.update(per_term_c=a.value, per_obs_c=b.value, per_obs_o=c.value)
        return out

    def constrain_pars(self, upars):
        u = np.ascontiguousarray(np.atleast_2d(_f64(upars)))
        out = np.empty_like(u)
        _ffi.check(_ffi.lib().bpc_constrain(self.handle, _ffi.dptr(u), u.shape[0], _ffi.dptr(out)))
        return out[0] if np.ndim(upars) == 1 else out

    def unconstrain_pars(self, theta):
        t = np.ascontiguousarray(np.atleast_2d(_f64(theta)))
        out = np.empty_like(t)
        _ffi.check(_ffi.lib().bpc_unconstrain(self.handle, _ffi.dptr(t), t.shape[0], _ffi.dptr(out)))
        return out[0] if np.ndim(theta) == 1 else out

    def sampling(self, dat, chains: int = 4, iter: int = 2000, warmup=None, init="random", control=None, seed: int = 1,
                 pooled_adaptation: bool = False, device=None, loo: bool = False, init_opt_steps: int = 0, init_opt_subsample: int = 0):
        """rstan::sampling — see bpinference_b200.fit.sampling."""
        from .fit import sampling
        return sampling(self, dat, chains=chains, iter=iter, warmup=warmup, init=init, control=control, seed=seed,
                        pooled_adaptation=pooled_adaptation, device=device, loo=loo, init_opt_steps=init_opt_steps,
                        init_opt_subsample=init_opt_subsample)

    def sampling_multi(self, dat, devices, chains: int = 4, iter: int = 2000, warmup=None, init="random", control=None, seed: int = 1,
                       pooled_adaptation: bool = False, init_opt_steps: int = 0):
        """the same run over several GPUs of this process (bpc_sample_multi) — see bpinference_b200.fit.sampling_multi."""
        from .fit import sampling_multi
        return sampling_multi(self, dat, devices, chains=chains, iter=iter, warmup=warmup, init=init, control=control, seed=seed,
                              pooled_adaptation=pooled_adaptation, init_opt_steps=init_opt_steps)

    def flops_per_chain(self, dat, upars):
        d = self._data(dat)
        u = np.ascontiguousarray(np.atleast_2d(_f64(upars)))
        out = np.empty(u.shape[0])
        _ffi.check(_ffi.lib().bpc_count_flops(self.handle, d.handle, _ffi.dptr(u), u.shape[0], _ffi.dptr(out)))
        return out

    def close(self):
        if getattr(self, "handle", None):
            _ffi.lib().bpc_model_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def stan_model(model, priors, simple_bd: bool = False, noise_model: bool = False) -> BpCudaModel:
    """generate() + rstan::stan_model in one step: returns the compiled CUDA model."""
    return BpCudaModel(model, priors, simple_bd=simple_bd, noise_model=noise_model)


def fp64_peak(device: int = 0, ms: float = 200.0):
    """(TFLOP/s, SM clock MHz) of a register-resident DFMA loop on `device`."""
    t, c = C.c_double(), C.c_double()
    _ffi.check(_ffi.lib().bpc_fp64_peak(device, ms, C.byref(t), C.byref(c)))
    return t.value, c.value


def fp64_tensor_peak(device: int = 0, ms: float = 200.0):
    """TFLOP/s of a register-resident DMMA (mma.sync m8n8k4 f64) loop on `device`."""
    t = C.c_double()
    _ffi.check(_ffi.lib().bpc_fp64_tensor_peak(device, ms, C.byref(t)))
    return t.value
