"""Synthetic workloads of the five BASELINE.json configs.

Data generation only: observations are drawn from the model's own Gaussian approximation at the
"true" parameters (x ~ N(mu, Sigma), rounded, clipped at 0), so that Sigma is well conditioned.
Nothing here is on the log-density hot path; the moments used for drawing come from a plain
batched numpy series and need not be exact.

Shapes follow the reference's example scripts (paths under /root/reference):
  cfg1  examples/one_type.R                 one-type template, N = 19*25, 4 chains
  cfg2  examples/two_type.R                 two-type, N = 5*50, 1024 chains
  cfg3  examples/one_type_logistic_birth.R  one-type template with env-dependent birth rate, N = 600
  cfg4  examples/apoptosis_with_noise.R     two-type with observation noise, N = 250
  cfg5  synthetic three-type                N = 100 000, 4096 chains (512 per GPU)
        5U: every observation has its own duration t_k ~ U(0.5, 5)   (ntimes_unique = N)
        5G: six distinct durations                                   (six (level, duration) groups)
        5K: 5U plus one env variable x_k ~ U(0, 1) per observation   (ndep_levels = N)
"""
from __future__ import annotations

import numpy as np


def _generator(e_mat, p_vec, r):
    """A [n,n], H [n,n,n] of the forward moment ODE  mu' = A^T mu,  S' = A^T S + S A + sum_i mu_i H_i."""
    E, n = e_mat.shape
    A = np.zeros((n, n))
    C = np.zeros((n, n, n))
    for e in range(E):
        i = int(p_vec[e]) - 1
        A[i] += r[e] * e_mat[e]
        A[i, i] -= r[e]
        C[i] += r[e] * (np.outer(e_mat[e], e_mat[e]) - np.diag(e_mat[e]))
    H = np.zeros((n, n, n))
    for l in range(n):
        el = np.zeros(n)
        el[l] = 1.0
        H[l] = C[l] + np.diag(A[l]) - np.outer(A[l], el) - np.outer(el, A[l])
    return A, H


def gaussian_moments(e_mat, p_vec, r, z0, t, substeps=8, terms=24):
    """batched mean / covariance of the population after duration t.  r: [E] or [N,E]; z0 [N,n]; t [N]."""
    e_mat = np.asarray(e_mat, dtype=float)
    z0 = np.atleast_2d(np.asarray(z0, dtype=float))
    N, n = z0.shape
    t = np.broadcast_to(np.asarray(t, dtype=float), (N,))
    r = np.asarray(r, dtype=float)
    if r.ndim == 1:
        A, H = _generator(e_mat, p_vec, r)
        A = np.broadcast_to(A, (N, n, n))
        H = np.broadcast_to(H, (N, n, n, n))
    else:
        AH = [_generator(e_mat, p_vec, rr) for rr in r]
        A = np.stack([a for a, _ in AH])
        H = np.stack([h for _, h in AH])
    mu, S = z0.copy(), np.zeros((N, n, n))
    h = t / substeps
    At = np.swapaxes(A, 1, 2)
    for _ in range(substeps):
        wm, wS = mu, S
        for k in range(terms):
            lm = np.einsum("nij,nj->ni", At, wm)
            lS = At @ wS + wS @ A + np.einsum("nl,nlij->nij", wm, H)
            c = (h / (k + 1))
            wm, wS = lm * c[:, None], lS * c[:, None, None]
            mu, S = mu + wm, S + wS
    return mu, S


def _draw(rng, mu, S):
    N, n = mu.shape
    Lc = np.linalg.cholesky(S + 1e-9 * np.eye(n))
    x = mu + np.einsum("nij,nj->ni", Lc, rng.standard_normal((N, n)))
    return np.clip(np.rint(x), 0, None)


def _trajectories(rng, e_mat, p_vec, r, z0, nsteps, nrep, dt=1.0, thin=None):
    """nrep replicates observed at nsteps+1 equally spaced times; returns (init_pop, final_pop) of the
    consecutive pairs, in the (replicate-major) order stan_data_from_simulation produces."""
    n = len(z0)
    cur = np.tile(np.asarray(z0, dtype=float), (nrep, 1))
    obs = [cur if thin is None else rng.binomial(cur.astype(np.int64), thin).astype(float)]
    for _ in range(nsteps):
        mu, S = gaussian_moments(e_mat, p_vec, r, cur, dt)
        cur = _draw(rng, mu, S)
        obs.append(cur if thin is None else rng.binomial(cur.astype(np.int64), thin).astype(float))
    obs = np.stack(obs, axis=1)                       # [rep, time, n]
    init = obs[:, :-1].reshape(-1, n)
    fin = obs[:, 1:].reshape(-1, n)
    return init, fin


def unconstrain(theta, priors):
    u = np.array(theta, dtype=float)
    for k, pr in enumerate(priors):
        b = pr.get("bounds")
        if b is not None:
            lo, hi = b
            s = (theta[k] - lo) / (hi - lo)
            u[k] = np.log(s / (1 - s))
    return u


def _finish(cfg, rng, chains, jitter=0.1):
    u_true = unconstrain(cfg["theta_true"], cfg["priors"])
    if cfg["kind"] == "multitype_noise":
        u_true = np.append(u_true, np.log(cfg["sigma_obs_true"]))
    cfg["u_true"] = u_true
    cfg["chains"] = chains
    cfg["upars"] = u_true[None, :] + jitter * rng.standard_normal((chains, len(u_true)))
    return cfg


def make_config(name: str, seed: int = 1, chains: int | None = None, nobs: int | None = None) -> dict:
    """returns dict(kind, e_mat, p_vec, func_deps, nparams, ndep, priors, data, theta_true, chains, upars)."""
    rng = np.random.default_rng(seed)
    if name == "cfg1":
        e_mat, p_vec = np.array([[2.0], [0.0]]), np.array([1, 1])
        theta = np.array([0.1, 0.05])
        nrep = 25 if nobs is None else max(1, nobs // 19)
        init, fin = _trajectories(rng, e_mat, p_vec, theta, [1000.0], 19, nrep)
        N = init.shape[0]
        data = dict(pop_vec=fin[:, 0], init_pop=init[:, 0], times=np.ones(N), function_var=np.zeros((1, 1)),
                    var_idx=np.ones(N, dtype=np.int32), ndep_levels=1, ndep=1, nparams=2, ndatapts=N)
        cfg = dict(name=name, kind="simple_bd", e_mat=e_mat, p_vec=p_vec, func_deps=["c[1]", "c[2]"], nparams=2, ndep=0,
                   priors=[dict(name="uniform", params=[0.0, 1.0], bounds=[0.0, 2.0])] * 2, data=data, theta_true=theta)
        return _finish(cfg, rng, chains or 4)
    if name == "cfg2":
        e_mat = np.array([[2, 0], [1, 1], [1, 1], [0, 2], [0, 0], [0, 0]], dtype=float)
        p_vec = np.array([1, 1, 2, 2, 1, 2])
        theta = np.array([0.20, 0.10, 0.05, 0.20, 0.25, 0.20])
        N = nobs or 250
        z0 = np.rint(np.array([1000.0, 500.0])[None, :] * rng.uniform(0.8, 1.5, size=(N, 1)))
        mu, S = gaussian_moments(e_mat, p_vec, theta, z0, 1.0)
        data = dict(ntypes=2, nevents=6, ndatapts=N, ntimes_unique=1, ndep_levels=1, ndep=1, nparams=6, e_mat=e_mat,
                    p_vec=p_vec, pop_vec=_draw(rng, mu, S), init_pop=z0, times=np.array([1.0]),
                    times_idx=np.ones(N, dtype=np.int32), function_var=np.zeros((1, 1)), var_idx=np.ones(N, dtype=np.int32))
        cfg = dict(name=name, kind="multitype", e_mat=e_mat, p_vec=p_vec, func_deps=["c[%d]" % (k + 1) for k in range(6)],
                   nparams=6, ndep=0, priors=[dict(name="normal", params=[0.0, 0.25], bounds=[0.0, 5.0])] * 6, data=data,
                   theta_true=theta)
        return _finish(cfg, rng, chains or 1024)
    if name == "cfg3":
        e_mat, p_vec = np.array([[2.0], [0.0]]), np.array([1, 1])
        theta = np.array([0.15, 0.2, 10.0, 0.5, 0.10])
        xs = np.array([0.0, 0.2, 0.4, 0.6, 0.8, 1.0])
        nrep = 20 if nobs is None else max(1, nobs // 30)
        init, fin, vidx = [], [], []
        for v, x in enumerate(xs):
            b = theta[0] + theta[1] / (1 + np.exp(theta[2] * (x - theta[3])))
            i_, f_ = _trajectories(rng, e_mat, p_vec, np.array([b, theta[4]]), [1000.0], 5, nrep)
            init.append(i_), fin.append(f_), vidx.append(np.full(i_.shape[0], v + 1))
        init, fin, vidx = np.concatenate(init), np.concatenate(fin), np.concatenate(vidx).astype(np.int32)
        N = init.shape[0]
        data = dict(pop_vec=fin[:, 0], init_pop=init[:, 0], times=np.ones(N), function_var=xs[:, None], var_idx=vidx,
                    ndep_levels=6, ndep=1, nparams=5, ndatapts=N)
        pri = [dict(name="normal", params=[0.0, 0.25], bounds=[0.0, 5.0]) for _ in range(5)]
        pri[2] = dict(name="uniform", params=[5.0, 15.0], bounds=[5.0, 15.0])
        cfg = dict(name=name, kind="simple_bd", e_mat=e_mat, p_vec=p_vec,
                   func_deps=["c[1] + c[2]/(1 + exp(c[3]*(x[1] - c[4])))", "c[5]"], nparams=5, ndep=1, priors=pri,
                   data=data, theta_true=theta)
        return _finish(cfg, rng, chains or 4)
    if name == "cfg4":
        e_mat = np.array([[2, 0], [0, 1], [0, 0]], dtype=float)
        p_vec = np.array([1, 1, 2])
        theta = np.array([0.08, 0.05, 0.1])
        nrep = 50 if nobs is None else max(1, nobs // 5)
        init, fin = _trajectories(rng, e_mat, p_vec, theta, [1000.0, 0.0], 5, nrep, thin=0.65)
        N = init.shape[0]
        data = dict(ntypes=2, nevents=3, ndatapts=N, ntimes_unique=1, ndep_levels=1, ndep=1, nparams=3, e_mat=e_mat,
                    p_vec=p_vec, pop_vec=fin, init_pop=init, times=np.array([1.0]), times_idx=np.ones(N, dtype=np.int32),
                    function_var=np.zeros((1, 1)), var_idx=np.ones(N, dtype=np.int32))
        cfg = dict(name=name, kind="multitype_noise", e_mat=e_mat, p_vec=p_vec, func_deps=["c[1]", "c[2]", "c[3]"],
                   nparams=3, ndep=0, priors=[dict(name="normal", params=[0.0, 0.25], bounds=[0.0, 2.0])] * 3, data=data,
                   theta_true=theta, sigma_obs_true=0.05)
        return _finish(cfg, rng, chains or 4)
    if name in ("cfg5U", "cfg5G", "cfg5K"):
        n = 3
        e_mat = np.zeros((9, 3))
        p_vec = np.zeros(9, dtype=np.int64)
        for i in range(3):
            e_mat[3 * i, i] = 2.0                       # division: two of type i
            e_mat[3 * i + 1, (i + 1) % 3] = 1.0         # transition to type i+1
            p_vec[3 * i: 3 * i + 3] = i + 1             # third event: death (no offspring)
        theta = np.array([0.25, 0.05, 0.10, 0.20, 0.04, 0.12, 0.22, 0.06, 0.09])
        N = nobs or 100_000
        z0 = np.rint(rng.uniform(500, 5000, size=(N, n)))
        func_deps = ["c[%d]" % (k + 1) for k in range(9)]
        P, K = 9, 0
        fvar, vidx, L = np.zeros((1, 1)), np.ones(N, dtype=np.int32), 1
        if name == "cfg5G":
            tu = np.array([0.5, 1.0, 2.0, 3.0, 4.0, 5.0])
            tidx = rng.integers(1, 7, size=N).astype(np.int32)
            t = tu[tidx - 1]
            r = theta
        else:
            t = rng.uniform(0.5, 5.0, size=N)
            tu, tidx = t.copy(), np.arange(1, N + 1, dtype=np.int32)
            r = theta
            if name == "cfg5K":
                x = rng.uniform(0.0, 1.0, size=N)
                theta = np.append(theta, 0.5)
                func_deps[0] = "c[1]*exp(-c[10]*x[1])"
                P, K = 10, 1
                fvar, vidx, L = x[:, None], np.arange(1, N + 1, dtype=np.int32), N
                r = np.tile(theta[:9], (N, 1))
                r[:, 0] = theta[0] * np.exp(-theta[9] * x)
        mu, S = gaussian_moments(e_mat, p_vec, r, z0, t)
        data = dict(ntypes=3, nevents=9, ndatapts=N, ntimes_unique=len(tu), ndep_levels=L, ndep=max(K, 1), nparams=P,
                    e_mat=e_mat, p_vec=p_vec, pop_vec=_draw(rng, mu, S), init_pop=z0, times=tu, times_idx=tidx,
                    function_var=fvar, var_idx=vidx)
        cfg = dict(name=name, kind="multitype", e_mat=e_mat, p_vec=p_vec, func_deps=func_deps, nparams=P, ndep=K,
                   priors=[dict(name="normal", params=[0.0, 0.25], bounds=[0.0, 5.0]) for _ in range(P)], data=data,
                   theta_true=theta)
        return _finish(cfg, rng, chains or 512)
    raise ValueError("unknown config " + name)


CONFIG_NAMES = ("cfg1", "cfg2", "cfg3", "cfg4", "cfg5U", "cfg5G", "cfg5K")
