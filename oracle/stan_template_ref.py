"""CPU ORACLE — literal numpy restatement of the Stan template's moment ODE (TEST INFRASTRUCTURE ONLY).

Follows stanfiles/multitype_template.stan line by line (paths under /root/reference):
  moment_ode        :5-62     (parameter unpacking :23-24, lambda/b_mat/a_mat :26-31, c_mat :33-43,
                               beta_mat :50-52, first moments :55, second moments :58)
  init_state        :85-94    (to_array_1d of a real[,] / real[,,] array is row-major)
  mu / sigma        :137-149
  score             :152      (multi_normal under `~`: additive constant dropped)
and multitype_noise_template.stan:153 for the noise variant.

The ODE is solved with SciPy: `rk45()` mimics what rstan's integrate_ode_rk45 does at its default
tolerances (rel = abs = 1e-6, Dormand-Prince 5(4); SciPy's controller is not Boost's, so this bounds
the gap rather than reproducing rstan bit for bit) and `exact()` uses DOP853 at 1e-13.  tests/ use
both to pin oracle/bp_oracle.hpp and to measure the closed-form-vs-RK45 gap.
"""
import numpy as np
from scipy.integrate import solve_ivp


def moment_ode(t, state, e_mat, r_mat):
    E, n = e_mat.shape
    lam = r_mat.sum(axis=0)                       # :27
    b_mat = r_mat.T @ e_mat                       # :28
    a_mat = b_mat - np.diag(lam)                  # :29-30
    c_mat = np.zeros((n, n, n))
    for a in range(n):                            # :33-43
        e_tmp = e_mat * e_mat[:, [a]]
        b2 = r_mat.T @ e_tmp
        for i in range(n):
            c_mat[i][a] = b2[i]
            c_mat[i][a, a] = b2[i, a] - b_mat[i, a]
    mt = state[:n * n].reshape((n, n), order="F")             # :47 to_matrix is column-major
    dt = state[n * n:].reshape((n, n * n), order="F")         # :48
    beta = np.zeros((n, n * n))
    for i in range(n):                                        # :50-52
        beta[i] = (mt.T @ c_mat[i] @ mt).reshape(-1, order="F")
    mtp = a_mat @ mt                                          # :55
    dtp = a_mat @ dt + beta                                   # :58
    return np.concatenate([mtp.reshape(-1, order="F"), dtp.reshape(-1, order="F")])


def init_state(n):
    m = np.eye(n)
    d = np.zeros((n, n, n))
    for i in range(n):
        d[i, i, i] = 1.0
    return np.concatenate([m.reshape(-1), d.reshape(-1)])


def r_matrix(e_mat, p_vec, r):
    """transformed parameters :108-116: r_mat[e, p_vec[e]] = rate_e, zero elsewhere."""
    E, n = e_mat.shape
    r_mat = np.zeros((E, n))
    r_mat[np.arange(E), np.asarray(p_vec) - 1] = r
    return r_mat


def solve(e_mat, p_vec, r, times, rtol, atol, method):
    e_mat = np.asarray(e_mat, dtype=float)
    n = e_mat.shape[1]
    times = np.atleast_1d(np.asarray(times, dtype=float))
    order = np.argsort(times)
    sol = solve_ivp(moment_ode, (0.0, float(times.max())), init_state(n), t_eval=times[order],
                    args=(e_mat, r_matrix(e_mat, p_vec, np.asarray(r, dtype=float))), rtol=rtol, atol=atol, method=method,
                    first_step=min(0.1, float(times.max())))
    out = np.empty((len(times), n * n + n ** 3))
    out[order] = sol.y.T
    return out


def rk45(e_mat, p_vec, r, times):
    return solve(e_mat, p_vec, r, times, 1e-6, 1e-6, "RK45")


def exact(e_mat, p_vec, r, times):
    return solve(e_mat, p_vec, r, times, 1e-13, 1e-13, "DOP853")


def mu_sigma(moments_row, z0):
    z0 = np.asarray(z0, dtype=float)
    n = len(z0)
    m_t = moments_row[:n * n].reshape((n, n), order="F")              # :139
    d_t = moments_row[n * n:].reshape((n, n * n), order="F")          # :140
    mu = m_t.T @ z0                                                   # :143
    temp = (z0 @ d_t).reshape((n, n), order="F")                      # :144
    sig = np.empty((n, n))
    for i in range(n):
        for j in range(n):
            sig[i, j] = temp[i, j] - z0 @ (m_t[:, i] * m_t[:, j])     # :147
    return mu, sig


def likelihood(e_mat, p_vec, r_levels, data, sigma_obs=None, solver=exact):
    """sum over observations of the multi_normal score (constant dropped) for one chain.
    r_levels: [L, E] rates per env level."""
    n, N = int(data["ntypes"]), int(data["ndatapts"])
    pop = np.asarray(data["pop_vec"], dtype=float).reshape(N, n)
    init = np.asarray(data["init_pop"], dtype=float).reshape(N, n)
    times = np.atleast_1d(np.asarray(data["times"], dtype=float))
    lp = 0.0
    mom = [solver(e_mat, p_vec, r_levels[v], times) for v in range(r_levels.shape[0])]     # :132-134
    for k in range(N):
        v, t = int(data["var_idx"][k]) - 1, int(data["times_idx"][k]) - 1
        mu, sig = mu_sigma(mom[v][t], init[k])
        if sigma_obs is not None:
            sig = sig + sigma_obs * np.diag(mu)
        sgn, ld = np.linalg.slogdet(sig)
        res = pop[k] - mu
        lp += -0.5 * ld - 0.5 * res @ np.linalg.solve(sig, res)
    return lp
