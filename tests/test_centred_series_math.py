"""The identities the centred / octet W paths rest on, checked in numpy with random data — no GPU, no library:

  forward   exp((tau + h) L) y0 = sum_j c_j(h) N_j y0,          N_j = L^j exp(tau L),       c_j(x) = x^j / j!
  reverse   W_n = sum_k c_n(t_k) b_k y0_k^T  =  sum_g sum_j c_(n-j)(tau_g) S^g_j,   S^g_j = sum_(k in g) c_j(h_k) b_k y0_k^T
  gradient  d/dL sum_k b_k^T exp(t_k L) y0_k  =  sum_i Y_i P^i,   Y_i = W_(i+1) + P Y_(i+1),   P = L^T     (bp_wpost)
  sub-steps y_s = F^s y_0, F = exp(h L):  dlp/dL = the same polynomial of W_j = c_j(h) Q,   Q = sum_i z_(i+1) y_i^T,  z_i = F^T z_(i+1)
            (bp_grouped's cooperative multi-step path)
"""
import math

import numpy as np
from scipy.linalg import expm


def _c(x, j):
    return x ** j / math.factorial(j)


def _setup(seed=0, D=9, n=3, K=40):
    rng = np.random.default_rng(seed)
    L = rng.normal(size=(D, D)) * 0.25
    y0 = np.zeros((K, D)); y0[:, :n] = rng.uniform(1.0, 5.0, size=(K, n))     # states start as (z0, 0)
    b = rng.normal(size=(K, D))
    t = np.sort(rng.uniform(0.5, 2.0, K))[::-1]
    return rng, L, y0, b, t


def test_forward_series_around_a_centre():
    _, L, y0, _, t = _setup()
    tau = 1.3
    M = 40
    N = [np.linalg.matrix_power(L, j) @ expm(tau * L) for j in range(M)]
    for k in range(len(t)):
        h = t[k] - tau
        y = sum(_c(h, j) * (N[j] @ y0[k]) for j in range(M))
        np.testing.assert_allclose(y, expm(t[k] * L) @ y0[k], rtol=1e-12, atol=1e-13)


def test_w_matrices_factor_through_the_groups_toeplitz_product():
    _, L, y0, b, t = _setup()
    M = 45
    W = [sum(_c(t[k], n_) * np.outer(b[k], y0[k]) for k in range(len(t))) for n_ in range(M)]
    groups = np.array_split(np.arange(len(t)), 5)
    W2 = [np.zeros_like(W[0]) for _ in range(M)]
    for idx in groups:
        tau = 0.5 * (t[idx].max() + t[idx].min())
        S = [sum(_c(t[k] - tau, j) * np.outer(b[k], y0[k]) for k in idx) for j in range(M)]
        for n_ in range(M):
            for j in range(n_ + 1):
                W2[n_] += _c(tau, n_ - j) * S[j]
    for n_ in range(M):
        np.testing.assert_allclose(W2[n_], W[n_], rtol=1e-11, atol=1e-13)


def _lbar_from_w(W, L):
    """sum_i Y_i P^i with Y_i = W_(i+1) + P Y_(i+1), P = L^T; W[j - 1] holds W_j."""
    P = L.T
    M = len(W)
    Y = [None] * M
    Y[M - 1] = W[M - 1]
    for i in range(M - 2, -1, -1):
        Y[i] = W[i] + P @ Y[i + 1]
    Z = Y[M - 1]
    for i in range(M - 2, -1, -1):
        Z = Y[i] + Z @ P
    return Z


def _fd_grad(f, L, eps=1e-6):
    G = np.zeros_like(L)
    for r in range(L.shape[0]):
        for c in range(L.shape[1]):
            E = np.zeros_like(L); E[r, c] = eps
            G[r, c] = (f(L + E) - f(L - E)) / (2 * eps)
    return G


def test_gradient_is_the_matrix_polynomial_of_the_w_matrices():
    _, L, y0, b, t = _setup(D=5, n=2, K=12)
    M = 45
    W = [sum(_c(t[k], j) * np.outer(b[k], y0[k]) for k in range(len(t))) for j in range(1, M + 1)]
    G = _lbar_from_w(W, L)
    f = lambda Lx: sum(b[k] @ expm(t[k] * Lx) @ y0[k] for k in range(len(t)))
    np.testing.assert_allclose(G, _fd_grad(f, L), rtol=2e-7, atol=1e-7)


def test_sub_steps_share_one_series_so_their_gradient_is_the_polynomial_of_c_j_times_q():
    rng, L, _, _, _ = _setup(D=5, n=2)
    D, n, s, t = 5, 2, 7, 3.0
    h = t / s
    Y0 = np.zeros((D, n)); Y0[:n, :n] = np.eye(n)
    Zs = rng.normal(size=(D, n))                       # cotangent of Y_s
    F = expm(h * L)
    Y = [Y0]
    for _ in range(s):
        Y.append(F @ Y[-1])
    Z, Q = Zs, np.zeros((D, D))
    for i in range(s - 1, -1, -1):
        Q += Z @ Y[i].T
        Z = F.T @ Z
    M = 40
    G = _lbar_from_w([_c(h, j) * Q for j in range(1, M + 1)], L)
    f = lambda Lx: float(np.sum(Zs * (np.linalg.matrix_power(expm(h * Lx), s) @ Y0)))
    np.testing.assert_allclose(G, _fd_grad(f, L), rtol=2e-7, atol=1e-7)
