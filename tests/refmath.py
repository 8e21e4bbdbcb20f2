"""Independent numpy statements of the moment math, used to pin the oracle (test helper).

* `vanloan_moments`  — scipy.linalg.expm of the Van Loan block matrix [[A, Cm],[0, A(+)A]]
                       (SURVEY.md §8a row A5): M = exp(At), D = M D(0) + J.
* `lp_complex`       — the per-observation score evaluated in complex arithmetic so that
                       complex-step differentiation gives d lp / d rate to ~1e-15 without any adjoint code.
"""
import math

import numpy as np
import scipy.linalg as sla


def generator(e_mat, p_vec, r):
    e_mat = np.asarray(e_mat, dtype=float)
    E, n = e_mat.shape
    r = np.asarray(r)
    A = np.zeros((n, n), r.dtype)
    C = np.zeros((n, n, n), r.dtype)
    for e in range(E):
        i = int(p_vec[e]) - 1
        A[i] += r[e] * e_mat[e]
        A[i, i] -= r[e]
        C[i] += r[e] * (np.outer(e_mat[e], e_mat[e]) - np.diag(e_mat[e]))
    H = np.zeros((n, n, n), r.dtype)
    for l in range(n):
        el = np.zeros(n)
        el[l] = 1
        H[l] = C[l] + np.diag(A[l]) - np.outer(A[l], el) - np.outer(el, A[l])
    return A, H, C


def vanloan_moments(e_mat, p_vec, r, t):
    """M [n,n] and raw second moments D [n, n*n] (column j + n*k) as the Stan template defines them."""
    A, _, C = generator(e_mat, p_vec, np.asarray(r, dtype=float))
    n = A.shape[0]
    K = np.kron(A, np.eye(n)) + np.kron(np.eye(n), A)
    Cm = np.stack([C[i].reshape(-1, order="F") for i in range(n)])
    Bm = np.block([[A, Cm], [np.zeros((n * n, n)), K]])
    Ex = sla.expm(Bm * t)
    M, J = Ex[:n, :n], Ex[:n, n:]
    D0 = np.zeros((n, n * n))
    for i in range(n):
        D0[i, i + n * i] = 1
    return M, M @ D0 + J


def _taylor(A, H, t, mu, S, terms=40, sub=8):
    h = t / sub
    for _ in range(sub):
        wm, wS = mu, S
        for k in range(terms):
            lm = A.T @ wm
            lS = A.T @ wS + wS @ A + np.einsum("l,lij->ij", wm, H)
            wm, wS = lm * (h / (k + 1)), lS * (h / (k + 1))
            mu, S = mu + wm, S + wS
    return mu, S


def lp_complex(e_mat, p_vec, r, t, z0, x, sigma_obs=None):
    A, H, _ = generator(e_mat, p_vec, r)
    n = len(z0)
    mu, S = _taylor(A, H, t, np.asarray(z0).astype(A.dtype), np.zeros((n, n), A.dtype))
    if sigma_obs is not None:
        S = S + sigma_obs * np.diag(mu)
    Lc = np.zeros_like(S)
    for j in range(n):
        d = S[j, j] - (Lc[j, :j] ** 2).sum()
        Lc[j, j] = np.sqrt(d)
        for i in range(j + 1, n):
            Lc[i, j] = (S[i, j] - (Lc[i, :j] * Lc[j, :j]).sum()) / Lc[j, j]
    res = np.asarray(x) - mu
    a = np.zeros_like(res)
    for i in range(n):
        a[i] = (res[i] - (Lc[i, :i] * a[:i]).sum()) / Lc[i, i]
    return -np.log(np.diag(Lc)).sum() - 0.5 * (a * a).sum()


def complex_step_grad(f, r, h=1e-30):
    r = np.asarray(r, dtype=float)
    g = np.zeros_like(r)
    for e in range(len(r)):
        rc = r.astype(complex)
        rc[e] += 1j * h
        g[e] = f(rc).imag / h
    return g


def sidx(i, j, n):
    """index of (i, j) in the packed upper triangle of a symmetric n x n matrix (bp_sidx of bp_kernels.cuh)."""
    if i > j:
        i, j = j, i
    return i * n - i * (i - 1) // 2 + (j - i)


def unpack_sym(packed, n):
    S = np.zeros((n, n))
    for i in range(n):
        for j in range(n):
            S[i, j] = packed[sidx(i, j, n)]
    return S


def moment_operator(e_mat, p_vec, r):
    """the closed linear operator L of the per-observation state y = (mu, packed Sigma):  mu' = A^T mu,
    Sigma' = A^T Sigma + Sigma A + sum_l mu_l H_l  (DESIGN.md §2), as a dense [D, D] matrix in packed coordinates,
    built column by column from the right-hand side applied to unit vectors."""
    A, H, _ = generator(e_mat, p_vec, np.asarray(r, dtype=float))
    n = A.shape[0]
    s = n * (n + 1) // 2
    D = n + s
    Lm = np.zeros((D, D))
    for c in range(D):
        mu = np.zeros(n)
        S = np.zeros((n, n))
        if c < n:
            mu[c] = 1.0
        else:
            for i in range(n):
                for j in range(i, n):
                    if sidx(i, j, n) == c - n:
                        S[i, j] = S[j, i] = 1.0
        dmu = A.T @ mu
        dS = A.T @ S + S @ A + np.einsum("l,lij->ij", mu, H)
        Lm[:n, c] = dmu
        for i in range(n):
            for j in range(i, n):
                Lm[n + sidx(i, j, n), c] = dS[i, j]
    return Lm
