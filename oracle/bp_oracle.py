"""CPU ORACLE — Python side (TEST INFRASTRUCTURE ONLY, NOT PRODUCT CODE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  The product (bpinference_b200 + libbpcuda.so) never does.

It restates, for one batch of unconstrained parameter vectors, what rstan's log_prob /
grad_log_prob return for the three bpinference Stan templates (reference paths under
/root/reference):

  * parameter declarations + bound transform   R/stan_utils.R:397-410 (emitted `real<lower,upper> ThetaK`)
                                               stanfiles/multitype_noise_template.stan:103 (`sigma_obs`)
  * rate expressions                           R/stan_utils.R:205-281 (grammar of exp_to_stan)
                                               stanfiles/multitype_template.stan:108-116, one_type_template.stan:20-26
  * priors  `ThetaK ~ name(params)`            R/stan_utils.R:395; list ALLOWED_DISTRIBUTIONS in R/sysdata.rda
  * likelihood                                 oracle/bp_oracle.hpp (C++; cites the template lines)

Heavy per-(chain, observation) arithmetic is done by oracle/libbporacle.so (built by oracle/Makefile);
the per-chain pieces (transform, rates and their Jacobian by forward-mode duals, priors, chain rule)
are numpy.  PARITY UNPINNED against rstan (see bp_oracle.hpp header): densities follow Stan's `~`
convention of dropping additive constants [recalled from the Stan reference manual, not checkable here].
"""
from __future__ import annotations

import ctypes
import math
import os
import re
import subprocess
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np
from scipy import special as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ST_NOT_PD, ST_RATE_NONFINITE, ST_ONETYPE_SINGULAR, ST_PRIOR_SUPPORT = 1, 2, 4, 8


def build(force: bool = False) -> str:
    """Compile oracle/libbporacle.so with the committed Makefile (g++ -O3 -fopenmp)."""
    so = os.path.join(_HERE, "libbporacle.so")
    if force or not os.path.exists(so):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)
        L.bpo_multitype_lp_grad.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, dp, ip, ctypes.c_int, dp, dp, ip, ip,
                                            ctypes.c_int, ctypes.c_int, dp, ctypes.c_int, dp, dp, ctypes.c_int,
                                            ctypes.c_double, ctypes.c_int, dp, dp, dp, ip]
        L.bpo_multitype_lp_grad.restype = ctypes.c_int
        L.bpo_multitype_flops.argtypes = [ctypes.c_int, ctypes.c_int, dp, ip, ctypes.c_int, dp, dp, ip, ip, ctypes.c_int,
                                          ctypes.c_int, dp, dp, ctypes.c_double, ctypes.c_int, ctypes.c_double]
        L.bpo_multitype_flops.restype = ctypes.c_longlong
        L.bpo_onetype_lp_grad.argtypes = [ctypes.c_int, ctypes.c_int, dp, dp, dp, ip, ctypes.c_int, ctypes.c_int, dp,
                                          ctypes.c_int, dp, dp, ip]
        L.bpo_onetype_lp_grad.restype = ctypes.c_int
        L.bpo_onetype_flops.argtypes = [ctypes.c_int, dp, dp, dp, ip, ctypes.c_int, dp]
        L.bpo_onetype_flops.restype = ctypes.c_longlong
        L.bpo_moments.argtypes = [ctypes.c_int, ctypes.c_int, dp, ip, dp, ctypes.c_double, ctypes.c_double, dp, dp]
        L.bpo_moments.restype = ctypes.c_int
        L.bpo_num_threads.restype = ctypes.c_int
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def _f(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _i(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


# ----------------------------------------------------------------------------- rate expressions
# Grammar of exp_to_stan (R/stan_utils.R:205-281): numeric literals, binary + - * / ^, unary + -,
# exp(.), log(.), parentheses, c[k] (1<=k<=nparams), x[k] (1<=k<=ndep).  R precedence: ^ (right
# associative) binds tighter than unary minus, which binds tighter than * /.
_TOK = re.compile(r"\s*(?:(\d+\.?\d*(?:[eE][+-]?\d+)?|\.\d+(?:[eE][+-]?\d+)?)|([A-Za-z_.][A-Za-z0-9_.]*)|(.))")


class ExprError(ValueError):
    pass


def _tokenize(s):
    out, pos = [], 0
    s = s.strip()
    while pos < len(s):
        m = _TOK.match(s, pos)
        if not m:
            raise ExprError("Invalid expression: " + s)
        pos = m.end()
        if m.group(1) is not None:
            out.append(("num", float(m.group(1))))
        elif m.group(2) is not None:
            out.append(("id", m.group(2)))
        else:
            out.append(("op", m.group(3)))
    out.append(("end", None))
    return out


class _Parser:
    def __init__(self, s, nparams, ndep):
        self.src, self.t, self.i, self.np_, self.nd = s, _tokenize(s), 0, nparams, ndep

    def peek(self):
        return self.t[self.i]

    def take(self):
        tok = self.t[self.i]
        self.i += 1
        return tok

    def expect(self, ch):
        k, v = self.take()
        if k != "op" or v != ch:
            raise ExprError("Invalid expression: " + self.src)

    def parse(self):
        e = self.additive()
        if self.peek()[0] != "end":
            raise ExprError("Invalid expression: " + self.src)
        return e

    def additive(self):
        e = self.multiplicative()
        while self.peek() in (("op", "+"), ("op", "-")):
            op = self.take()[1]
            e = (op, e, self.multiplicative())
        return e

    def multiplicative(self):
        e = self.unary()
        while self.peek() in (("op", "*"), ("op", "/")):
            op = self.take()[1]
            e = (op, e, self.unary())
        return e

    def unary(self):
        if self.peek() in (("op", "-"), ("op", "+")):
            op = self.take()[1]
            return ("neg" if op == "-" else "pos", self.unary())
        return self.power()

    def power(self):
        base = self.primary()
        if self.peek() == ("op", "^"):
            self.take()
            return ("^", base, self.unary())   # right associative; exponent may carry a unary sign
        return base

    def primary(self):
        k, v = self.take()
        if k == "num":
            return ("num", v)
        if k == "op" and v == "(":
            e = self.additive()
            self.expect(")")
            return e
        if k == "id":
            if self.peek() == ("op", "("):
                if v not in ("exp", "log"):
                    raise ExprError("Invalid expression: " + self.src)
                self.take()
                arg = self.additive()
                if self.peek() == ("op", ","):
                    raise ExprError("log and exp take only one parameter")
                self.expect(")")
                return (v, arg)
            if self.peek() == ("op", "[") and v in ("c", "x"):
                self.take()
                kk, num = self.take()
                if kk != "num" or num != int(num):
                    raise ExprError("Invalid expression: " + self.src)
                self.expect("]")
                num = int(num)
                if v == "c":
                    if num > self.np_ or num <= 0:
                        raise ExprError("Parameter c[%d] goes beyond the number of parameters specified." % num)
                    return ("c", num - 1)
                if num > self.nd or num <= 0:
                    raise ExprError("Variable x[%d] goes beyond the number of dependent variables specified." % num)
                return ("x", num - 1)
        raise ExprError("Invalid expression: " + self.src)


def parse_rate(s: str, nparams: int, ndep: int):
    return _Parser(s, nparams, ndep).parse()


class Dual:
    """value [C, L] with gradient [C, L, P] w.r.t. the constrained parameters Theta."""

    __slots__ = ("v", "g")

    def __init__(self, v, g):
        self.v, self.g = v, g


def _eval(node, theta, xs):
    C, P = theta.shape
    Ln = xs.shape[0]
    k = node[0]
    if k == "num":
        return Dual(np.full((C, Ln), node[1]), np.zeros((C, Ln, P)))
    if k == "c":
        g = np.zeros((C, Ln, P))
        g[:, :, node[1]] = 1.0
        return Dual(np.repeat(theta[:, node[1]][:, None], Ln, 1), g)
    if k == "x":
        return Dual(np.repeat(xs[:, node[1]][None, :], C, 0), np.zeros((C, Ln, P)))
    if k in ("neg", "pos"):
        a = _eval(node[1], theta, xs)
        return Dual(-a.v, -a.g) if k == "neg" else a
    if k in ("exp", "log"):
        a = _eval(node[1], theta, xs)
        if k == "exp":
            e = np.exp(a.v)
            return Dual(e, a.g * e[..., None])
        with np.errstate(all="ignore"):
            return Dual(np.log(a.v), a.g / a.v[..., None])
    a, b = _eval(node[1], theta, xs), _eval(node[2], theta, xs)
    with np.errstate(all="ignore"):
        if k == "+":
            return Dual(a.v + b.v, a.g + b.g)
        if k == "-":
            return Dual(a.v - b.v, a.g - b.g)
        if k == "*":
            return Dual(a.v * b.v, a.g * b.v[..., None] + b.g * a.v[..., None])
        if k == "/":
            q = a.v / b.v
            return Dual(q, (a.g - b.g * q[..., None]) / b.v[..., None])
        if k == "^":
            p = np.power(a.v, b.v)
            g = a.g * (b.v * np.power(a.v, b.v - 1.0))[..., None]
            if np.any(b.g != 0.0):
                g = g + b.g * (p * np.log(a.v))[..., None]
            return Dual(p, g)
    raise ExprError("bad node")


# ----------------------------------------------------------------------------- priors
# log densities up to additive constants (Stan `~` statement), value and d/dtheta; -inf outside support.
_SQ2, _SQPI = math.sqrt(2.0), math.sqrt(math.pi)


def _log_erfc_and_grad(z):
    """log(erfc(z)) and its derivative w.r.t. z, stable for large |z|."""
    zc = np.where(z > 0, z, 0.0)
    lx = np.where(z > 0, np.log(sp.erfcx(zc)) - zc * zc, np.log(sp.erfc(np.where(z > 0, 0.0, z))))
    d = np.where(z > 0, -2.0 / _SQPI / sp.erfcx(zc), -2.0 / _SQPI * np.exp(-z * z) / sp.erfc(np.where(z > 0, 0.0, z)))
    return lx, d


def prior_logpdf(name: str, p: Sequence[float], y: np.ndarray):
    """returns (lp, dlp/dy, in_support) elementwise."""
    y = np.asarray(y, dtype=np.float64)
    ok = np.ones_like(y, dtype=bool)
    with np.errstate(all="ignore"):
        if name == "normal":
            z = (y - p[0]) / p[1]
            lp, g = -0.5 * z * z, -z / p[1]
        elif name == "exp_mod_normal":
            mu, s, lam = p
            z = (mu + lam * s * s - y) / (_SQ2 * s)
            le, dle = _log_erfc_and_grad(z)
            lp, g = -lam * y + le, -lam + dle * (-1.0 / (_SQ2 * s))
        elif name == "skew_normal":
            xi, om, al = p
            z = (y - xi) / om
            le, dle = _log_erfc_and_grad(-al * z / _SQ2)
            lp, g = -0.5 * z * z + le, -z / om + dle * (-al / (_SQ2 * om))
        elif name == "student_t":
            nu, mu, s = p
            z = (y - mu) / s
            lp, g = -(nu + 1) / 2 * np.log1p(z * z / nu), -(nu + 1) * z / (s * nu * (1 + z * z / nu))
        elif name == "cauchy":
            z = (y - p[0]) / p[1]
            lp, g = -np.log1p(z * z), -2 * z / (p[1] * (1 + z * z))
        elif name == "double_exponential":
            lp, g = -np.abs(y - p[0]) / p[1], -np.sign(y - p[0]) / p[1]
        elif name == "logistic":
            z = (y - p[0]) / p[1]
            lp = -z - 2 * np.log1p(np.exp(-z))
            g = (-1 + 2 * np.exp(-z) / (1 + np.exp(-z))) / p[1]
        elif name == "gumbel":
            z = (y - p[0]) / p[1]
            lp, g = -z - np.exp(-z), (-1 + np.exp(-z)) / p[1]
        elif name == "lognormal":
            ok = y > 0
            ly = np.log(y)
            z = (ly - p[0]) / p[1]
            lp, g = -ly - 0.5 * z * z, (-1 - z / p[1]) / y
        elif name == "chi_square":
            ok = y >= 0
            lp, g = (p[0] / 2 - 1) * np.log(y) - y / 2, (p[0] / 2 - 1) / y - 0.5
        elif name == "inv_chi_square":
            ok = y > 0
            lp, g = -(p[0] / 2 + 1) * np.log(y) - 0.5 / y, -(p[0] / 2 + 1) / y + 0.5 / (y * y)
        elif name == "scaled_inv_chi_square":
            ok = y > 0
            nu, s = p
            lp, g = -(nu / 2 + 1) * np.log(y) - nu * s * s / (2 * y), -(nu / 2 + 1) / y + nu * s * s / (2 * y * y)
        elif name == "exponential":
            ok = y >= 0
            lp, g = -p[0] * y, np.full_like(y, -p[0])
        elif name == "gamma":
            ok = y >= 0
            lp, g = (p[0] - 1) * np.log(y) - p[1] * y, (p[0] - 1) / y - p[1]
        elif name == "inv_gamma":
            ok = y > 0
            lp, g = -(p[0] + 1) * np.log(y) - p[1] / y, -(p[0] + 1) / y + p[1] / (y * y)
        elif name == "weibull":
            ok = y >= 0
            al, s = p
            lp, g = (al - 1) * np.log(y) - np.power(y / s, al), (al - 1) / y - al / s * np.power(y / s, al - 1)
        elif name == "frechet":
            ok = y > 0
            al, s = p
            lp, g = -(al + 1) * np.log(y) - np.power(y / s, -al), -(al + 1) / y + al / s * np.power(y / s, -al - 1)
        elif name == "rayleigh":
            ok = y >= 0
            lp, g = np.log(y) - y * y / (2 * p[0] ** 2), 1 / y - y / p[0] ** 2
        elif name == "pareto":
            ok = y >= p[0]
            lp, g = -(p[1] + 1) * np.log(y), -(p[1] + 1) / y
        elif name == "pareto_type_2":
            mu, lam, al = p
            ok = y >= mu
            lp, g = -(al + 1) * np.log1p((y - mu) / lam), -(al + 1) / (lam + y - mu)
        elif name == "beta":
            ok = (y >= 0) & (y <= 1)
            lp, g = (p[0] - 1) * np.log(y) + (p[1] - 1) * np.log1p(-y), (p[0] - 1) / y - (p[1] - 1) / (1 - y)
        elif name == "uniform":
            ok = (y >= p[0]) & (y <= p[1])
            lp, g = np.zeros_like(y), np.zeros_like(y)
        else:
            raise ValueError("Invalid prior name!")
    ok = ok & np.isfinite(lp)
    return np.where(ok, lp, -np.inf), np.where(ok, g, 0.0), ok


# ----------------------------------------------------------------------------- model container
@dataclass
class OracleModel:
    kind: str                       # 'multitype' | 'multitype_noise' | 'simple_bd'
    e_mat: np.ndarray               # E x n
    p_vec: np.ndarray               # E, 1-based
    func_deps: List[str]
    nparams: int
    ndep: int
    priors: List[dict]              # {name, params, bounds or None}
    _ast: list = field(default_factory=list)

    def __post_init__(self):
        self.e_mat = np.asarray(self.e_mat, dtype=np.float64)
        if self.e_mat.ndim == 1:
            self.e_mat = self.e_mat[:, None]
        self.p_vec = np.asarray(self.p_vec, dtype=np.int32)
        self._ast = [parse_rate(s, self.nparams, self.ndep) for s in self.func_deps]

    @property
    def dim(self):
        return self.nparams + (1 if self.kind == "multitype_noise" else 0)


def has_bounds(b) -> bool:
    """R passes bounds = NA for "no bounds" (R/stan_utils.R:397)."""
    if b is None:
        return False
    b = np.asarray(b, dtype=np.float64)
    return b.size == 2 and not np.any(np.isnan(b))


def constrain(model: OracleModel, u: np.ndarray):
    """u [C, dim] -> theta [C, P], sigma_obs [C] or None, log-Jacobian [C], dtheta/du [C, dim], dlogJ/du [C, dim]."""
    C = u.shape[0]
    P = model.nparams
    theta = np.empty((C, P))
    dth = np.ones((C, model.dim))
    lj = np.zeros(C)
    dlj = np.zeros((C, model.dim))
    for k, pr in enumerate(model.priors):
        b = pr.get("bounds")
        if not has_bounds(b):
            theta[:, k] = u[:, k]
            continue
        lo, hi = float(b[0]), float(b[1])
        s = sp.expit(u[:, k])
        theta[:, k] = lo + (hi - lo) * s
        dth[:, k] = (hi - lo) * s * (1 - s)
        au = np.abs(u[:, k])
        lj += math.log(hi - lo) - au - 2 * np.log1p(np.exp(-au))
        dlj[:, k] = 1 - 2 * s
    sig = None
    if model.kind == "multitype_noise":
        sig = np.exp(u[:, P])
        dth[:, P] = sig
        lj += u[:, P]
        dlj[:, P] = 1.0
    return theta, sig, lj, dth, dlj


def rates(model: OracleModel, theta: np.ndarray, function_var: np.ndarray):
    """r [C, L, E] and dr/dtheta [C, L, E, P]."""
    xs = np.asarray(function_var, dtype=np.float64)
    if xs.ndim == 1:
        xs = xs[:, None]
    duals = [_eval(a, theta, xs) for a in model._ast]
    r = np.stack([d.v for d in duals], axis=2)
    J = np.stack([d.g for d in duals], axis=2)
    return r, J


def log_prob(model: OracleModel, data: dict, upars: np.ndarray, jacobian: bool = True, grouped: Optional[bool] = None,
             precision: int = 0, theta_step: float = 1.0, nthreads: int = 0):
    """lp [C], grad [C, dim], status [C] at unconstrained parameter vectors upars [C, dim]
    (declaration order Theta1..ThetaP then sigma_obs, multitype_noise_template.stan:99-104)."""
    u = np.atleast_2d(_f(upars))
    C = u.shape[0]
    P = model.nparams
    theta, sig, lj, dth, dlj = constrain(model, u)
    r, J = rates(model, theta, data["function_var"])
    Ln, E = r.shape[1], r.shape[2]
    status = np.zeros(C, dtype=np.int32)
    lp = np.zeros(C)
    rbar = np.zeros((C, Ln, E))
    sbar = np.zeros(C)
    L = lib()
    rc = _f(r)
    bad = ~np.isfinite(rc).all(axis=(1, 2))
    rc = np.where(np.isfinite(rc), rc, 0.0)
    if model.kind == "simple_bd":
        N = int(data["ndatapts"])
        pv, ip, tm, vi = _f(data["pop_vec"]).ravel(), _f(data["init_pop"]).ravel(), _f(data["times"]).ravel(), _i(data["var_idx"])
        L.bpo_onetype_lp_grad(precision, N, _dp(pv), _dp(ip), _dp(tm), _ip(vi), Ln, C, _dp(rc), nthreads, _dp(lp), _dp(rbar),
                              _ip(status))
    else:
        n, N = int(data["ntypes"]), int(data["ndatapts"])
        em = np.asfortranarray(_f(model.e_mat))
        pv = np.asfortranarray(_f(data["pop_vec"]).reshape(N, n))
        ip = np.asfortranarray(_f(data["init_pop"]).reshape(N, n))
        tm, vi, ti = _f(np.atleast_1d(data["times"])), _i(data["var_idx"]), _i(data["times_idx"])
        Tn = tm.shape[0]
        if grouped is None:
            grouped = len(set(zip(vi.tolist(), ti.tolist()))) * 2 * n <= N
        sg = _f(sig) if sig is not None else None
        rcode = L.bpo_multitype_lp_grad(precision, n, E, _dp(em.ravel(order="F")), _ip(_i(model.p_vec)), N,
                                        _dp(pv.ravel(order="F")), _dp(ip.ravel(order="F")), _ip(vi), _ip(ti), Ln, Tn, _dp(tm),
                                        C, _dp(rc), _dp(sg) if sg is not None else None, int(bool(grouped)), theta_step,
                                        nthreads, _dp(lp), _dp(rbar), _dp(sbar), _ip(status))
        if rcode:
            raise ValueError("unsupported ntypes")
    status[bad] |= ST_RATE_NONFINITE
    # priors on the constrained scale
    gth = np.einsum("cle,clep->cp", rbar, J)
    for k, pr in enumerate(model.priors):
        pl, pg, ok = prior_logpdf(pr["name"], pr["params"], theta[:, k])
        lp += pl
        gth[:, k] += pg
        status[~ok] |= ST_PRIOR_SUPPORT
    grad = np.zeros((C, model.dim))
    grad[:, :P] = gth * dth[:, :P]
    if sig is not None:
        # sigma_obs ~ normal(0, .1) on sigma_obs > 0   (multitype_noise_template.stan:131)
        lp += -0.5 * (sig / 0.1) ** 2
        grad[:, P] = (sbar - sig / 0.01) * dth[:, P]
    if jacobian:
        lp += lj
        grad += dlj
    fail = status != 0
    lp[fail] = -np.inf
    grad[fail] = 0.0
    return lp, grad, status


def moments(e_mat, p_vec, r, t, theta_step: float = 1.0):
    """single-ancestor mean matrix M [n, n] and covariances V [n, n, n]."""
    em = np.asarray(e_mat, dtype=np.float64)
    if em.ndim == 1:
        em = em[:, None]
    E, n = em.shape
    M, V = np.zeros((n, n)), np.zeros((n, n, n))
    rc = lib().bpo_moments(n, E, _dp(np.asfortranarray(em).ravel(order="F")), _ip(_i(p_vec)), _dp(_f(r)), float(t), theta_step,
                           _dp(M), _dp(V))
    if rc:
        raise ValueError("bpo_moments failed (%d)" % rc)
    return M, V


def mu_sigma(e_mat, p_vec, r, z0, t):
    """what R/calculate_moments.R:79-86 returns: mu = M^T z0, Sigma = sum_l z0_l V_l."""
    M, V = moments(e_mat, p_vec, r, t)
    z0 = np.asarray(z0, dtype=np.float64)
    return M.T @ z0, np.einsum("l,lij->ij", z0, V)



def log_lik(model: OracleModel, data: dict, upars: np.ndarray) -> np.ndarray:
    """per-observation log_lik [C, ndatapts] in the order of the data list, WITH the normalising constants: the
    generated-quantities blocks stanfiles/multitype_loo_block.stan:15-32 (`multi_normal_lpdf(pop_vec[k] | mu_t, sigma_t)`) and
    simple_birth_death_loo_block.stan:3-12 (`normal_lpdf(pop_vec[k] | mu, sigma)`), which generate(..., loo = T) appends
    (R/stan_utils.R:173-175,189-191).  Literal restatement, one observation at a time, SciPy densities.

    Reference quirk kept: generate() appends the SAME multitype LOO block to the noise template (R/stan_utils.R:189-191), and that
    block builds sigma_t without the sigma_obs * diag(mu_t) term of multitype_noise_template.stan:153 — so log_lik of a noise model
    is the noise-free multi_normal density."""
    from scipy import stats
    u = np.atleast_2d(_f(upars))
    C = u.shape[0]
    theta, sig, *_ = constrain(model, u)
    r, _ = rates(model, theta, data["function_var"])
    N = int(data["ndatapts"])
    vi = _i(data["var_idx"]) - 1
    out = np.full((C, N), np.nan)
    if model.kind == "simple_bd":
        pv, ip, tm = _f(data["pop_vec"]).ravel(), _f(data["init_pop"]).ravel(), _f(data["times"]).ravel()
        for c in range(C):
            b, d = r[c, vi, 0], r[c, vi, 1]
            with np.errstate(all="ignore"):
                e = np.exp(tm * (b - d))
                mu = e * ip
                sd = np.sqrt((e * (b * (2 * e - 1) - d) / (b - d) - np.exp(2 * tm * (b - d))) * ip)
                out[c] = stats.norm.logpdf(pv, mu, sd)
        return out
    n = int(data["ntypes"])
    pv, ip = _f(data["pop_vec"]).reshape(N, n), _f(data["init_pop"]).reshape(N, n)
    tm, ti = _f(np.atleast_1d(data["times"])), _i(data["times_idx"]) - 1
    for c in range(C):
        cache = {}
        for k in range(N):
            key = (int(vi[k]), int(ti[k]))
            if key not in cache:
                cache[key] = moments(model.e_mat, model.p_vec, r[c, key[0]], float(tm[key[1]]))
            M, V = cache[key]
            mu = M.T @ ip[k]                                  # multitype_loo_block.stan:22
            S = np.einsum("l,lij->ij", ip[k], V)              # :23-28
            try:
                out[c, k] = stats.multivariate_normal.logpdf(pv[k], mu, 0.5 * (S + S.T), allow_singular=False)   # :31
            except Exception:
                out[c, k] = np.nan
    return out


def transformed_parameters(model: OracleModel, data: dict, upars: np.ndarray):
    """the transformed-parameters block as rstan stores it after the loop over dependent-variable levels
    (multitype_template.stan:105-117, one_type_template.stan:17-28): `r_mat` keeps the values of the LAST level (it is
    re-initialised at the top of every iteration), column-major flattening; `theta[i]` = append_array(to_array_1d(e_mat),
    to_array_1d(r_mat)) per level (to_array_1d of a matrix is column-major), flattened by rstan as theta[1,1], theta[2,1], ...
    (first index fastest).  Returns (names, values [C, ncols]).  One-type template: r_mat[2, ndep_levels] only."""
    u = np.atleast_2d(_f(upars))
    C = u.shape[0]
    th, *_ = constrain(model, u)
    r, _ = rates(model, th, data["function_var"])          # [C, L, E]
    Ln, E = r.shape[1], r.shape[2]
    if model.kind == "simple_bd":
        names = ["r_mat[%d,%d]" % (e + 1, v + 1) for v in range(Ln) for e in range(2)]
        vals = np.stack([r[:, v, e] for v in range(Ln) for e in range(2)], axis=1)
        return names, vals
    n = model.e_mat.shape[1]
    pz = np.asarray(model.p_vec, dtype=int) - 1

    def rmat(level):
        m = np.zeros((C, E, n))
        for e in range(E):
            m[:, e, pz[e]] = r[:, level, e]
        return m
    last = rmat(Ln - 1)
    names = ["r_mat[%d,%d]" % (e + 1, j + 1) for j in range(n) for e in range(E)]
    cols = [last[:, e, j] for j in range(n) for e in range(E)]
    tl = []
    for v in range(Ln):
        m = rmat(v)
        ev = np.broadcast_to(model.e_mat.ravel(order="F")[None, :], (C, E * n))
        tl.append(np.concatenate([ev, m.transpose(0, 2, 1).reshape(C, E * n)], axis=1))   # [C, 2 E n]
    for q in range(2 * E * n):
        for v in range(Ln):
            names.append("theta[%d,%d]" % (v + 1, q + 1))
            cols.append(tl[v][:, q])
    return names, np.stack(cols, axis=1)


def flops_per_chain(model: OracleModel, data: dict, theta: np.ndarray, sigma_obs: float = -1.0, grouped: bool = True,
                    theta_step: float = 1.0) -> int:
    """exact operation count of one chain's likelihood+adjoint for the shipped algorithm."""
    r, _ = rates(model, np.atleast_2d(_f(theta)), data["function_var"])
    rc = _f(r[0])
    L = lib()
    if model.kind == "simple_bd":
        N = int(data["ndatapts"])
        return int(L.bpo_onetype_flops(N, _dp(_f(data["pop_vec"]).ravel()), _dp(_f(data["init_pop"]).ravel()),
                                       _dp(_f(data["times"]).ravel()), _ip(_i(data["var_idx"])), rc.shape[0], _dp(rc)))
    n, N = int(data["ntypes"]), int(data["ndatapts"])
    em = np.asfortranarray(_f(model.e_mat))
    pv = np.asfortranarray(_f(data["pop_vec"]).reshape(N, n))
    ip = np.asfortranarray(_f(data["init_pop"]).reshape(N, n))
    tm = _f(np.atleast_1d(data["times"]))
    return int(L.bpo_multitype_flops(n, rc.shape[1], _dp(em.ravel(order="F")), _ip(_i(model.p_vec)), N, _dp(pv.ravel(order="F")),
                                     _dp(ip.ravel(order="F")), _ip(_i(data["var_idx"])), _ip(_i(data["times_idx"])), rc.shape[0],
                                     tm.shape[0], _dp(tm), _dp(rc), float(sigma_obs), int(grouped), theta_step))


def num_threads() -> int:
    return int(lib().bpo_num_threads())
