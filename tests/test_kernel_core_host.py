"""The per-thread core of the CUDA kernels (the BP_HD device functions of bp_kernels.cuh plus the generated rate /
assembly code), compiled by g++ and run on the host (tests/hostsim), against the CPU oracle: this pins the kernel
ARITHMETIC without a GPU.  The -m gpu tests then pin the kernels themselves."""
import numpy as np
import pytest

import bpinference_b200 as bp
from bpinference_b200 import synthetic as syn
from oracle import bp_oracle as bo
from tests import hostsim_util as hs


def _setup(name, nobs):
    cfg = syn.make_config(name, chains=2, nobs=nobs)
    mod = bp.bp_model(cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"])
    eng = bp.stan_model(mod, cfg["priors"], simple_bd=cfg["kind"] == "simple_bd", noise_model=cfg["kind"] == "multitype_noise")
    om = bo.OracleModel(cfg["kind"], cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"], cfg["priors"])
    return cfg, eng, om


@pytest.mark.parametrize("name,nobs", [("cfg1", None), ("cfg2", 60), ("cfg3", None), ("cfg4", 60), ("cfg5U", 40), ("cfg5K", 40), ("cfg5G", 40)])
def test_kernel_core_matches_oracle(name, nobs):
    cfg, eng, om = _setup(name, nobs)
    u = cfg["upars"]
    theta, sig, lj, dth, dlj = bo.constrain(om, u)
    lp_o, g_o, st_o = bo.log_prob(om, cfg["data"], u, jacobian=False)
    P = om.nparams
    for c in range(2):
        th = np.append(theta[c], sig[c]) if sig is not None else theta[c]
        lp, cb, sb, st, apps, _ = hs.likelihood(eng, cfg["data"], th)
        pri = [bo.prior_logpdf(pr["name"], pr["params"], theta[c:c + 1, k]) for k, pr in enumerate(om.priors)]
        prior = sum(p[0][0] for p in pri) - (0.5 * (sig[c] / 0.1) ** 2 if sig is not None else 0.0)
        assert st == 0 and st_o[c] == 0
        assert abs(lp - (lp_o[c] - prior)) <= 1e-12 * abs(lp_o[c])
        gth = (cb + np.array([p[1][0] for p in pri])) * dth[c, :P]
        np.testing.assert_allclose(gth, g_o[c, :P], rtol=1e-10, atol=1e-12 * np.abs(g_o[c]).max())
        if sig is not None:
            assert abs((sb - sig[c] / 0.01) * sig[c] - g_o[c, P]) <= 1e-10 * abs(g_o[c, P])


def test_substeps_and_recompute_path():
    """rates x duration large enough that the series needs several sub-steps (the recompute branch of the reverse sweep)."""
    cfg, eng, om = _setup("cfg2", 12)
    d = dict(cfg["data"])
    d["times"] = np.array([9.0])
    u = cfg["upars"] + 1.5            # rates around 1: ||tL|| ~ 40 -> ~20 sub-steps
    theta, _, _, dth, _ = bo.constrain(om, u)
    lp_o, g_o, st_o = bo.log_prob(om, d, u, jacobian=False)
    lp, cb, sb, st, apps, _ = hs.likelihood(eng, d, theta[0])
    assert apps / 12 >= 75            # several sub-steps of up to 25 applications each
    pri = [bo.prior_logpdf(pr["name"], pr["params"], theta[0:1, k]) for k, pr in enumerate(om.priors)]
    if st_o[0] == 0:
        assert abs(lp - (lp_o[0] - sum(p[0][0] for p in pri))) <= 1e-10 * abs(lp_o[0])
        gth = (cb + np.array([p[1][0] for p in pri])) * dth[0]
        np.testing.assert_allclose(gth, g_o[0], rtol=1e-8, atol=1e-10 * np.abs(g_o[0]).max())
    else:
        assert st != 0


@pytest.mark.parametrize("name", ["cfg2", "cfg5U", "cfg4", "cfg1"])
def test_flop_model_equals_counted_operations(name):
    """bpc_flop_model (the constants behind bench.py's roofline.achieved) against a recount of the same device
    functions with a counting scalar: within 1% when the series fits the shared-memory slots (short durations), and
    never above the recount (recomputed lower terms are executed but are not algorithmic work) otherwise."""
    cfg, eng, om = _setup(name, 24 if name != "cfg1" else None)
    theta, sig, *_ = bo.constrain(om, cfg["upars"])
    th = np.append(theta[0], sig[0]) if sig is not None else theta[0]
    fm = eng.flop_model()
    N = cfg["data"]["ndatapts"]
    for scale, exact in ((0.05, True), (1.0, False)):
        d = dict(cfg["data"])
        d["times"] = np.asarray(d["times"], dtype=float) * scale
        lp, cb, sb, st, apps, ops = hs.likelihood(eng, d, th, counted=True)
        model = apps * (fm["per_app_fwd"] + fm["per_app_bwd"]) + N * fm["per_obs"]
        counted = float(ops[1])
        assert counted > 0
        if exact or eng.kind == 2:
            assert abs(model - counted) <= 0.01 * counted, (model, counted)
        else:
            assert model <= 1.01 * counted and counted <= 1.25 * model, (model, counted)


@pytest.mark.parametrize("name", ["cfg5U", "cfg2", "cfg4"])
def test_octet_flop_constants_and_adjugate_score(name):
    """the constants behind roofline.achieved on the octet path (per term of the short series, per observation outside the terms:
    csrc/flops.cpp per_term_c / per_obs_o) against a recount of the octet formula with a counting scalar (tests/hostsim:
    hs_centred_obs performs the arithmetic of one (chain, observation) of bp_fusedo in scalar form), EXACTLY; the formula's numbers
    against the direct series; and the adjugate score of bp_fusedo against the LDL^T score of the other kernels."""
    import ctypes as C
    cfg, eng, om = _setup(name, 24)
    fm = eng.flop_model()
    n = eng.ntypes
    S = n * (n + 1) // 2
    D = n + S
    dp = C.POINTER(C.c_double)
    theta, sig, *_ = bo.constrain(om, cfg["upars"])
    r, _ = bo.rates(om, theta, cfg["data"]["function_var"])
    e_mat, p_vec = np.asarray(cfg["e_mat"], float), np.asarray(cfg["p_vec"])
    # N_j = L^j exp(tau L) [e_1 .. e_n] from the oracle's moments: column l of N_0 = state from one type-l ancestor; N_j by differences of
    # the exact states is ill-conditioned, so build L explicitly from unit-vector applications of the ODE right-hand side (refmath)
    from tests import refmath
    Lmat = refmath.moment_operator(e_mat, p_vec, r[0, 0])           # [D, D] packed coordinates
    import scipy.linalg as sl
    tau, h = 1.3, 0.011
    E0 = sl.expm(tau * Lmat)[:, :n]
    for J in (7, 6):
        Nj = np.stack([np.linalg.matrix_power(Lmat, j) @ E0 for j in range(J + 1)])      # [J+1, D, n]
        z0 = np.asarray(cfg["data"]["init_pop"], float).reshape(-1, n)[0]
        y_exact = sl.expm((tau + h) * Lmat)[:, :n] @ z0
        xo = y_exact[:n] + 0.7 * np.sqrt(np.diag(refmath.unpack_sym(y_exact[n:], n)))
        sg = float(sig[0]) if sig is not None else 0.0
        for counted in (False, True):
            Lh = hs.build(eng, counted=counted)
            y, b, sp = np.zeros(D), np.zeros(D), np.zeros((J + 1, n, D))
            det, q, ops = C.c_double(), C.c_double(), (C.c_ulonglong * 2)()
            rc = Lh.hs_centred_obs(J, np.ascontiguousarray(Nj).ctypes.data_as(dp), z0.ctypes.data_as(dp), np.ascontiguousarray(xo).ctypes.data_as(dp),
                                   C.c_double(h), C.c_double(sg), y.ctypes.data_as(dp), C.byref(det), C.byref(q), b.ctypes.data_as(dp),
                                   sp.ctypes.data_as(dp), ops)
            assert rc == 0
            if counted:
                assert ops[0] == (J + 1) * fm["per_term_c"] + fm["per_obs_o"], (ops[0], fm)
            else:
                np.testing.assert_allclose(y, y_exact, rtol=1e-12)         # the short series around tau reproduces exp((tau + h) L) z0
                # S' rows are X[(j, l)] b: the W-matrix factorisation's input
                cj = np.cumprod(np.concatenate([[1.0], h / np.arange(1, J + 1)]))
                np.testing.assert_allclose(sp, cj[:, None, None] * z0[None, :, None] * b[None, None, :], rtol=1e-13)
                # both scores on the same state
                mu, Sg = y_exact[:n].copy(), y_exact[n:].copy()
                if sig is not None:
                    for i in range(n):
                        Sg[refmath.sidx(i, i, n)] += sg * mu[i]
                lp1, b1, b2 = C.c_double(), np.zeros(D), np.zeros(D)
                rc2 = Lh.hs_score_both(np.ascontiguousarray(xo).ctypes.data_as(dp), mu.ctypes.data_as(dp), Sg.ctypes.data_as(dp), C.byref(lp1),
                                       b1.ctypes.data_as(dp), C.byref(det), C.byref(q), b2.ctypes.data_as(dp))
                assert rc2 == 0
                np.testing.assert_allclose(-0.5 * (np.log(det.value) + q.value), lp1.value, rtol=1e-13)
                np.testing.assert_allclose(b2, b1, rtol=1e-11, atol=1e-13 * np.abs(b1).max())


@pytest.mark.parametrize("n", [1, 2, 3])
def test_s_prime_block_layout_is_a_bijection_and_contiguous_for_bp_toep(n, tmp_path):
    """bp_osg_off (the position of S'[(j, l)][column] inside a group's block, written by bp_fusedo and read by bp_toep): the function's own
    text from csrc/bp_kernels.cuh compiled for the host.  Every element of the 8 n x 8 D block has its own slot, the 8 x 8 elements
    (all j, the 8 columns of one tile) of one ancestor l are 64 consecutive doubles and the n ancestors of a tile follow each other --
    what makes a CTA of bp_toep read n x 512 contiguous bytes per group."""
    import ctypes
    import os
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "bpinference_b200", "csrc", "bp_kernels.cuh")).read()
    m = re.search(r"__device__ __forceinline__ int bp_osg_off\(int row, int col\) \{.*?\n\}\n", text, re.S)
    assert m, "bp_osg_off not found"
    src = tmp_path / "osg.cpp"
    src.write_text("#define __device__\n#define __forceinline__ inline\n#define BP_N %d\n#define BP_CJ 8\n%s\nextern \"C\" int osg(int r, int c) { return bp_osg_off(r, c); }\n"
                   % (n, m.group(0)))
    lib = tmp_path / "osg.so"
    subprocess.run(["g++", "-O1", "-shared", "-fPIC", str(src), "-o", str(lib)], check=True)
    f = ctypes.CDLL(str(lib)).osg
    D = n + n * (n + 1) // 2
    rows, cols = 8 * n, 8 * D
    off = np.array([[f(r, c) for c in range(cols)] for r in range(rows)])
    assert sorted(off.ravel().tolist()) == list(range(rows * cols))
    for ct in range(D):
        for l in range(n):
            blk = np.array([[off[j * n + l, ct * 8 + c8] for c8 in range(8)] for j in range(8)])
            base = (ct * n + l) * 64
            assert (blk == base + np.arange(64).reshape(8, 8)).all()
