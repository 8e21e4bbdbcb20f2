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
