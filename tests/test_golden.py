"""Committed golden vectors (tests/golden/lp_grad_golden.json, made by tests/golden/make_golden.py): the CPU oracle
reproduces them (-m "not gpu") and the CUDA path matches them (-m gpu), without reading anything outside the repo."""
import json
import os

import numpy as np
import pytest

import bpinference_b200 as bp
from oracle import bp_oracle as bo

G = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lp_grad_golden.json")))


def _data(case):
    d = dict(case["data"])
    for k in ("pop_vec", "init_pop", "times", "function_var", "e_mat"):
        if k in d:
            d[k] = np.asarray(d[k], dtype=float)
    for k in ("var_idx", "times_idx", "p_vec"):
        if k in d:
            d[k] = np.asarray(d[k], dtype=np.int32)
    return d


@pytest.mark.parametrize("case", G["cases"], ids=[c["name"] for c in G["cases"]])
def test_oracle_reproduces_golden(case):
    om = bo.OracleModel(case["kind"], case["e_mat"], case["p_vec"], case["func_deps"], case["nparams"], case["ndep"], case["priors"])
    for jac in (True, False):
        exp = case["expected"]["jacobian_%d" % jac]
        lp, g, st = bo.log_prob(om, _data(case), np.asarray(case["upars"]), jacobian=jac)
        assert st.tolist() == exp["status"]
        np.testing.assert_allclose(lp, exp["lp"], rtol=1e-12)
        np.testing.assert_allclose(g, exp["grad"], rtol=1e-10, atol=1e-11 * np.abs(exp["grad"]).max())


def test_oracle_reproduces_vignette_values():
    v = G["vignettes"]["two_type"]
    mu, S = bo.mu_sigma(v["e_mat"], v["p_vec"], v["rates"], v["z0"], v["t"])
    np.testing.assert_allclose(mu, v["mean_printed"], rtol=1e-14)
    np.testing.assert_allclose(S.ravel(), v["cov_printed_desolve_1e-6"], rtol=4e-5)
    v = G["vignettes"]["one_type"]
    mu, S = bo.mu_sigma([[2.0], [0.0]], [1, 1], v["rates"], v["z0"], v["t"])
    np.testing.assert_allclose(mu, v["mean_printed"], rtol=1e-14)
    np.testing.assert_allclose(S.ravel(), v["var_printed_desolve_1e-6"], rtol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("case", G["cases"], ids=[c["name"] for c in G["cases"]])
def test_cuda_matches_golden(case):
    mod = bp.bp_model(np.asarray(case["e_mat"], dtype=float), case["p_vec"], case["func_deps"], case["nparams"], case["ndep"])
    eng = bp.stan_model(mod, case["priors"], simple_bd=case["kind"] == "simple_bd", noise_model=case["kind"] == "multitype_noise")
    for grouping in ((0,) if case["kind"] == "simple_bd" else (1, 2)):
        d = eng.data(_data(case), grouping=grouping)
        for jac in (True, False):
            exp = case["expected"]["jacobian_%d" % jac]
            lp, g, st = eng.log_prob_grad(d, np.asarray(case["upars"]), adjust_transform=jac)
            assert st.tolist() == exp["status"]
            np.testing.assert_allclose(lp, exp["lp"], rtol=1e-11)
            np.testing.assert_allclose(g, exp["grad"], rtol=1e-9, atol=1e-10 * np.abs(exp["grad"]).max())
