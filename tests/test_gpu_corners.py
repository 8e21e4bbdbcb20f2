"""GPU parity of the corners round 1 left untested: every one of the 22 prior kernels on the device, unbounded
declarations through log_prob / grad_log_prob, per-observation log_lik element by element, and the octet path against the
oracle at the bench shape.  Everything goes through the C ABI (ctypes mirror); the CPU oracle is the checker."""
import numpy as np
import pytest

import bpinference_b200 as bp
from bpinference_b200 import synthetic as syn
from oracle import bp_oracle as bo

pytestmark = pytest.mark.gpu

# name, params (valid for R/stan_utils.R:418-459), bounds inside the support
ALL_PRIORS = [
    ("normal", [0.1, 0.25], (0.0, 2.0)), ("exp_mod_normal", [0.1, 0.2, 3.0], (0.0, 2.0)), ("skew_normal", [0.3, 0.5, -2.0], (-1.0, 3.0)),
    ("student_t", [4.0, 0.2, 0.7], (-3.0, 3.0)), ("cauchy", [0.1, 0.6], (-2.0, 4.0)), ("double_exponential", [0.4, 0.3], (-1.0, 2.0)),
    ("logistic", [0.2, 0.5], (-2.0, 2.0)), ("gumbel", [0.1, 0.4], (-1.0, 3.0)), ("lognormal", [-1.0, 0.7], (0.0, 4.0)),
    ("chi_square", [3.0], (0.0, 9.0)), ("inv_chi_square", [2.5], (0.0, 5.0)), ("scaled_inv_chi_square", [3.0, 0.6], (0.0, 5.0)),
    ("exponential", [2.0], (0.0, 6.0)), ("gamma", [2.5, 3.0], (0.0, 5.0)), ("inv_gamma", [3.0, 1.5], (0.0, 5.0)),
    ("weibull", [1.7, 0.9], (0.0, 4.0)), ("frechet", [2.2, 0.8], (0.0, 6.0)), ("rayleigh", [0.7], (0.0, 4.0)),
    ("pareto", [0.5, 2.0], (0.5, 7.0)), ("pareto_type_2", [0.1, 0.8, 2.5], (0.1, 6.0)), ("beta", [2.0, 3.5], (0.0, 1.0)),
    ("uniform", [0.0, 1.5], (0.0, 1.5)),
]


def _one_type_with_priors(priors):
    """one-type template, birth = c[1], death = c[2]; the remaining parameters only carry a prior and a transform."""
    P = len(priors)
    cfg = syn.make_config("cfg1", chains=1)
    mod = bp.bp_model_simple_birth_death(["c[1]", "c[2]"], P, 0)
    eng = bp.stan_model(mod, priors, simple_bd=True)
    om = bo.OracleModel("simple_bd", [[2.0], [0.0]], [1, 1], ["c[1]", "c[2]"], P, 0, priors)
    return cfg["data"], eng, om


@pytest.mark.parametrize("jacobian", [True, False])
def test_every_prior_kernel_bounded(jacobian):
    """all 22 ALLOWED_DISTRIBUTIONS (R/sysdata.rda, emitted as `ThetaK ~ name(params)` at R/stan_utils.R:395) in ONE model, each on
    its own bounded parameter: value and gradient of bpc_log_prob against the oracle's prior_logpdf + bound transform."""
    priors = [bp.prior_dist(nm, par, list(b)) for nm, par, b in ALL_PRIORS]
    # parameters 1, 2 are the rates: normal / exp_mod_normal on (0, 2)
    data, eng, om = _one_type_with_priors(priors)
    rng = np.random.default_rng(5)
    u = rng.normal(0.0, 1.2, size=(24, len(priors)))
    u[:, 0] = np.log(0.1 / 1.9) + 0.2 * rng.standard_normal(24)
    u[:, 1] = np.log(0.05 / 1.95) + 0.2 * rng.standard_normal(24)
    lp, g, st = eng.log_prob_grad(data, u, adjust_transform=jacobian)
    lp_o, g_o, st_o = bo.log_prob(om, data, u, jacobian=jacobian)
    assert st.tolist() == st_o.tolist() and (st == 0).all()
    np.testing.assert_allclose(lp, lp_o, rtol=1e-11)
    np.testing.assert_allclose(g, g_o, rtol=1e-9, atol=1e-10 * np.abs(g_o).max())
    # each kernel on its own: only parameter k moves -> the change of lp is that prior's (plus its Jacobian term)
    base = u[:1].copy()
    for k in range(2, len(priors)):
        v = np.repeat(base, 5, axis=0)
        v[:, k] = np.linspace(-2.5, 2.5, 5)
        a, ga, _ = eng.log_prob_grad(data, v, adjust_transform=jacobian)
        b, gb, _ = bo.log_prob(om, data, v, jacobian=jacobian)
        np.testing.assert_allclose(a - a[0], b - b[0], rtol=1e-9, atol=1e-10, err_msg=ALL_PRIORS[k][0])
        np.testing.assert_allclose(ga[:, k], gb[:, k], rtol=1e-9, atol=1e-11, err_msg=ALL_PRIORS[k][0])


def test_every_prior_kernel_unbounded_and_support():
    """bounds = NA (`real ThetaK;`, R/stan_utils.R:397-410): Theta = u, no Jacobian term; a value outside the prior's support
    rejects the chain with BPC_ST_PRIOR_SUPPORT exactly where the oracle does."""
    priors = [bp.prior_dist(nm, par) for nm, par, _ in ALL_PRIORS]
    priors[0] = bp.prior_dist("normal", [0.1, 0.25])          # birth, unbounded
    priors[1] = bp.prior_dist("exp_mod_normal", [0.1, 0.2, 3.0])   # death, unbounded
    data, eng, om = _one_type_with_priors(priors)
    P = len(priors)
    rng = np.random.default_rng(6)
    inside = np.array([0.5 * (b[0] + b[1]) for _, _, b in ALL_PRIORS])
    u = inside[None, :] + 0.2 * rng.uniform(-1, 1, size=(8 + P, P)) * np.array([min(1.0, b[1] - b[0]) for _, _, b in ALL_PRIORS])[None, :]
    u[:, 0] = 0.1 + 0.01 * rng.standard_normal(len(u))
    u[:, 1] = 0.05 + 0.005 * rng.standard_normal(len(u))
    for k in range(2, P):                                      # chain 8 + k: parameter k far below its range
        u[8 + k, k] = -3.0
    lp, g, st = eng.log_prob_grad(data, u)
    lp_o, g_o, st_o = bo.log_prob(om, data, u)
    assert st.tolist() == st_o.tolist()
    assert (st[:8] == 0).all()
    positive = {"lognormal", "chi_square", "inv_chi_square", "scaled_inv_chi_square", "exponential", "gamma", "inv_gamma", "weibull",
                "frechet", "rayleigh", "pareto", "pareto_type_2", "beta", "uniform"}
    for k in range(2, P):
        assert bool(st[8 + k] & 8) == (ALL_PRIORS[k][0] in positive), ALL_PRIORS[k][0]
    ok = st == 0
    np.testing.assert_allclose(lp[ok], lp_o[ok], rtol=1e-11)
    np.testing.assert_allclose(g[ok], g_o[ok], rtol=1e-9, atol=1e-10 * np.abs(g_o[ok]).max())
    assert np.isneginf(lp[~ok]).all() and (g[~ok] == 0).all()
    lpn, gn, _ = eng.log_prob_grad(data, u, adjust_transform=False)   # nothing to adjust: identical
    assert (lpn[ok] == lp[ok]).all() and (gn[ok] == g[ok]).all()


@pytest.mark.parametrize("name,nobs,chains,grouping", [("cfg2", None, 20, 0), ("cfg2", None, 20, 1), ("cfg5U", 4000, 16, 0), ("cfg4", None, 4, 0)])
def test_unbounded_declarations_through_multitype_log_prob(name, nobs, chains, grouping):
    """the multitype kernels (grouped, octet, noise) with `real ThetaK;` declarations: lp and gradient against the oracle."""
    cfg = syn.make_config(name, chains=chains, nobs=nobs)
    priors = [dict(name="normal", params=[0.1, 0.25], bounds=None) for _ in cfg["priors"]]
    priors[-1] = dict(name="gamma", params=[2.0, 8.0], bounds=None)
    mod = bp.bp_model(cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"])
    eng = bp.stan_model(mod, priors, noise_model=cfg["kind"] == "multitype_noise")
    om = bo.OracleModel(cfg["kind"], cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"], priors)
    rng = np.random.default_rng(2)
    u = np.tile(cfg["theta_true"], (chains, 1)) * (1 + 0.05 * rng.standard_normal((chains, cfg["nparams"])))
    if cfg["kind"] == "multitype_noise":
        u = np.concatenate([u, np.log(0.05) + 0.1 * rng.standard_normal((chains, 1))], axis=1)
    d = eng.data(cfg["data"], grouping=grouping)
    for jac in (True, False):
        lp, g, st = eng.log_prob_grad(d, u, adjust_transform=jac)
        lp_o, g_o, st_o = bo.log_prob(om, cfg["data"], u, jacobian=jac)
        assert (st == 0).all() and (st_o == 0).all()
        np.testing.assert_allclose(lp, lp_o, rtol=1e-11)
        np.testing.assert_allclose(g, g_o, rtol=1e-9, atol=1e-10 * np.abs(g_o).max())
    np.testing.assert_allclose(eng.constrain_pars(u)[:, :cfg["nparams"]], u[:, :cfg["nparams"]], rtol=0, atol=0)


@pytest.mark.parametrize("name,nobs", [("cfg2", None), ("cfg5U", 300), ("cfg5K", 200), ("cfg5G", 240), ("cfg4", None), ("cfg1", None), ("cfg3", None)])
def test_log_lik_per_observation(name, nobs):
    """bpc_log_lik element by element against the literal per-observation oracle (stanfiles/multitype_loo_block.stan:15-32,
    simple_birth_death_loo_block.stan:3-12), on a SHUFFLED data list with mixed durations: the device works on a table sorted by
    duration and un-permutes its output, so a wrong permutation shows here (sums cannot see it)."""
    cfg = syn.make_config(name, chains=3, nobs=nobs)
    rng = np.random.default_rng(12)
    d = dict(cfg["data"])
    N = int(d["ndatapts"])
    perm = rng.permutation(N)
    for k in ("pop_vec", "init_pop", "var_idx"):
        d[k] = np.asarray(d[k])[perm]
    if cfg["kind"] == "simple_bd":
        d["times"] = np.asarray(d["times"], dtype=float)[perm] * rng.choice([0.5, 1.0, 2.0], size=N)
    else:
        if name in ("cfg2", "cfg4", "cfg5G"):       # mix several durations into a data list that had one
            d["times"] = np.array([1.0, 0.4, 2.5, 1.7])
            d["ntimes_unique"] = 4
            d["times_idx"] = rng.integers(1, 5, size=N).astype(np.int32)
        else:
            d["times_idx"] = np.asarray(d["times_idx"])[perm]
    mod = bp.bp_model(cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"])
    eng = bp.stan_model(mod, cfg["priors"], simple_bd=cfg["kind"] == "simple_bd", noise_model=cfg["kind"] == "multitype_noise")
    om = bo.OracleModel(cfg["kind"], cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"], cfg["priors"])
    ll = eng.log_lik(d, cfg["upars"])
    ll_o = bo.log_lik(om, d, cfg["upars"])
    assert ll.shape == (3, N) and np.isfinite(ll_o).all()
    np.testing.assert_allclose(ll, ll_o, rtol=1e-10, atol=1e-10)
    # and a shuffled list gives the shuffled answer
    ll0 = eng.log_lik(cfg["data"], cfg["upars"]) if name in ("cfg5U", "cfg5K") else None
    if ll0 is not None:
        np.testing.assert_allclose(ll, ll0[:, perm], rtol=1e-13)


def test_octet_path_matches_oracle_at_bench_shape():
    """the headline kernel at the bench shape (BASELINE.json configs[4]: 100 000 observations; 64 of the 512 chains per GPU, 8 octets)
    against the CPU oracle on every observation — 6.4e6 evaluations, a few seconds of host time."""
    cfg = syn.make_config("cfg5U", seed=1, chains=64, nobs=100000)
    mod = bp.bp_model(cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"])
    eng = bp.stan_model(mod, cfg["priors"])
    om = bo.OracleModel(cfg["kind"], cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"], cfg["priors"])
    rng = np.random.default_rng(1000)
    u = cfg["u_true"][None, :] + 0.1 * rng.standard_normal((64, len(cfg["u_true"])))      # bench.py's chain positions
    d = eng.data(cfg["data"])
    assert "bp_fusedo" in d.path(64)[0]
    lp, g, st = eng.log_prob_grad(d, u)
    lp_o, g_o, st_o = bo.log_prob(om, cfg["data"], u, grouped=False)
    assert (st == 0).all() and (st_o == 0).all()
    np.testing.assert_allclose(lp, lp_o, rtol=1e-11)
    np.testing.assert_allclose(g, g_o, rtol=1e-9, atol=1e-10 * np.abs(g_o).max())


@pytest.mark.parametrize("chains,kernel", [(16, "bp_groupb"), (130, "bp_groupc")])
def test_grouped_path_with_phase_b_on_its_own_grid(chains, kernel):
    """grouped data with >= 16 384 observations: bp_grouped (phase A) -> bp_groupb ((tiles, chains) CTAs) or, from 128 chains, bp_groupc
    (one thread per chain, tiles staged in shared memory; 130 chains: a ragged last block) -> bp_grouped (phase A', epilogue) against
    the oracle and against the single-launch path (BPC_GROUPB=0); a not-positive-definite observation rejects exactly its chains;
    deterministic."""
    import os
    cfg = syn.make_config("cfg5G", chains=chains, nobs=20000)
    mod = bp.bp_model(cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"])
    eng = bp.stan_model(mod, cfg["priors"])
    om = bo.OracleModel(cfg["kind"], cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"], cfg["priors"])
    u = cfg["upars"].copy()
    u[1] += 0.75
    u[2] -= 1.5
    u[3] += 2.2                                    # sub-steps in phases A / A'
    d = eng.data(cfg["data"])
    assert d.grouped
    lp, g, st = eng.log_prob_grad(d, u)
    lp_o, g_o, st_o = bo.log_prob(om, cfg["data"], u)
    assert st.tolist() == st_o.tolist() and (st == 0).all()
    np.testing.assert_allclose(lp, lp_o, rtol=1e-11)
    for c in range(chains):
        np.testing.assert_allclose(g[c], g_o[c], rtol=1e-9, atol=1e-10 * np.abs(g_o[c]).max())
    b = eng.log_prob_grad(d, u)
    assert (b[0] == lp).all() and (b[1] == g).all()
    os.environ["BPC_GROUPB"] = "0"
    try:
        d1 = eng.data(cfg["data"])
        lp1, g1, st1 = eng.log_prob_grad(d1, u)
    finally:
        del os.environ["BPC_GROUPB"]
    np.testing.assert_allclose(lp, lp1, rtol=1e-12)
    np.testing.assert_allclose(g, g1, rtol=1e-9, atol=1e-11 * np.abs(g1).max())
    bad = dict(cfg["data"])
    ip = np.array(bad["init_pop"], dtype=float)
    ip[777] = 0.0                                  # Sigma = 0 for one observation: multi_normal rejects every chain
    bad["init_pop"] = ip
    lpb, gb, stb = eng.log_prob_grad(bad, u)
    assert (stb & 1).all() and np.isneginf(lpb).all() and (gb == 0).all()


def _unbox(v):
    """what jsonlite::write_json(auto_unbox = TRUE) makes of an R value: matrices row by row, length-one vectors as scalars"""
    a = np.asarray(v)
    if a.ndim == 0 or (a.ndim == 1 and a.size == 1):
        return a.reshape(-1)[0].item()
    return a.tolist()


@pytest.mark.parametrize("name", ["cfg2", "cfg3", "cfg4"])
def test_rstan_dump_comparer_reads_a_dump_shaped_like_parity_dump_R(name):
    """tools/compare_rstan_dump.py (the Python half of SURVEY.md §8(f) N4) on a dump with the layout r-package/parity_dump.R writes
    through jsonlite -- matrices row by row, `upars` / `grad` one COLUMN per fixture point, length-one vectors unboxed, log densities up
    to the additive constant Stan drops -- with the CPU oracle standing in for rstan.  Checks the comparer's plumbing (model and priors
    rebuilt from the JSON, data list, orientation, differences of lp); the numbers rstan would put there need a machine with R."""
    import importlib.util
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("compare_rstan_dump", os.path.join(root, "tools", "compare_rstan_dump.py"))
    cmp_mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cmp_mod)
    cfg = syn.make_config(name, chains=5)
    om = bo.OracleModel(cfg["kind"], cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"], cfg["priors"])
    u = cfg["upars"]
    lp, g, st = bo.log_prob(om, cfg["data"], u)
    lp_nj, _, _ = bo.log_prob(om, cfg["data"], u, jacobian=False)
    assert (st == 0).all()
    entry = dict(name=name, simple_bd=cfg["kind"] == "simple_bd", noise_model=cfg["kind"] == "multitype_noise",
                 model=dict(e_mat=_unbox(cfg["e_mat"]), p_vec=_unbox(cfg["p_vec"]), func_deps=list(cfg["func_deps"]), nparams=cfg["nparams"], ndep=cfg["ndep"]),
                 priors=[dict(name=p["name"], params=_unbox(p["params"]), bounds=_unbox(p["bounds"]) if p.get("bounds") is not None else "NA") for p in cfg["priors"]],
                 dat={k: _unbox(v) for k, v in cfg["data"].items()}, upars=u.T.tolist(),
                 lp=(lp + 1234.5).tolist(), lp_nojac=(lp_nj - 77.0).tolist(), grad=g.T.tolist())
    if name == "cfg2":
        entry["log_lik"] = bo.log_lik(om, cfg["data"], u[:1])[0].tolist()
    out = cmp_mod.compare(json.loads(json.dumps(entry)))
    assert out["status"] == [0] * 5
    assert out["lp_diff_rel"] < 1e-10 and out["lp_nojac_diff_rel"] < 1e-10 and out["grad_rel"] < 1e-8
    if name == "cfg2":
        assert out["log_lik_rel"] < 1e-10
