"""GPU parity: libbpcuda (through the C ABI, via the ctypes mirror) against the CPU oracle on the same seeded inputs.

Tolerances: fp64 throughout; BASELINE.json's north star asks for 1e-9 relative on log-density and gradient,
the tests below hold the CUDA path to 1e-11 / 1e-9 of the oracle (the remaining difference is FMA contraction
and summation order)."""
import numpy as np
import pytest

import bpinference_b200 as bp
from bpinference_b200 import synthetic as syn
from oracle import bp_oracle as bo

pytestmark = pytest.mark.gpu


def _pair(cfg):
    mod = bp.bp_model(cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"])
    eng = bp.stan_model(mod, cfg["priors"], simple_bd=cfg["kind"] == "simple_bd", noise_model=cfg["kind"] == "multitype_noise")
    om = bo.OracleModel(cfg["kind"], cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"], cfg["priors"])
    return eng, om


@pytest.mark.parametrize("name,nobs,chains", [("cfg1", None, 4), ("cfg2", None, 64), ("cfg3", None, 4), ("cfg4", None, 4),
                                              ("cfg5U", 3000, 16), ("cfg5G", 3000, 16), ("cfg5K", 1500, 8)])
@pytest.mark.parametrize("jacobian", [True, False])
def test_log_prob_and_gradient_match_oracle(name, nobs, chains, jacobian):
    cfg = syn.make_config(name, chains=chains, nobs=nobs)
    eng, om = _pair(cfg)
    lp, grad, st = eng.log_prob_grad(cfg["data"], cfg["upars"], adjust_transform=jacobian)
    lp_o, g_o, st_o = bo.log_prob(om, cfg["data"], cfg["upars"], jacobian=jacobian)
    assert (st == 0).all() and (st_o == 0).all()
    np.testing.assert_allclose(lp, lp_o, rtol=1e-11)
    np.testing.assert_allclose(grad, g_o, rtol=1e-9, atol=1e-10 * np.abs(g_o).max())


def test_status_bits_match_oracle():
    cfg = syn.make_config("cfg1", chains=3)
    eng, om = _pair(cfg)
    u = cfg["upars"].copy()
    u[1, 0] = 3.0                      # Theta1 = 2 inv_logit(3) > 1: outside uniform(0,1) (examples/one_type.R:7)
    u[2, :] = 0.0                      # birth == death: 0/0 in one_type_template.stan:43
    lp, grad, st = eng.log_prob_grad(cfg["data"], u)
    lp_o, g_o, st_o = bo.log_prob(om, cfg["data"], u)
    assert st.tolist() == st_o.tolist()
    assert st[0] == 0 and st[1] & 8 and st[2] & 4
    assert np.isneginf(lp[1]) and np.isneginf(lp[2]) and (grad[1:] == 0).all()
    cfg = syn.make_config("cfg2", chains=2, nobs=20)
    eng, om = _pair(cfg)
    d = dict(cfg["data"])
    d["init_pop"] = np.zeros_like(d["init_pop"])       # Sigma = 0: multi_normal rejects
    lp, grad, st = eng.log_prob_grad(d, cfg["upars"])
    assert (st & 1).all() and np.isneginf(lp).all()


def test_deterministic_and_chain_order_independent():
    cfg = syn.make_config("cfg5U", chains=12, nobs=2000)
    eng, _ = _pair(cfg)
    d = eng.data(cfg["data"])
    a = eng.log_prob_grad(d, cfg["upars"])
    b = eng.log_prob_grad(d, cfg["upars"])
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all()          # bit-identical: fixed reduction order, no atomics on doubles
    perm = np.random.default_rng(0).permutation(12)
    c = eng.log_prob_grad(d, cfg["upars"][perm])
    assert (c[0] == a[0][perm]).all() and (c[1] == a[1][perm]).all()


def test_log_lik_matches_sum_and_constants():
    # (cfg4, the noise template, is covered element by element in test_gpu_corners.py: the reference's LOO block has no sigma_obs term)
    for name in ("cfg2", "cfg1"):
        cfg = syn.make_config(name, chains=3)
        eng, om = _pair(cfg)
        ll = eng.log_lik(cfg["data"], cfg["upars"])
        n = 1 if cfg["kind"] == "simple_bd" else cfg["data"]["ntypes"]
        lp_o, _, _ = bo.log_prob(om, cfg["data"], cfg["upars"], jacobian=False)
        theta, sig, *_ = bo.constrain(om, cfg["upars"])
        prior = sum(bo.prior_logpdf(pr["name"], pr["params"], theta[:, k])[0] for k, pr in enumerate(om.priors))
        if sig is not None:
            prior = prior - 0.5 * (sig / 0.1) ** 2
        N = cfg["data"]["ndatapts"]
        np.testing.assert_allclose(ll.sum(axis=1) + 0.5 * n * N * np.log(2 * np.pi), lp_o - prior, rtol=1e-11)


@pytest.mark.parametrize("name,nobs,chains", [("cfg2", None, 33), ("cfg4", None, 5), ("cfg5G", 5000, 9), ("cfg5K", 700, 3)])
def test_grouped_kernel_equals_fused_kernel(name, nobs, chains):
    """the template-shaped two-phase path (moments once per (level, duration) group) and the per-observation fused
    kernel are two evaluations of the same function: equal to rounding; both match the oracle."""
    cfg = syn.make_config(name, chains=chains, nobs=nobs)
    eng, om = _pair(cfg)
    if name == "cfg5K":          # a few repeated (level, duration) pairs among many distinct ones
        d = dict(cfg["data"])
        d["var_idx"] = np.asarray(d["var_idx"]).copy()
        d["times_idx"] = np.asarray(d["times_idx"]).copy()
        d["var_idx"][::2] = 1 + (np.arange(len(d["var_idx"][::2])) % 5)
        d["times_idx"][::2] = 1 + (np.arange(len(d["times_idx"][::2])) % 3)
        cfg["data"] = d
    df = eng.data(cfg["data"], grouping=1)
    dg = eng.data(cfg["data"], grouping=2)
    assert not df.grouped and dg.grouped and dg.ngroups >= 1
    lf, gf, sf = eng.log_prob_grad(df, cfg["upars"])
    lg, gg, sg = eng.log_prob_grad(dg, cfg["upars"])
    assert (sf == 0).all() and (sg == 0).all()
    np.testing.assert_allclose(lg, lf, rtol=1e-12)
    np.testing.assert_allclose(gg, gf, rtol=1e-9, atol=1e-11 * np.abs(gf).max())
    lp_o, g_o, _ = bo.log_prob(om, cfg["data"], cfg["upars"])
    np.testing.assert_allclose(lg, lp_o, rtol=1e-11)
    np.testing.assert_allclose(gg, g_o, rtol=1e-9, atol=1e-10 * np.abs(g_o).max())
    auto = eng.data(cfg["data"])
    assert auto.grouped == (name != "cfg5K")


def test_grouped_status_and_large_norm():
    cfg = syn.make_config("cfg2", chains=4, nobs=40)
    eng, om = _pair(cfg)
    d = dict(cfg["data"])
    d["times"] = np.array([9.0])
    u = cfg["upars"] + 1.5                       # rates ~ 1, ||tL|| ~ 40: many sub-steps in phase A / A'
    dg = eng.data(d, grouping=2)
    lg, gg, sg = eng.log_prob_grad(dg, u)
    lp_o, g_o, st_o = bo.log_prob(om, d, u)
    assert sg.tolist() == st_o.tolist()
    ok = st_o == 0
    np.testing.assert_allclose(lg[ok], lp_o[ok], rtol=1e-10)
    np.testing.assert_allclose(gg[ok], g_o[ok], rtol=1e-8, atol=1e-9 * np.abs(g_o).max())
    d0 = dict(cfg["data"])
    d0["init_pop"] = np.zeros_like(d0["init_pop"])
    l0, g0, s0 = eng.log_prob_grad(eng.data(d0, grouping=2), cfg["upars"])
    assert (s0 & 1).all() and np.isneginf(l0).all() and (g0 == 0).all()


def test_w_path_and_ring_fallback_in_one_call():
    """per-observation data with chain-level rates: chains whose series fit one sub-step take the W path (tensor-core
    accumulation + per-chain matrix polynomial), chains with large ||tL|| fall back to the ring kernel — in the same call."""
    cfg = syn.make_config("cfg5U", chains=10, nobs=1500)
    eng, om = _pair(cfg)
    u = cfg["upars"].copy()
    u[::2] += 2.2                      # every other chain: rates ~ 9x larger -> 2 ||A|| t_max > theta_cap -> sub-steps
    u[1::4] += 0.75                    # rates ~ 2x larger: 2 ||A|| t_max ~ 3.3, the long (up to 32 terms) single-step series of the W path
    d = eng.data(cfg["data"], grouping=1)
    lp, g, st = eng.log_prob_grad(d, u)
    lp_o, g_o, st_o = bo.log_prob(om, cfg["data"], u)
    assert st.tolist() == st_o.tolist()
    ok = st_o == 0
    assert ok.sum() >= 5
    np.testing.assert_allclose(lp[ok], lp_o[ok], rtol=1e-11)
    np.testing.assert_allclose(g[ok], g_o[ok], rtol=1e-9, atol=1e-10 * np.abs(g_o[ok]).max())
    # the two paths agree with each other chain by chain as well
    import os
    os.environ["BPC_NO_WPATH"] = "1"
    try:
        d2 = eng.data(cfg["data"], grouping=1)
        lp2, g2, st2 = eng.log_prob_grad(d2, u)
    finally:
        del os.environ["BPC_NO_WPATH"]
    assert st2.tolist() == st.tolist()
    np.testing.assert_allclose(lp[ok], lp2[ok], rtol=1e-12)
    np.testing.assert_allclose(g[ok], g2[ok], rtol=1e-9, atol=1e-11 * np.abs(g2[ok]).max())


@pytest.mark.parametrize("n", [4, 5])
def test_four_and_five_type_models(n):
    """ntypes = 4, 5 (the oracle's and the kernels' upper range): ring kernel, grouped kernel, moments and log_lik."""
    rng = np.random.default_rng(3)
    E = 3 * n
    e_mat = np.zeros((E, n))
    p_vec = np.zeros(E, dtype=np.int64)
    for i in range(n):
        e_mat[3 * i, i] = 2.0
        e_mat[3 * i + 1, (i + 1) % n] = 1.0
        e_mat[3 * i + 1, i] = 1.0 if i % 2 else 0.0     # some asymmetric division events
        p_vec[3 * i: 3 * i + 3] = i + 1
    theta = rng.uniform(0.03, 0.3, size=E)
    N = 240
    z0 = np.rint(rng.uniform(300, 3000, size=(N, n)))
    tu = np.array([0.7, 1.5, 3.0])
    tidx = rng.integers(1, 4, size=N).astype(np.int32)
    mu, S = syn.gaussian_moments(e_mat, p_vec, theta, z0, tu[tidx - 1])
    data = dict(ntypes=n, nevents=E, ndatapts=N, ntimes_unique=3, ndep_levels=1, ndep=1, nparams=E, e_mat=e_mat, p_vec=p_vec,
                pop_vec=syn._draw(rng, mu, S), init_pop=z0, times=tu, times_idx=tidx, function_var=np.zeros((1, 1)),
                var_idx=np.ones(N, dtype=np.int32))
    priors = [dict(name="normal", params=[0.0, 0.25], bounds=[0.0, 5.0]) for _ in range(E)]
    cfg = dict(kind="multitype", e_mat=e_mat, p_vec=p_vec, func_deps=["c[%d]" % (k + 1) for k in range(E)], nparams=E, ndep=0, priors=priors)
    eng, om = _pair(cfg)
    u = syn.unconstrain(theta, priors)[None, :] + 0.1 * rng.standard_normal((5, E))
    lp_o, g_o, st_o = bo.log_prob(om, data, u)
    assert (st_o == 0).all()
    for grouping in (1, 2):
        lp, g, st = eng.log_prob_grad(eng.data(data, grouping=grouping), u)
        assert (st == 0).all()
        np.testing.assert_allclose(lp, lp_o, rtol=1e-11)
        np.testing.assert_allclose(g, g_o, rtol=1e-9, atol=1e-10 * np.abs(g_o).max())
    mu_d, sig_d = eng.moments(data, u[:1])
    th0 = eng.constrain_pars(u[0])
    mu_h, S_h = syn.gaussian_moments(e_mat, p_vec, th0, z0, tu[tidx - 1], substeps=16, terms=30)
    np.testing.assert_allclose(mu_d[0], mu_h, rtol=1e-10)
    np.testing.assert_allclose(sig_d[0], S_h, rtol=1e-8, atol=1e-8 * np.abs(S_h).max())
    ll = eng.log_lik(data, u)
    lp_nj, _, _ = bo.log_prob(om, data, u, jacobian=False)
    th, *_ = bo.constrain(om, u)
    prior = sum(bo.prior_logpdf(pr["name"], pr["params"], th[:, k])[0] for k, pr in enumerate(om.priors))
    np.testing.assert_allclose(ll.sum(axis=1) + 0.5 * n * N * np.log(2 * np.pi), lp_nj - prior, rtol=1e-11)


def test_full_size_properties_cfg5U():
    """BASELINE.json's full observation count (100 000 observations, 3-type model; 32 of the 512 chains per GPU to keep the
    test short): size-independent properties instead of an oracle pass that would take minutes on the CPU —
    additivity over a split of the observations, invariance under a permutation of the observations, the sum of the
    per-observation log_lik, a directional finite difference of the gradient, and a sampled oracle comparison."""
    cfg = syn.make_config("cfg5U", chains=32, nobs=100000)
    eng, om = _pair(cfg)
    d = cfg["data"]
    u = cfg["upars"]
    N = d["ndatapts"]
    lp, g, st = eng.log_prob_grad(d, u, adjust_transform=False)
    assert (st == 0).all() and np.isfinite(lp).all() and np.isfinite(g).all()
    theta, *_ = bo.constrain(om, u)
    prior = sum(bo.prior_logpdf(pr["name"], pr["params"], theta[:, k])[0] for k, pr in enumerate(om.priors))
    dth = theta * (1 - theta / 5.0)                       # d theta / d u of the (0, 5) bound transform
    pgrad = np.stack([bo.prior_logpdf(pr["name"], pr["params"], theta[:, k])[1] for k, pr in enumerate(om.priors)], axis=1) * dth

    def subset(idx):
        s = dict(d)
        s["ndatapts"] = len(idx)
        for k in ("pop_vec", "init_pop"):
            s[k] = np.asarray(d[k])[idx]
        s["var_idx"] = np.asarray(d["var_idx"])[idx]
        s["times_idx"] = np.asarray(d["times_idx"])[idx]
        return s
    rng = np.random.default_rng(0)
    perm = rng.permutation(N)
    a, b = perm[: N // 3], perm[N // 3:]
    la, ga, _ = eng.log_prob_grad(subset(a), u, adjust_transform=False)
    lb, gb, _ = eng.log_prob_grad(subset(b), u, adjust_transform=False)
    np.testing.assert_allclose(la + lb - prior, lp, rtol=1e-12)                  # likelihood is a sum over observations
    np.testing.assert_allclose(ga + gb - pgrad, g, rtol=1e-9, atol=1e-10 * np.abs(g).max())
    lq, gq, _ = eng.log_prob_grad(subset(perm), u, adjust_transform=False)       # order of the data list is irrelevant
    np.testing.assert_allclose(lq, lp, rtol=1e-13)
    np.testing.assert_allclose(gq, g, rtol=1e-10, atol=1e-11 * np.abs(g).max())
    ll = eng.log_lik(d, u[:4])
    np.testing.assert_allclose(ll.sum(axis=1) + 0.5 * 3 * N * np.log(2 * np.pi), (lp - prior)[:4], rtol=1e-12)
    v = rng.standard_normal(u.shape)
    h = 1e-6
    fd = (eng.log_prob(d, u + h * v, adjust_transform=False) - eng.log_prob(d, u - h * v, adjust_transform=False)) / (2 * h)
    np.testing.assert_allclose((g * v).sum(axis=1), fd, rtol=2e-6)
    idx = np.sort(rng.choice(N, 2000, replace=False))                            # sampled oracle comparison, 2 chains
    lo, go, so = bo.log_prob(om, subset(idx), u[:2], jacobian=False)
    ls, gs, ss = eng.log_prob_grad(subset(idx), u[:2], adjust_transform=False)
    np.testing.assert_allclose(ls, lo, rtol=1e-11)
    np.testing.assert_allclose(gs, go, rtol=1e-9, atol=1e-10 * np.abs(go).max())


@pytest.mark.parametrize("nobs,chunks", [(20000, None), (60000, None), (60000, "3")])
def test_centred_w_path_matches_oracle_and_w_path(nobs, chunks):
    """dense sorted durations: the centred W path (bp_ctable + bp_fusedc: short series around shared centres, Toeplitz
    product into the W matrices) against the oracle, against bp_fusedw on the same data, with chains of small and large
    ||A|| t (long centre series) and one chain that needs sub-steps (ring fallback) in the same call."""
    import os
    cfg = syn.make_config("cfg5U", chains=6, nobs=nobs)
    eng, om = _pair(cfg)
    u = cfg["upars"].copy()
    u[1] += 0.75                       # 2 ||A|| t_max ~ 3.3: centre series of up to 32 terms
    u[2] -= 1.5                        # small rates: short series everywhere
    u[3] += 2.2                        # sub-steps: bp_fused
    if chunks:
        os.environ["BPC_CGROUP_CHUNKS"] = chunks
    try:
        d = eng.data(cfg["data"])
    finally:
        os.environ.pop("BPC_CGROUP_CHUNKS", None)
    assert d.path(6)[0].startswith("bp_ctable")
    lp, g, st = eng.log_prob_grad(d, u)
    lp_o, g_o, st_o = bo.log_prob(om, cfg["data"], u)
    assert st.tolist() == st_o.tolist() and (st == 0).all()
    np.testing.assert_allclose(lp, lp_o, rtol=1e-11)
    np.testing.assert_allclose(g, g_o, rtol=1e-9, atol=1e-10 * np.abs(g_o).max())
    os.environ["BPC_NO_CPATH"] = "1"
    try:
        d2 = eng.data(cfg["data"])
    finally:
        del os.environ["BPC_NO_CPATH"]
    assert d2.path(6)[0].startswith("bp_fusedw")
    lp2, g2, st2 = eng.log_prob_grad(d2, u)
    np.testing.assert_allclose(lp, lp2, rtol=1e-12)
    np.testing.assert_allclose(g, g2, rtol=1e-9, atol=1e-11 * np.abs(g2).max())
    b = eng.log_prob_grad(d, u)        # deterministic
    assert (b[0] == lp).all() and (b[1] == g).all()


@pytest.mark.parametrize("name,nobs,chains,grouping", [("cfg5U", 20000, 19, None), ("cfg5U", 60000, 16, None), ("cfg2", None, 24, 1),
                                                       ("cfg4", None, 17, 1)])
def test_octet_path_matches_oracle_and_centred_path(name, nobs, chains, grouping):
    """octet path (bp_otable + bp_fusedo + bp_toep: eight chains share the tensor tiles of a pass) against the
    oracle and against bp_fusedc on the same data; a ragged last octet, chains with long and short centre series, one chain
    that needs sub-steps (ring fallback, listed by bp_otable) in the same call; two-type and noise models on ungrouped data."""
    import os
    cfg = syn.make_config(name, chains=chains, nobs=nobs)
    eng, om = _pair(cfg)
    u = cfg["upars"].copy()
    if name == "cfg5U":
        u[1] += 0.75                   # 2 ||A|| t_max ~ 3.3: centre series of up to 32 terms
        u[2] -= 1.5                    # small rates: short series everywhere
        u[11] += 2.2                   # sub-steps: bp_fused
    kw = {} if grouping is None else {"grouping": grouping}
    d = eng.data(cfg["data"], **kw)
    assert "bp_fusedo" in d.path(chains)[0]
    lp, g, st = eng.log_prob_grad(d, u)
    lp_o, g_o, st_o = bo.log_prob(om, cfg["data"], u)
    assert st.tolist() == st_o.tolist() and (st == 0).all()
    np.testing.assert_allclose(lp, lp_o, rtol=1e-11)
    np.testing.assert_allclose(g, g_o, rtol=1e-9, atol=1e-10 * np.abs(g_o).max())
    os.environ["BPC_OPATH"] = "0"
    try:
        d2 = eng.data(cfg["data"], **kw)
        assert "bp_fusedc" in d2.path(chains)[0]
        lp2, g2, st2 = eng.log_prob_grad(d2, u)
    finally:
        del os.environ["BPC_OPATH"]
    np.testing.assert_allclose(lp, lp2, rtol=1e-12)
    np.testing.assert_allclose(g, g2, rtol=1e-9, atol=1e-11 * np.abs(g2).max())
    b = eng.log_prob_grad(d, u)        # deterministic
    assert (b[0] == lp).all() and (b[1] == g).all()
    # fewer chains than the octet threshold on the same handle: back to bp_fusedc, same numbers
    lp3, g3, st3 = eng.log_prob_grad(d, u[:5])
    np.testing.assert_allclose(lp3, lp[:5], rtol=1e-12)
    np.testing.assert_allclose(g3, g[:5], rtol=1e-9, atol=1e-11 * np.abs(g).max())


def test_octet_path_with_more_groups_than_one_batch_of_the_group_table():
    """narrow groups (8 chunks instead of 32): more than 128 groups, so bp_otable's N_0 loop (one thread per group, batches of 128)
    takes a second batch, bp_fusedo runs short groups and bp_toep sums over many of them; against the oracle."""
    import os
    cfg = syn.make_config("cfg5U", chains=16, nobs=40000)
    eng, om = _pair(cfg)
    u = cfg["upars"].copy()
    u[1] += 0.75
    u[2] -= 1.5
    os.environ["BPC_CGROUP_CHUNKS"] = "8"
    try:
        d = eng.data(cfg["data"])
    finally:
        del os.environ["BPC_CGROUP_CHUNKS"]
    name, ctas, _ = d.path(16)
    assert "bp_fusedo" in name and ctas > 128, (name, ctas)
    lp, g, st = eng.log_prob_grad(d, u)
    lp_o, g_o, st_o = bo.log_prob(om, cfg["data"], u)
    assert st.tolist() == st_o.tolist() and (st == 0).all()
    np.testing.assert_allclose(lp, lp_o, rtol=1e-11)
    np.testing.assert_allclose(g, g_o, rtol=1e-9, atol=1e-10 * np.abs(g_o).max())


def test_octet_path_one_type_multitype_template():
    """ntypes = 1 under the multitype template with per-observation durations: the octet path's smallest shape
    (D = 2: two (cc, d) tiles, one row tile) against the oracle and against bp_fusedc."""
    import os
    rng = np.random.default_rng(7)
    N, chains = 5000, 16
    t = rng.uniform(0.5, 3.0, N)
    z0 = np.rint(rng.uniform(500, 2000, N))
    x = np.rint(z0 * np.exp(0.05 * t) + rng.standard_normal(N) * np.sqrt(0.15 * z0 * t))
    mod = bp.bp_model([[2], [0]], [1, 1], ["c[1]", "c[2]"], 2, 0)
    priors = [bp.prior_dist("uniform", [0, 1], [0, 2])] * 2
    data = bp.create_stan_data(mod, x, z0, t)
    eng = bp.stan_model(mod, priors)
    om = bo.OracleModel("multitype", [[2.0], [0.0]], [1, 1], ["c[1]", "c[2]"], 2, 0, priors)
    u = np.log(np.array([0.1, 0.05]) / (2 - np.array([0.1, 0.05])))[None, :] + 0.1 * rng.standard_normal((chains, 2))
    d = eng.data(data)
    assert "bp_fusedo" in d.path(chains)[0]
    lp, g, st = eng.log_prob_grad(d, u)
    lp_o, g_o, st_o = bo.log_prob(om, data, u)
    assert (st == 0).all() and (st_o == 0).all()
    np.testing.assert_allclose(lp, lp_o, rtol=1e-11)
    np.testing.assert_allclose(g, g_o, rtol=1e-9, atol=1e-10 * np.abs(g_o).max())
    os.environ["BPC_OPATH"] = "0"
    try:
        d2 = eng.data(data)
        lp2, g2, st2 = eng.log_prob_grad(d2, u)
    finally:
        del os.environ["BPC_OPATH"]
    np.testing.assert_allclose(lp, lp2, rtol=1e-12)
    np.testing.assert_allclose(g, g2, rtol=1e-9, atol=1e-11 * np.abs(g2).max())


@pytest.mark.parametrize("env", ["BPC_O2", "BPC_O9"])
@pytest.mark.parametrize("name,nobs,chains,grouping", [("cfg5U", 40000, 19, None), ("cfg2", None, 24, 1)])
def test_octet_kernel_variants_agree_bitwise(env, name, nobs, chains, grouping):
    """the measured alternatives of the octet passes (bp_fusedo2: tensor warps + score warps with named barriers; bp_fusedo9: n = 3,
    two barriers per pass) do the arithmetic of bp_fusedo: identical lp and gradient for bp_fusedo2 (same order of operations), round-off
    agreement for bp_fusedo9; a ragged last octet and a chain on the sub-step fallback included.  (The variant is chosen when a model's kernels are loaded.)"""
    import os
    cfg = syn.make_config(name, chains=chains, nobs=nobs)
    u = cfg["upars"].copy()
    if name == "cfg5U":
        u[1] += 0.75
        u[2] -= 1.5
        u[11] += 2.2
    kw = {} if grouping is None else {"grouping": grouping}
    eng, _ = _pair(cfg)
    d = eng.data(cfg["data"], **kw)
    assert "bp_fusedo" in d.path(chains)[0]
    lp, g, st = eng.log_prob_grad(d, u)
    os.environ[env] = "1"
    try:
        eng2, _ = _pair(cfg)
        d2 = eng2.data(cfg["data"], **kw)
        lp2, g2, st2 = eng2.log_prob_grad(d2, u)
    finally:
        del os.environ[env]
    assert st.tolist() == st2.tolist()
    if env == "BPC_O2":                  # same operations in the same order
        assert (lp == lp2).all() and (g == g2).all()
    else:                                # bp_fusedo9 sums the observations of a pass in another order
        np.testing.assert_allclose(lp2, lp, rtol=1e-13)
        np.testing.assert_allclose(g2, g, rtol=1e-10, atol=1e-12 * np.abs(g).max())


@pytest.mark.parametrize("name,nobs,chains", [("cfg2", None, 12), ("cfg5G", 1200, 12), ("cfg4", None, 8)])
def test_grouped_multi_step_cooperative_path(name, nobs, chains):
    """bp_grouped picks its path per chain: single-step cooperative (rates near the truth), cooperative with sub-steps
    (F = exp(h L) as a matrix, states Y_i kept, Lbar from the D-column polynomial) for 2 ||A|| t up to a few tens, and the
    per-thread series beyond what the sub-step states' shared memory holds.  All three against the oracle and against the
    per-observation kernel in one call (this is what a sampler pass sees during warm-up)."""
    cfg = syn.make_config(name, chains=chains, nobs=nobs)
    eng, om = _pair(cfg)
    u = cfg["upars"].copy()
    P = cfg["nparams"]
    u[1, :P] += 1.2      # rates x ~3
    u[2, :P] += 2.0      # x ~6: sub-steps
    u[3, :P] += 2.8      # x ~10
    u[4, :P] += 3.5      # close to the upper bound
    u[5, :P] = 6.0       # at the bound: many sub-steps
    u[6, 0] += 3.0       # one large rate
    dg = eng.data(cfg["data"], grouping=2)
    df = eng.data(cfg["data"], grouping=1)
    lg, gg, sg = eng.log_prob_grad(dg, u)
    lf, gf, sf = eng.log_prob_grad(df, u)
    lp_o, g_o, st_o = bo.log_prob(om, cfg["data"], u)
    assert sg.tolist() == st_o.tolist() == sf.tolist()
    ok = st_o == 0
    assert ok.sum() >= chains - 3
    np.testing.assert_allclose(lg[ok], lp_o[ok], rtol=1e-10)
    np.testing.assert_allclose(lg[ok], lf[ok], rtol=1e-10)
    for c in np.flatnonzero(ok):     # gradients chain by chain: their scales differ by orders of magnitude
        np.testing.assert_allclose(gg[c], g_o[c], rtol=1e-8, atol=1e-9 * np.abs(g_o[c]).max(), err_msg="chain %d vs oracle" % c)
        np.testing.assert_allclose(gg[c], gf[c], rtol=1e-8, atol=1e-9 * np.abs(gf[c]).max(), err_msg="chain %d vs bp_fused" % c)
    b = eng.log_prob_grad(dg, u)
    assert (b[0][ok] == lg[ok]).all() and (b[1][ok] == gg[ok]).all()
