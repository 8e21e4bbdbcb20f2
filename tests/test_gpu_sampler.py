"""GPU tests of the device-resident NUTS driver and of the moment kernel against the reference's golden values."""
import numpy as np
import pytest

import bpinference_b200 as bp
from bpinference_b200 import synthetic as syn
from oracle import bp_oracle as bo

pytestmark = pytest.mark.gpu


def _engine(cfg):
    mod = bp.bp_model(cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"])
    return bp.stan_model(mod, cfg["priors"], simple_bd=cfg["kind"] == "simple_bd", noise_model=cfg["kind"] == "multitype_noise")


def test_vignette_golden_moments_on_gpu():
    """vignettes/two-type-inference.pdf (chunk two-type-inference.Rmd:38-44) and vignettes/first-model.pdf
    (first-model.Rmd:63-67): calculate_moments on the device reproduces the printed means to 14 digits and the printed
    covariances to the accuracy deSolve gave them (1e-5 relative); the exact values are held to 1e-11."""
    E2 = [[2, 0], [1, 1], [1, 1], [0, 2], [0, 0], [0, 0]]
    m = bp.calculate_moments(E2, [1, 1, 2, 2, 1, 2], [0.20, 0.10, 0.05, 0.20, 0.25, 0.20], [100, 50], 1)
    np.testing.assert_allclose(m["mu_mat"].ravel(), [97.8054257205992, 59.8852387358044], rtol=1e-14)
    np.testing.assert_allclose(m["sigma_mat"].ravel(), [45.0418961482491, 3.02422137683959, 3.02422137683959, 32.0727485045799], rtol=4e-5)
    np.testing.assert_allclose([m["sigma_mat"][0, 0], m["sigma_mat"][0, 1], m["sigma_mat"][1, 1]], [45.04201918116, 3.02413190691, 32.07279597047], rtol=1e-11)
    m1 = bp.calculate_moments([[2.0], [0.0]], [1, 1], [0.25, 0.10], [1000], 1)
    np.testing.assert_allclose(m1["mu_mat"].ravel(), [1161.83424272828], rtol=1e-14)
    np.testing.assert_allclose(m1["sigma_mat"].ravel(), [438.725421279098], rtol=1e-5)
    np.testing.assert_allclose(m1["sigma_mat"].ravel(), [438.72398464468], rtol=1e-12)


def test_template_state_matches_literal_ode():
    """the device moments in the template's own layout (m_t, d_t of multitype_template.stan:137-140) against the literal numpy
    restatement of the template's ODE integrated at 1e-13 (oracle/stan_template_ref.py)."""
    from oracle import stan_template_ref as ref
    rng = np.random.default_rng(4)
    for n in (1, 2, 3):
        E = 2 * n + 2
        e_mat = rng.integers(0, 3, size=(E, n)).astype(float)
        p_vec = rng.integers(1, n + 1, size=E)
        r = rng.uniform(0.02, 0.6, size=E)
        t = float(rng.uniform(0.4, 4.0))
        m_t, d_t = bp.template_moments(e_mat, p_vec, r, t)
        row = ref.exact(e_mat, p_vec, r, [t])[0]
        np.testing.assert_allclose(m_t, row[:n * n].reshape((n, n), order="F"), rtol=1e-9, atol=1e-11)
        ode_d = row[n * n:].reshape((n, n * n), order="F")
        np.testing.assert_allclose(d_t, ode_d, rtol=1e-8, atol=1e-9 * max(1.0, np.abs(ode_d).max()))


def test_one_type_closed_form_equals_multitype_kernel():
    """one_type_template.stan:42-43 against the n = 1 moment kernel (two template families, one answer)."""
    for b, d, t, z0 in [(0.25, 0.10, 1.0, 1000.0), (0.1, 0.3, 4.0, 50.0), (2.0, 1.5, 0.7, 7.0), (4.5, 4.0, 5.0, 3.0)]:
        g = b - d
        e = np.exp(t * g)
        var = (e * (b * (2 * e - 1) - d) / g - e * e) * z0
        m = bp.calculate_moments([[2.0], [0.0]], [1, 1], [b, d], [z0], t)
        np.testing.assert_allclose(m["mu_mat"][0, 0], e * z0, rtol=1e-13)
        np.testing.assert_allclose(m["sigma_mat"][0, 0], var, rtol=1e-11)


def test_simulated_moments_agree_with_kernel():
    """tests/testthat/test_simulation.R:109-127 with the reference's tolerances scaled to 2000 replicates."""
    E2 = np.array([[2, 0], [1, 1], [1, 1], [0, 2], [0, 0], [0, 0]], dtype=float)
    P2 = [1, 1, 2, 2, 1, 2]
    th = [0.20, 0.10, 0.05, 0.20, 0.25, 0.20]
    mod = bp.bp_model(E2, P2, ['c[%d]' % k for k in range(1, 7)], 6, 0)
    sim = bp.bpsims(mod, th, [100, 50], [5], 2000, rng=np.random.default_rng(11))
    mom = bp.calculate_moments(E2, P2, th, [100, 50], 5)
    x = sim.iloc[:, 2:4].to_numpy()
    assert abs(x[:, 0].mean() - mom["mu_mat"][0, 0]) < 2 and abs(x[:, 1].mean() - mom["mu_mat"][1, 0]) < 2
    c = np.cov(x.T)
    assert abs(c[0, 0] - mom["sigma_mat"][0, 0]) < 16 and abs(c[1, 1] - mom["sigma_mat"][1, 1]) < 16 and abs(c[0, 1] - mom["sigma_mat"][0, 1]) < 16


def _grid_posterior_cfg1(cfg, eng):
    """posterior mean / sd of (Theta1, Theta2) by brute-force quadrature of the device log density on a grid."""
    g1, g2 = np.linspace(0.05, 0.14, 181), np.linspace(0.005, 0.09, 171)
    T1, T2 = np.meshgrid(g1, g2, indexing="ij")
    th = np.stack([T1.ravel(), T2.ravel()], axis=1)
    lp = eng.log_prob(cfg["data"], eng.unconstrain_pars(th), adjust_transform=False)
    w = np.exp(lp - lp.max())
    w /= w.sum()
    mean = (w[:, None] * th).sum(0)
    sd = np.sqrt((w[:, None] * (th - mean) ** 2).sum(0))
    return mean, sd


def test_nuts_posterior_matches_quadrature():
    """cfg1 (examples/one_type.R): NUTS draws against the posterior integrated on a grid with the same log density."""
    cfg = syn.make_config("cfg1", chains=8)
    eng = _engine(cfg)
    mean_q, sd_q = _grid_posterior_cfg1(cfg, eng)
    fit = eng.sampling(cfg["data"], chains=8, iter=2000, control=dict(adapt_delta=0.95), seed=11)
    flat = fit.draws.reshape(-1, 2)
    ess = fit.ess_bulk()
    assert (fit.rhat < 1.05).all() and bp.check_div(fit) == 0 and (ess > 400).all()
    assert bp.check_rhat(fit) == "Rhat looks reasonable for all parameters"
    mcse = sd_q / np.sqrt(ess)
    assert (np.abs(flat.mean(0) - mean_q) < 5 * mcse + 1e-4).all(), (flat.mean(0), mean_q, mcse)
    np.testing.assert_allclose(flat.std(0), sd_q, rtol=0.15)
    assert 0.85 < fit.accept_stat__.mean() < 0.999
    assert "saturated the maximum tree depth of 10" in bp.check_exceeded_treedepth(fit)


def test_sampler_reproducible_and_shard_independent():
    """same seed -> identical draws; a chain's draws depend on its GLOBAL id only (Philox key), not on how chains are grouped."""
    cfg = syn.make_config("cfg4", chains=6)
    eng = _engine(cfg)
    d = eng.data(cfg["data"])
    a = eng.sampling(d, chains=6, iter=300, seed=5)
    b = eng.sampling(d, chains=6, iter=300, seed=5)
    assert (a.draws == b.draws).all() and (a.n_leapfrog__ == b.n_leapfrog__).all()
    c = eng.sampling(d, chains=6, iter=300, seed=6)
    assert not (a.draws == c.draws).all()
    assert a.draws.shape == (6, 150, 4) and a.names == ["Theta1", "Theta2", "Theta3", "sigma_obs"]
    assert (a.draws[:, :, :3] > 0).all() and (a.draws[:, :, :3] < 2).all() and (a.draws[:, :, 3] > 0).all()


def test_pooled_adaptation_and_user_inits():
    cfg = syn.make_config("cfg2", chains=32, nobs=120)
    eng = _engine(cfg)
    init = bp.uniform_initialize(np.tile([0.05, 0.4], (6, 1)), 32, rng=np.random.default_rng(2))
    fit = eng.sampling(cfg["data"], chains=32, iter=600, init=init, control=dict(adapt_delta=0.9, max_treedepth=10), seed=9, pooled_adaptation=True)
    assert (fit.rhat < 1.1).all(), fit.rhat
    assert fit.info["total_chains"] == 32 and fit.info["leapfrogs"] > 0
    # the posterior is centred near the simulation parameters for the identifiable net growth rates
    flat = fit.draws.reshape(-1, 6)
    net1 = flat[:, 0] - flat[:, 4] - flat[:, 1] * 0    # birth - death of type 1 (transfer terms cancel in expectation)
    assert abs(np.median(net1) - (0.20 - 0.25)) < 0.1


def test_device_simulator_matches_analytic_moments():
    """tests/testthat/test_simulation.R:109-127 at the reference's own size and tolerances (10 000 replicates, |d mean| < 1,
    |d var|, |d cov| < 7), with the replicates simulated by the Gillespie kernel; plus the host simulator's column layout."""
    E2 = np.array([[2, 0], [1, 1], [1, 1], [0, 2], [0, 0], [0, 0]], dtype=float)
    P2 = [1, 1, 2, 2, 1, 2]
    th = [0.20, 0.10, 0.05, 0.20, 0.25, 0.20]
    mod = bp.bp_model(E2, P2, ['c[%d]' % k for k in range(1, 7)], 6, 0)
    sim = bp.bpsims(mod, th, [100, 50], [5], 10000, device=0, seed=3)
    assert list(sim.columns) == ["times", "rep", "X1", "X2"] and len(sim) == 10000 and (sim["times"] == 5).all()
    mom = bp.calculate_moments(E2, P2, th, [100, 50], 5)
    x = sim.iloc[:, 2:4].to_numpy()
    assert abs(x[:, 0].mean() - mom["mu_mat"][0, 0]) < 1 and abs(x[:, 1].mean() - mom["mu_mat"][1, 0]) < 1
    c = np.cov(x.T)
    assert abs(c[0, 0] - mom["sigma_mat"][0, 0]) < 7 and abs(c[1, 1] - mom["sigma_mat"][1, 1]) < 7 and abs(c[0, 1] - mom["sigma_mat"][0, 1]) < 7
    # trajectories: first row is z0, rows are recorded at every requested time, reproducible by seed
    tr = bp.bp_cuda(E2, th, P2, [1000, 500], np.arange(6), 64, seed=9)
    assert tr.shape == (64, 6, 2) and (tr[:, 0] == [1000, 500]).all() and (tr >= 0).all()
    assert (bp.bp_cuda(E2, th, P2, [1000, 500], np.arange(6), 64, seed=9) == tr).all()
    dat = bp.stan_data_from_simulation(bp.bpsims(mod, th, [1000, 500], np.arange(6), 50, device=0), mod)
    assert dat["ndatapts"] == 250 and dat["ntimes_unique"] == 1 and dat["times"][0] == 1


def test_sampler_survives_other_calls_on_its_data_handle():
    """the sampler replays a CUDA graph with the data handle's workspace baked in.  Other entry points on the same handle
    with another chain count re-tile / reallocate that workspace between bpc_sampler_run calls; the sampler must notice (workspace
    generation), re-capture, and produce exactly the draws of an undisturbed run.  Also: a second sampler on a handle whose first
    sampler (and its stream) is gone."""
    import ctypes as C
    from bpinference_b200 import _ffi
    cfg = syn.make_config("cfg2", chains=32, nobs=120)
    eng = _engine(cfg)
    d = eng.data(cfg["data"])
    L = _ffi.lib()

    def run(disturb):
        args = _ffi.SampleArgs(32, 120, 60, 0.8, 8, 3, 0, None, 0, 0.0, 0, 0, None)
        h = C.c_void_p()
        _ffi.check(L.bpc_sampler_create(eng.handle, d.handle, C.byref(args), C.byref(h)))
        st = C.c_int(0)
        k = 0
        while st.value != 1:
            _ffi.check(L.bpc_sampler_run(h, 160, C.byref(st)))
            if disturb:
                k += 1
                n = 5 + (k % 3) * 20                       # 5, 25, 45 chains: different tilings and workspace sizes
                eng.log_prob_grad(d, cfg["upars"][:1].repeat(n, axis=0))
                if k % 4 == 0:
                    eng.log_lik(d, cfg["upars"][:2])
        draws = np.empty((32, 60, 6))
        nl = np.empty((32, 60), dtype=np.int32)
        _ffi.check(L.bpc_sampler_draws(h, _ffi.dptr(draws), None, None, None, _ffi.iptr(nl), None, None))
        L.bpc_sampler_free(h)
        return draws, nl
    a, na = run(False)
    b, nb = run(True)
    assert (a == b).all() and (na == nb).all()
    assert np.isfinite(a).all() and a.std() > 0


def test_stanfit_shaped_columns():
    """N2: columns of the fit in rstan's order — Theta1..P, r_mat (column-major), theta[i, ] = (e_mat, r_mat), lp__ — so that
    `samples[,1:6]` of vignettes/two-type-inference.Rmd:63-69 reads the six rates; the values against the oracle's restatement of the
    transformed-parameters block (multitype_template.stan:105-117, one_type_template.stan:19-27)."""
    for name, chains in (("cfg2", 4), ("cfg3", 4), ("cfg4", 4)):
        cfg = syn.make_config(name, chains=chains, nobs=60 if name == "cfg2" else None)
        eng = _engine(cfg)
        fit = eng.sampling(cfg["data"], chains=chains, iter=60, warmup=30, seed=2, loo=(name == "cfg2"))
        om = bo.OracleModel(cfg["kind"], cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"], cfg["priors"])
        flat = fit.draws.reshape(-1, fit.draws.shape[2])
        names, vals = bo.transformed_parameters(om, cfg["data"], syn_unconstrain(eng, flat))
        tn, tv = fit.transformed_parameters()
        assert list(tn) == names
        np.testing.assert_allclose(tv.reshape(-1, tv.shape[2]), vals, rtol=1e-13, atol=0)
        cols = fit.colnames
        P = cfg["nparams"]
        assert cols[:P] == ["Theta%d" % (k + 1) for k in range(P)] and cols[-1] == "lp__"
        M = fit.as_matrix()
        assert M.shape == (chains * 30, len(cols))
        np.testing.assert_array_equal(M[:, :P], flat[:, :P])
        np.testing.assert_array_equal(M[:, -1], fit.lp__.reshape(-1))
        ex = fit.extract()
        if name == "cfg2":
            assert cols[P] == "r_mat[1,1]" and cols[P + 12] == "theta[1,1]" and cols[-2] == "log_lik[60]"
            assert ex["r_mat"].shape == (chains * 30, 6, 2) and ex["theta"].shape == (chains * 30, 1, 24) and ex["log_lik"].shape == (chains * 30, 60)
            np.testing.assert_array_equal(ex["r_mat"][:, 0, 0], flat[:, 0])          # r_mat[1, p_vec[1]] = c[1]
            np.testing.assert_array_equal(ex["r_mat"][:, 2, 1], flat[:, 2])          # event 3 has parent type 2
            assert (ex["r_mat"][:, 2, 0] == 0).all()
            ll = bo.log_lik(om, cfg["data"], syn_unconstrain(eng, flat[:3]))
            np.testing.assert_allclose(ex["log_lik"][:3], ll, rtol=1e-10)
            s = fit.summary()
            assert list(s)[:6] == cols[:6] and list(s)[-1] == "lp__" and "r_mat[1,1]" in s
        if name == "cfg3":
            assert ex["r_mat"].shape == (chains * 30, 2, 6) and "theta" not in ex
        if name == "cfg4":
            assert cols[P] == "sigma_obs" and cols[P + 1] == "r_mat[1,1]"


def syn_unconstrain(eng, theta):
    return eng.unconstrain_pars(theta)
