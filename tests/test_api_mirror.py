"""The reference's own testthat suites, restated against the Python mirror of its R interface.

Each test names the reference test it mirrors (paths under /root/reference/tests/testthat).  They run without a
GPU: validation, the expression front end (C++, through the C ABI), prior strings, data conversion, simulator."""
import warnings

import numpy as np
import pytest

import bpinference_b200 as bp

E1 = np.array([[2.0], [0.0]])
E2 = np.array([[2, 0], [1, 1], [1, 1], [0, 2], [0, 0], [0, 0]], dtype=float)
P2 = [1, 1, 2, 2, 1, 2]
LOGI = 'c[1] + c[2]/(1 + exp(c[3]*(x[1] - c[4])))'

BAD_EXPR = [("c[10]", 5, 5, "Parameter c[10] goes beyond the number of parameters specified."),
            ("c[7]", 5, 5, "Parameter c[7] goes beyond the number of parameters specified."),
            ("c[0]", 5, 5, "Parameter c[0] goes beyond the number of parameters specified."),
            ("x[10]", 5, 5, "Variable x[10] goes beyond the number of dependent variables specified."),
            ("x[7]", 5, 5, "Variable x[7] goes beyond the number of dependent variables specified."),
            ("x[0]", 5, 5, "Variable x[0] goes beyond the number of dependent variables specified."),
            ("c[6] + x[-1]", 6, 5, "Invalid expression: x[-1]"), ("c[-6] + x[3]", 6, 5, "Invalid expression: c[-6]"),
            ("c[abcd]", 1, 1, "Invalid expression: c[abcd]"), ("c[]", 6, 5, "Invalid expression: c[]"),
            ("c[4]*a[5]", 6, 5, "Invalid expression: a[5]"), ("c[1+1]", 6, 5, "Invalid expression: c[1 + 1]"),
            ("c[3] + c11", 6, 5, "Invalid expression: c11"), ("c[[3]]", 6, 5, "Invalid expression: c[[3]]"),
            ('"hi mom"', 6, 5, 'Invalid expression: "hi mom"'), ("x[2] & c[3]", 6, 5, "Invalid expression: x[2] & c[3]"),
            ("c(1)", 6, 5, "Invalid expression: c(1)"), ("exp(c[1],c[3])", 6, 5, "log and exp take only one parameter"),
            ("exp(c[1],c[3], c[2])", 6, 5, "log and exp take only one parameter"), ("sin(x[1])", 6, 5, "Invalid expression: sin(x[1])")]


@pytest.mark.parametrize("expr,np_,nd,msg", BAD_EXPR)
def test_expression_parsing_invalid(expr, np_, nd, msg):
    """test_expression_parsing.R:3-47 (exp_to_stan and check_valid raise the same 20 messages)."""
    for fn in (bp.exp_to_stan, bp.check_valid):
        with pytest.raises(ValueError) as e:
            fn(expr, np_, nd)
        assert str(e.value) == msg


def test_expression_parsing_translations():
    """test_expression_parsing.R:49-63."""
    cases = {"c[1]": "Theta1", "x[5]": "function_var[i,5]", "exp(c[3]) + log(x[2])": "( exp(Theta3) ) + ( log(function_var[i,2]) )",
             "log(x[4])*(c[5]*(c[4]*x[1]))": "( log(function_var[i,4]) ) * ( ( Theta5 ) * ( ( Theta4 ) * ( function_var[i,1] ) ) )",
             "c[1]*exp(x[1]*x[2])": "( Theta1 ) * ( exp(( function_var[i,1] ) * ( function_var[i,2] )) )"}
    for src, want in cases.items():
        assert bp.exp_to_stan(src, 5, 5) == want
        bp.check_valid(src, 5, 5)


def test_r_precedence_and_deparse():
    assert bp.deparse("-c[1]^2") == "-c[1]^2"
    assert bp.deparse("2^-1*c[2]") == "2^-1 * c[2]"
    assert bp.deparse("c[1]/c[2]-c[3]") == "c[1]/c[2] - c[3]"
    assert bp.deparse("1e5*c[1]+.5") == "1e+05 * c[1] + 0.5"
    assert bp.exp_to_stan("-c[1]^2", 2, 0) == "-(( Theta1 ) ^ ( 2 ))"
    assert bp.exp_to_stan("(c[1])", 2, 0) == "Theta1"


def test_bpmodel_invalid():
    """test_bpmodel.R:3-58."""
    f2 = [LOGI, 'c[5] - c[6]/(1 + exp(c[7]*(x[1] - c[11])))']
    with pytest.raises(ValueError, match=r"Parameter c\[11\] goes beyond the number of parameters specified."):
        bp.bp_model(E1, [1, 1], f2, 8, 1)
    dims = "Dimensions of e_mat, p_vec, and func_deps do not agree"
    with pytest.raises(ValueError, match=dims):
        bp.bp_model(np.array([[2.0], [3.0], [0.0]]), [1, 1], ['c[1]', 'c[2]', 'c[3]'], 8, 1)
    em = np.array([[2.0, 1.0], [0.0, 1.0]])
    with pytest.raises(ValueError, match=dims):
        bp.bp_model(em, [1, 1, 1], ['c[1]', 'c[2]'], 8, 1)
    pmsg = "p_vec must be a vector of parents for each bith event"
    with pytest.raises(ValueError, match=pmsg):
        bp.bp_model(em, [1, -1], ['c[1]', 'c[2]'], 8, 1)
    with pytest.raises(ValueError, match=dims):
        bp.bp_model(em, [1, 1], ['c[1]', 'c[2]', 'c[3]'], 8, 1)
    with pytest.raises(ValueError, match=r"Variable x\[3\] goes beyond the number of dependent variables specified."):
        bp.bp_model(em, [1, 1], ['c[1]', 'c[2]*x[3]'], 8, 1)
    with pytest.raises(ValueError, match=pmsg):
        bp.bp_model(em, [1, 1.5], ['c[1]', 'c[2]*x[3]'], 8, 4)
    with pytest.raises(ValueError, match="nparams should be an integer number of parameters"):
        bp.bp_model(em, [1, 1], ['c[1]', 'c[2]*x[3]'], -8, 1)
    with pytest.raises(ValueError, match="Invalid expression: hi"):
        bp.bp_model(em, [1, 1], ['hi', 'mom'], 8, 1)
    num = "e_mat, p_mat, ndep, and nparams must all be numeric!"
    for args in [("all", [1, 1], ['c[1]', 'c[2]'], 2, 1), (em, "your", ['c[1]', 'c[2]'], 2, 1), (em, [1, 1], ['c[1]', 'c[2]'], "base", 1),
                 (em, [1, 1], ['c[1]', 'c[2]'], 2, "arebelongtous")]:
        with pytest.raises(ValueError, match=num):
            bp.bp_model(*args)
    with pytest.raises(ValueError, match="func_deps must be a vector of strings relating brith rates to dependent variables"):
        bp.bp_model(em, [1, 1], print, 2, 1)


def test_bpmodel_valid_roundtrip():
    """test_bpmodel.R:60-70: fields and deparse(func_deps) round-trip."""
    fd = ['c[1] + c[2]/(1 + exp(c[3] * (x[1] - c[4])))', 'c[5] - c[6]/(1 + exp(c[7] * (x[1] - c[8])))']
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        mod = bp.bp_model(E1, [1, 1], fd, 8, 1)
    assert (mod["e_mat"] == E1).all() and list(mod["p_vec"]) == [1, 1]
    assert mod["func_deps"] == fd
    with pytest.warns(UserWarning, match="e_mat is a vector, not a matrix. Converting to a one-column matrix"):
        bp.bp_model([2, 0], [1, 1], ['c[1]', 'c[2]'], 2, 0)


PRIOR_BAD = [(("are you", [1, 2, 3], [2, 3]), "Invalid prior name!"), ((24, [1, 2, 3], [2, 3]), "prior name must be a string!"),
             (("normall", [1, 2, 3], [2, 3]), "Invalid prior name!"), (("normal", [1, 2, 3], [2, 3]), "Incorrect number of parameters for prior distribution"),
             (("normal", [1, 0], [2, 3]), "Positive parameter required here!"), (("normal", [1, -3], [2, 3]), "Positive parameter required here!"),
             (("lognormal", [1, -3], [2, 3]), "Positive parameter required here!"), (("exponential", [1, -3], [2, 3]), "Incorrect number of parameters for prior distribution"),
             (("exponential", [0], [2, 3]), "Positive parameter required here!"), (("exponential", [-1], [2, 3]), "Positive parameter required here!"),
             (("exponential", "feeling", [2, 3]), "parameters should be numeric!"), (("gamma", [1, 1], "it"), "bounds should be a list of 2 numbers!"),
             (("gamma", [1, 1], "now"), "bounds should be a list of 2 numbers!"), (("gamma", [1, 1], [4, 5, 6]), "bounds should be a list of 2 numbers!"),
             (("gamma", [1, 1], [25, 24]), "Error: lower bounder greater or equal to upper bound!"), (("gamma", [-1, 1], [24, 25]), "Positive parameter required here!"),
             (("uniform", [1, -1], [-2, 2]), "upper bound must be greater than lower in uniform distribution!"),
             (("MR KRABS?!?!?!?", [1, 2, 3], [2, 3]), "Invalid prior name!")]


@pytest.mark.parametrize("args,msg", PRIOR_BAD)
def test_priors_invalid(args, msg):
    """test_priors.R:3-22."""
    with pytest.raises(ValueError) as e:
        bp.prior_dist(*args)
    assert str(e.value) == msg


def test_priors_parsed():
    """test_priors.R:24-33: exact declaration and sampling-statement strings."""
    assert bp.parse_prior(bp.prior_dist("normal", [0, 1], [-1, 1]), 1) == ["\tTheta1 ~ normal(0,1);\n", "\treal<lower=-1, upper=1> Theta1;\n"]
    assert bp.parse_prior(bp.prior_dist("uniform", [0, 2], [0, 3]), 4) == ["\tTheta4 ~ uniform(0,2);\n", "\treal<lower=0, upper=3> Theta4;\n"]
    assert bp.parse_prior(bp.prior_dist("normal", [0, .4324]), 1) == ["\tTheta1 ~ normal(0,0.4324);\n", "\treal Theta1;\n"]
    assert bp.parse_prior(bp.prior_dist("gamma", [.5, 5.77], [0, 58.18]), 1) == ["\tTheta1 ~ gamma(0.5,5.77);\n", "\treal<lower=0, upper=58.18> Theta1;\n"]


def test_all_22_distributions_listed():
    assert len(bp.ALLOWED_DISTRIBUTIONS) == 22
    for s in bp.ALLOWED_DISTRIBUTIONS:
        name, n, cons = s.split()
        params = [1.5] * int(n)
        if name == "uniform":
            params = [0.0, 2.0]
        bp.prior_dist(name, params, [0.1, 0.9])


def test_generate_prior_count():
    """test_generation.R:3-13: wrong number of priors."""
    mod = bp.bp_model(E1, [1, 1], ['c[1]', 'c[2]'], 2, 0)
    for k in (1, 3):
        with pytest.raises(ValueError, match="incorrect number of priors for model!"):
            bp.stan_model(mod, [bp.prior_dist("uniform", [0, 1], [0, 2])] * k)
    with pytest.raises(ValueError, match="model is not a simple birth-death process!"):
        bp.stan_model(bp.bp_model(E2, P2, ['c[%d]' % k for k in range(1, 7)], 6, 0), [bp.prior_dist("uniform", [0, 1], [0, 2])] * 6, simple_bd=True)
    with pytest.raises(ValueError, match="observational noise models are not available for one-type processes"):
        bp.stan_model(mod, [bp.prior_dist("uniform", [0, 1], [0, 2])] * 2, noise_model=True)


def test_simulation_errors():
    """test_simulation.R:3-22."""
    mod = bp.bp_model(E1, [1, 1], ['c[1]', 'c[2]'], 2, 0)
    with pytest.raises(ValueError, match="Initial popualation vector has wrong size!"):
        bp.bpsims(mod, [0.25, 0.10], [1000, 500], np.arange(6), 5)
    with pytest.raises(ValueError, match="Wrong number of parameters were provided for the model!"):
        bp.bpsims(mod, [0.25, 0.10, .1, .5], [1000], np.arange(6), 5)
    mod = bp.bp_model(E1, [1, 1], ['c[1]*log(x[1])', 'c[2]'], 2, 1)
    with pytest.raises(ValueError, match="Incorrect number of dependent variables were provided for the model!"):
        bp.bpsims(mod, [0.25, 0.10], [1000], np.arange(6), 5)


def test_simulation_shapes_and_conversion():
    """test_simulation.R:24-107, 129-152 and test_conversion.R:3-23."""
    rng = np.random.default_rng(7)
    times = np.arange(6)
    mod = bp.bp_model(E1, [1, 1], ['c[1]', 'c[2]'], 2, 0)
    sim = bp.bpsims(mod, [0.25, 0.10], [1000], times, 5, rng=rng)
    assert sim.iloc[0, 2] == 1000 and len(sim) == len(times) * 5
    mod2 = bp.bp_model(E2, P2, ['c[%d]' % k for k in range(1, 7)], 6, 0)
    sim2 = bp.bpsims(mod2, [0.20, 0.10, 0.05, 0.20, 0.25, 0.20], [1000, 500], times, 5, rng=rng)
    assert sim2.iloc[0, 2] == 1000 and sim2.iloc[0, 3] == 500 and len(sim2) == 30
    dat2 = bp.stan_data_from_simulation(sim2, mod2)
    assert dat2["ndatapts"] == 25 and dat2["ntypes"] == 2 and dat2["nevents"] == 6 and dat2["ntimes_unique"] == 1
    assert dat2["function_var"].shape == (1, 1) and dat2["ndep"] == 1 and dat2["ndep_levels"] == 1 and (dat2["var_idx"] == 1).all()
    assert (dat2["init_pop"][0] == [1000, 500]).all() and dat2["pop_vec"].shape == (25, 2)
    assert (bp.bpsims(mod, [0.25, 0.10], [1000], [1], 5, rng=rng)["times"] == 1).all()
    # functional dependencies
    modl = bp.bp_model(E1, [1, 1], [LOGI, 'c[5]'], 5, 1)
    reps = 1 + rng.poisson(4, size=6)
    c_mat = np.array([0.0, 0.2, 0.4, 0.6, 0.8, 1.0]).reshape(-1, 1)
    siml = bp.bpsims(modl, [0.15, .2, 10, 0.5, 0.10], [1000], times, reps, c_mat, rng=rng)
    assert len(siml) == len(times) * reps.sum() and siml.iloc[0, 4] == 1000 and siml.iloc[0, 3] == 0.0 and siml.iloc[-1, 3] == 1.0
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        datl = bp.stan_data_from_simulation(siml, modl)
    assert datl["ndep_levels"] == 6 and datl["ndatapts"] == 5 * reps.sum()
    siml = bp.bpsims(modl, [0.15, .2, 10, 0.5, 0.10], [1000], np.arange(1, 16, 3), reps, c_mat, rng=rng)
    datl = bp.stan_data_from_simulation(siml, modl)
    assert len(datl["times"]) == 1 and datl["times"][0] == 3
    # two dependent variables, two types
    fd = [LOGI, 'c[5]', 'c[6]', 'c[7] + c[8]/(1 + exp(c[9]*(x[2] - c[10])))', 'c[11]', 'c[12]']
    mod12 = bp.bp_model(E2, P2, fd, 12, 2)
    cm2 = np.tile(c_mat, (1, 2))
    sim12 = bp.bpsims(mod12, [0.15, .2, 10, 0.5, 0.05, 0.05, 0.15, .2, 10, 0.5, 0.10, 0.10], [1000, 1000], times, reps, cm2, rng=rng)
    assert len(sim12) == len(times) * reps.sum()
    assert list(sim12.iloc[0, 3:7]) == [0.0, 0.0, 1000, 1000] and list(sim12.iloc[-1, 3:5]) == [1.0, 1.0]
    # conversion errors (test_conversion.R:18-21)
    mods = bp.bp_model_simple_birth_death(['c[1]', 'c[2]'], 2, 0)
    sims = bp.bpsims(mods, [0.1, 0.05], [1000], np.arange(1, 6), 25, rng=rng)
    with pytest.warns(UserWarning):
        dat = bp.stan_data_from_simulation(sims, mods, simple_bd=True) if False else bp.create_stan_data(mods, sims.iloc[1:, 2].to_numpy(), sims.iloc[:-1, 2].to_numpy(), np.ones(len(sims) - 1), simple_bd=True)
    msg = "times, initial populations, and final populations must all have same number of rows!"
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for a, b, c in [(dat["pop_vec"], dat["init_pop"], dat["times"][1:]), (dat["pop_vec"], dat["init_pop"][1:], dat["times"]),
                        (dat["pop_vec"][1:], dat["init_pop"], dat["times"])]:
            with pytest.raises(ValueError, match=msg):
                bp.create_stan_data(mods, a, b, c, simple_bd=True)
    with pytest.warns(UserWarning, match="final_pop is a vector, not a matrix. Converting to a one-column matrix"):
        bp.create_stan_data(mods, dat["pop_vec"], dat["init_pop"], dat["times"], simple_bd=True)
    assert set(dat) == {"pop_vec", "init_pop", "times", "function_var", "var_idx", "ndep_levels", "ndep", "nparams", "ndatapts"}


def test_calculate_moments_validation():
    """test_moments.R:3-29 (argument validation happens before any device work)."""
    em = np.array([[2.0], [0.0]])
    with pytest.raises(ValueError, match="Incorrect number of rate parameters provided for model!"):
        bp.calculate_moments(em, [1, 1], [0.25, 0.10, .5], [1000], 5)
    with pytest.raises(ValueError, match="Dimensions of birth events matrix disagree with dimension of initial population vector!"):
        bp.calculate_moments(em, [1, 1], [0.25, 0.10], [1000, 500], 5)
    with pytest.raises(ValueError, match="Dimensions of birth events matrix disagree with dimension of parent vector!"):
        bp.calculate_moments(em, [1, 1, 1], [0.25, 0.10], [1000], 5)
    par = "Parent vector must have integer entries between 1 and the number of types!"
    for pv in ([1, 2], [0, 2]):
        with pytest.raises(ValueError, match=par):
            bp.calculate_moments(em, pv, [0.25, 0.10], [1000], 5)
    em4 = np.array([[2, 0], [0, 0], [0, 0], [2, 0]], dtype=float).reshape(4, 2)
    with pytest.raises(ValueError, match=par):
        bp.calculate_moments(em4, [1, 2, 1.5, 2], [0.2, 0.20, .1, .1], [1000, 500], 5)


def test_uniform_initialize():
    init = bp.uniform_initialize(np.array([[0, 1], [5, 15]]), 3, rng=np.random.default_rng(0))
    assert len(init) == 3 and list(init[0]) == ["Theta1", "Theta2"] and 5 <= init[2]["Theta2"] <= 15
