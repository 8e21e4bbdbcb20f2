"""per-chain means of the grouped three-type data set of tests/test_gpu_sampler_stats.py (why is split R-hat of Theta8/9 at 1.09?)"""
import numpy as np
import bpinference_b200 as bp
e = np.zeros((9, 3)); p = np.zeros(9, dtype=int)
for i in range(3):
    e[3 * i, i] = 2.0; e[3 * i + 1, (i + 1) % 3] = 1.0; p[3 * i: 3 * i + 3] = i + 1
th = np.array([0.25, 0.05, 0.10, 0.20, 0.04, 0.12, 0.22, 0.06, 0.09])
mod = bp.bp_model(e, p, ["c[%d]" % k for k in range(1, 10)], 9, 0)
sim = bp.bpsims(mod, th, [2000, 1500, 1000], np.array([0, 0.5, 1.5, 3.0, 5.0, 7.5, 10.0]), 400, device=0, seed=13)
dat = bp.stan_data_from_simulation(sim, mod)
eng = bp.stan_model(mod, [bp.prior_dist("normal", [0, 0.25], [0, 5])] * 9)
d = eng.data(dat)
init = bp.uniform_initialize(np.tile([0.0, 1.0], (9, 1)), 64, rng=np.random.default_rng(1))
fit = eng.sampling(d, chains=64, iter=600, warmup=300, init=init, control=dict(adapt_delta=0.8, metric="dense_e"), seed=1, pooled_adaptation=True, init_opt_steps=300)
print("rhat", fit.rhat.round(3), "depth", fit.treedepth__.mean(), "div", fit.divergent__.sum())
cm = fit.draws.mean(1); cs = fit.draws.std(1); lpm = fit.lp__.mean(1)
order = np.argsort(cm[:, 7])
for c in order:
    print("chain %2d  lp %.2f  th8 %.4f +- %.4f  th9 %.4f +- %.4f  th7 %.4f  div %d depth %.2f" % (c, lpm[c], cm[c, 7], cs[c, 7], cm[c, 8], cs[c, 8], cm[c, 6], fit.divergent__[c].sum(), fit.treedepth__[c].mean()))
print("ess", fit.ess_bulk().round(0))
