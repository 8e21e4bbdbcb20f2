# parity_dump.R — for someone with R + rstan: dumps rstan's log_prob / grad_log_prob (and, for the two-type model, the per-observation
# log_lik of generate(..., loo = T)) at fixture points for the FIVE configurations of BASELINE.json, so that libbpcuda can be compared
# with rstan itself (closes the "parity unpinned" gap, SURVEY.md §8c / §8(f) N4; DESIGN.md §2).
#   Rscript parity_dump.R out.json          then:  python tools/compare_rstan_dump.py out.json
# Each entry holds everything the comparison needs: the model (e_mat, p_vec, func_deps as deparsed strings, priors, template
# switches), the Stan data list, the unconstrained fixture points (one column per point), and rstan's answers.
library(bpinference); library(rstan); library(jsonlite)
set.seed(1)

dump_one <- function(name, mod, priors, dat, upars, simple_bd = FALSE, noise_model = FALSE, loo = FALSE) {
  sm <- stan_model(model_code = generate(mod, priors, simple_bd = simple_bd, noise_model = noise_model, loo = loo))
  fit <- sampling(sm, data = dat, chains = 1, iter = 1, algorithm = "Fixed_param", refresh = 0)
  out <- list(name = name, simple_bd = simple_bd, noise_model = noise_model,
              model = list(e_mat = mod$e_mat, p_vec = mod$p_vec, func_deps = vapply(mod$func_deps, function(f) paste(deparse(f), collapse = ""), ""),
                           nparams = mod$nparams, ndep = mod$ndep),
              priors = lapply(priors, unclass), dat = dat, upars = upars,
              lp = apply(upars, 2, function(u) log_prob(fit, u, adjust_transform = TRUE)),
              lp_nojac = apply(upars, 2, function(u) log_prob(fit, u, adjust_transform = FALSE)),
              grad = apply(upars, 2, function(u) as.vector(grad_log_prob(fit, u, adjust_transform = TRUE))))
  if (loo) {   # generated quantities at the first fixture point: constrain_pars runs the whole program
    cp <- constrain_pars(fit, upars[, 1])
    out$log_lik <- cp$log_lik
  }
  out
}

unc <- function(theta, lo, hi) { z <- (theta - lo) / (hi - lo); log(z / (1 - z)) }
res <- list()

# cfg1: examples/one_type.R
mod1 <- bp_model_simple_birth_death(c("c[1]", "c[2]"), 2, 0)
pri1 <- rep(list(list(name = "uniform", params = c(0, 1), bounds = c(0, 2))), 2)
sim1 <- bpsims(mod1, c(0.1, 0.05), c(1000), seq(1, 20), 25)
dat1 <- stan_data_from_simulation(sim1, mod1, simple_bd = TRUE)
up1 <- sapply(1:16, function(i) unc(c(0.1, 0.05) * exp(rnorm(2, 0, 0.2)), 0, 2))
res$cfg1 <- dump_one("cfg1", mod1, pri1, dat1, up1, simple_bd = TRUE)

# cfg2: examples/two_type.R (with the LOO block)
e2 <- rbind(c(2, 0), c(1, 1), c(1, 1), c(0, 2), c(0, 0), c(0, 0)); p2 <- c(1, 1, 2, 2, 1, 2)
mod2 <- bp_model(e2, p2, sprintf("c[%d]", 1:6), 6, 0)
pri2 <- rep(list(list(name = "normal", params = c(0, .25), bounds = c(0, 5))), 6)
th2 <- c(0.20, 0.10, 0.05, 0.20, 0.25, 0.20)
dat2 <- stan_data_from_simulation(bpsims(mod2, th2, c(1000, 500), seq(0, 5), 50), mod2)
up2 <- sapply(1:16, function(i) unc(th2 * exp(rnorm(6, 0, 0.15)), 0, 5))
res$cfg2 <- dump_one("cfg2", mod2, pri2, dat2, up2, loo = TRUE)

# cfg3: examples/one_type_logistic_birth.R
mod3 <- bp_model(matrix(c(2, 0), ncol = 1), c(1, 1), c("c[1] + c[2]/(1 + exp(c[3]*(x[1] - c[4])))", "c[5]"), 5, 1)
pri3 <- rep(list(list(name = "normal", params = c(0, .25), bounds = c(0, 5))), 5)
pri3[[3]] <- list(name = "uniform", params = c(5, 15), bounds = c(5, 15))
th3 <- c(0.15, .2, 10, 0.5, 0.10)
c_mat <- matrix(c(0.0, 0.2, 0.4, 0.6, 0.8, 1.0), ncol = 1)
dat3 <- stan_data_from_simulation(bpsims(mod3, th3, c(1000), seq(0, 5), rep(20, 6), c_mat), mod3, simple_bd = TRUE)
lo3 <- c(0, 0, 5, 0, 0); hi3 <- c(5, 5, 15, 5, 5)
up3 <- sapply(1:16, function(i) unc(th3 * exp(rnorm(5, 0, 0.1)), lo3, hi3))
res$cfg3 <- dump_one("cfg3", mod3, pri3, dat3, up3, simple_bd = TRUE)

# cfg4: examples/apoptosis_with_noise.R
mod4 <- bp_model(rbind(c(2, 0), c(0, 1), c(0, 0)), c(1, 1, 2), c("c[1]", "c[2]", "c[3]"), 3, 0)
pri4 <- rep(list(list(name = "normal", params = c(0, .25), bounds = c(0, 2))), 3)
th4 <- c(0.08, 0.05, .1)
sim4 <- bpsims(mod4, th4, c(1000, 0), seq(0, 5), 50)
sim4$X1 <- sapply(sim4$X1, function(x) rbinom(1, x, .65)); sim4$X2 <- sapply(sim4$X2, function(x) rbinom(1, x, .65))
dat4 <- stan_data_from_simulation(sim4, mod4)
up4 <- sapply(1:16, function(i) c(unc(th4 * exp(rnorm(3, 0, 0.15)), 0, 2), log(0.05) + rnorm(1, 0, 0.2)))
res$cfg4 <- dump_one("cfg4", mod4, pri4, dat4, up4, noise_model = TRUE)

# cfg5: the synthetic 3-type process of SURVEY.md §8(d), reduced to what rstan's ODE route finishes in minutes: 6 durations x 40 replicates
e5 <- matrix(0, 9, 3); p5 <- rep(0, 9)
for (i in 1:3) { e5[3 * i - 2, i] <- 2; e5[3 * i - 1, i %% 3 + 1] <- 1; p5[(3 * i - 2):(3 * i)] <- i }
mod5 <- bp_model(e5, p5, sprintf("c[%d]", 1:9), 9, 0)
pri5 <- rep(list(list(name = "normal", params = c(0, .25), bounds = c(0, 5))), 9)
th5 <- c(0.25, 0.05, 0.10, 0.20, 0.04, 0.12, 0.22, 0.06, 0.09)
dat5 <- stan_data_from_simulation(bpsims(mod5, th5, c(2000, 1500, 1000), c(0, 0.5, 1.5, 3.5, 6.5, 10.5, 15.5), 40), mod5)
up5 <- sapply(1:16, function(i) unc(th5 * exp(rnorm(9, 0, 0.1)), 0, 5))
res$cfg5 <- dump_one("cfg5", mod5, pri5, dat5, up5)

write_json(res, commandArgs(trailingOnly = TRUE)[1], digits = NA, auto_unbox = TRUE)
