# parity_dump.R — for someone with R + rstan: dumps rstan's log_prob / grad_log_prob at fixture points so that
# libbpcuda can be compared with rstan itself (closes the "parity unpinned" gap, SURVEY.md §8c / DESIGN.md §2).
#   Rscript parity_dump.R out.json          then:  python tools/compare_rstan_dump.py out.json
library(bpinference); library(rstan); library(jsonlite)
set.seed(1)
e_mat <- rbind(c(2, 0), c(1, 1), c(1, 1), c(0, 2), c(0, 0), c(0, 0)); p_vec <- c(1, 1, 2, 2, 1, 2)
mod <- bp_model(e_mat, p_vec, sprintf("c[%d]", 1:6), 6, 0)
priors <- rep(list(list(name = "normal", params = c(0, .25), bounds = c(0, 5))), 6)
sim <- bpsims(mod, c(0.20, 0.10, 0.05, 0.20, 0.25, 0.20), c(1000, 500), seq(0, 5), 50)
dat <- stan_data_from_simulation(sim, mod)
sm <- stan_model(model_code = generate(mod, priors))
fit <- sampling(sm, data = dat, chains = 1, iter = 1, algorithm = "Fixed_param", refresh = 0)
upars <- matrix(rnorm(6 * 16, -3, 0.3), nrow = 6)
out <- list(dat = dat, upars = upars, lp = apply(upars, 2, function(u) log_prob(fit, u, adjust_transform = TRUE)),
            grad = apply(upars, 2, function(u) grad_log_prob(fit, u, adjust_transform = TRUE)))
write_json(out, commandArgs(trailingOnly = TRUE)[1], digits = NA, auto_unbox = TRUE)
