# bpcuda.R — R wrappers a bpinference maintainer adds next to R/stan_utils.R.  Same arguments as generate()
# (R/stan_utils.R:130) and the same `dat` list; every exported signature of the package stays as it is.
# All arithmetic happens in libbpcuda (include/bpcuda.h) through the .Call shim src/bpcuda_rcall.c.

#' compile a model for the GPU engine: the stand-in for rstan::stan_model(model_code = generate(model, priors, ...))
#' @export
bp_cuda_model <- function(model, priors, simple_bd = FALSE, noise_model = FALSE) {
  if (is.null(attr(model, "class")) || attr(model, "class") != "bp_model") stop("model must be a bp_model object!")
  m <- unclass(model)
  m$func_deps <- vapply(model$func_deps, function(f) paste(deparse(f), collapse = ""), "")
  kind <- if (simple_bd) 2L else if (noise_model) 1L else 0L
  structure(list(ptr = .Call("bpc_R_model_create", m, lapply(priors, unclass), kind), simple_bd = simple_bd,
                 nparams = model$nparams, noise = noise_model, ntypes = ncol(model$e_mat), nevents = nrow(model$e_mat),
                 e_mat = model$e_mat, p_vec = model$p_vec), class = "bp_cuda_model")
}

#' number of CUDA devices libbpcuda sees
#' @export
bp_cuda_devices <- function() .Call("bpc_R_device_count")

#' device copy of the Stan data list (create_stan_data / stan_data_from_simulation, R/stan_utils.R:12-112)
#' @export
bp_cuda_data <- function(cm, dat) structure(list(ptr = .Call("bpc_R_data_create", cm$ptr, dat, cm$simple_bd), dat = dat), class = "bp_cuda_data")

#' rstan::log_prob / rstan::grad_log_prob for one vector or a matrix with one column per chain
#' @export
bp_log_prob <- function(cm, cd, upars, adjust_transform = TRUE, gradient = FALSE) {
  r <- .Call("bpc_R_log_prob", cm$ptr, cd$ptr, as.matrix(upars), adjust_transform, gradient)
  lp <- r[[1]]
  lp[r[[3]] != 0L] <- -Inf
  if (gradient) attr(lp, "gradient") <- r[[2]]
  lp
}

#' per-observation log_lik (ndatapts x chains) of generate(..., loo = T): stanfiles/multitype_loo_block.stan, simple_birth_death_loo_block.stan
#' @export
bp_log_lik <- function(cm, cd, upars) .Call("bpc_R_log_lik", cm$ptr, cd$ptr, as.matrix(upars), as.integer(cd$dat$ndatapts))

#' mean vector and covariance matrix of every (chain, observation): calculate_moments (R/calculate_moments.R:11-88) on the device
#' @export
bp_moments <- function(cm, cd, upars) {
  r <- .Call("bpc_R_moments", cm$ptr, cd$ptr, as.matrix(upars), as.integer(cd$dat$ndatapts), as.integer(cm$ntypes))
  list(mu_mat = r[[1]], sigma_mat = r[[2]])
}

#' Gillespie replicates on the device: what bp() returns for every replicate (R/simulation.R:9-56), ntypes x ntimes x nrep
#' @export
bp_cuda_simulate <- function(e_mat, p_vec, r_vec, z0, times, nrep, seed = sample.int(.Machine$integer.max, 1), device = 0L)
  .Call("bpc_R_simulate", e_mat, as.integer(p_vec), r_vec, z0, times, as.integer(nrep), as.numeric(seed), as.integer(device))

# columns in rstan's order: Theta1..P (sigma_obs), r_mat column-major, theta[i, ] = c(e_mat, r_mat) with the first index fastest, lp__
# (multitype_template.stan:105-117; vignettes/two-type-inference.Rmd:63-69 reads samples[, 1:6])
bp_fit_matrix <- function(cm, cd, fit) {
  dim3 <- dim(fit$draws)
  flat <- matrix(fit$draws, nrow = dim3[1])                                  # dim x (ns * chains)
  L <- cd$dat$ndep_levels
  rates <- .Call("bpc_R_transformed", cm$ptr, cd$ptr, flat, as.integer(L), as.integer(if (cm$simple_bd) 2L else cm$nevents))
  pars <- t(flat)
  colnames(pars) <- dimnames(fit$draws)[[1]]
  if (cm$simple_bd) {
    tp <- t(matrix(rates, nrow = 2 * L))
    colnames(tp) <- as.vector(outer(1:2, 1:L, function(e, v) sprintf("r_mat[%d,%d]", e, v)))
  } else {
    E <- cm$nevents; n <- cm$ntypes; nd <- ncol(flat)
    rm <- array(0, c(E, n, L, nd))
    for (e in 1:E) rm[e, cm$p_vec[e], , ] <- rates[e, , ]
    r_last <- t(matrix(rm[, , L, ], nrow = E * n))
    colnames(r_last) <- as.vector(outer(1:E, 1:n, function(e, j) sprintf("r_mat[%d,%d]", e, j)))
    th <- array(0, c(L, 2 * E * n, nd))
    for (v in 1:L) { th[v, 1:(E * n), ] <- as.vector(cm$e_mat); th[v, (E * n + 1):(2 * E * n), ] <- matrix(rm[, , v, ], nrow = E * n) }
    thm <- t(matrix(th, nrow = L * 2 * E * n))
    colnames(thm) <- as.vector(outer(1:L, 1:(2 * E * n), function(v, q) sprintf("theta[%d,%d]", v, q)))
    tp <- cbind(r_last, thm)
  }
  cbind(pars, tp, lp__ = as.vector(fit$lp__))
}

#' rstan::sampling(stan_mod, data = dat, chains, iter, warmup, init, control = list(adapt_delta, max_treedepth)).
#' devices: 0-based device ids; more than one shards `chains` over them inside libbpcuda (bpc_sample_multi: one host thread per
#' device, ncclCommInitAll for the pooled warm-up statistics and R-hat) — the stand-in for options(mc.cores = ...).
#' @export
bp_sampling <- function(cm, cd, chains = 4, iter = 2000, warmup = floor(iter / 2), init = "random",
                        control = list(adapt_delta = 0.8, max_treedepth = 10), seed = sample.int(.Machine$integer.max, 1),
                        devices = 0L, pooled_adaptation = FALSE, init_opt_steps = 0L) {
  inits <- if (is.list(init)) sapply(init, function(ch) unlist(ch[sprintf("Theta%d", seq_len(cm$nparams))])) else NULL
  delta <- if (is.null(control$adapt_delta)) 0.8 else control$adapt_delta
  depth <- if (is.null(control$max_treedepth)) 10L else as.integer(control$max_treedepth)
  metric <- if (identical(control$metric, "dense_e")) 1L else 0L     # dense_e needs pooled_adaptation = TRUE
  if (length(devices) > 1) {
    if (chains %% length(devices) != 0) stop("chains must be a multiple of the number of devices")
    r <- .Call("bpc_R_sample_multi", cm$ptr, cd$dat, cm$simple_bd, as.integer(devices), as.integer(chains / length(devices)), as.integer(iter),
               as.integer(warmup), delta, depth, as.numeric(seed), inits, pooled_adaptation, as.integer(init_opt_steps), metric)
  } else {
    r <- .Call("bpc_R_sample", cm$ptr, cd$ptr, as.integer(chains), as.integer(iter), as.integer(warmup), delta, depth, as.numeric(seed), inits,
               pooled_adaptation, as.integer(init_opt_steps), metric)
  }
  names(r) <- c("draws", "lp__", "divergent__", "treedepth__", "n_leapfrog__", "accept_stat__", "stepsize__", "Rhat")
  dimnames(r$draws)[[1]] <- c(sprintf("Theta%d", seq_len(cm$nparams)), if (cm$noise) "sigma_obs")
  names(r$Rhat) <- dimnames(r$draws)[[1]]
  structure(r, class = "bp_cuda_fit")
}

#' as.matrix(fit) of a stanfit: draws x columns in rstan's column order
#' @export
as.matrix.bp_cuda_fit <- function(x, cm, cd, ...) bp_fit_matrix(cm, cd, x)

#' the reference's diagnostics (R/stan_utils.R:493-533) for a bp_cuda_fit
#' @export
check_div_cuda <- function(fit) sum(fit$divergent__)
#' @export
check_exceeded_treedepth_cuda <- function(fit, max_depth = 10) {
  n <- sum(fit$treedepth__ == max_depth); N <- length(fit$treedepth__)
  sprintf("%s of %s iterations saturated the maximum tree depth of %s (%s%%)", n, N, max_depth, 100 * n / N)
}
#' @export
check_rhat_cuda <- function(fit) {
  bad <- which((fit$Rhat > 1.1 | is.infinite(fit$Rhat)) & !is.nan(fit$Rhat))
  for (k in bad) print(sprintf("Rhat for parameter %s is %s!", names(fit$Rhat)[k], fit$Rhat[k]))
  if (length(bad) == 0) "Rhat looks reasonable for all parameters" else "Rhat above 1.1 indicates that the chains very likely have not mixed"
}
