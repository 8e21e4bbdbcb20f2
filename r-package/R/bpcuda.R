# bpcuda.R — R wrappers a bpinference maintainer adds next to R/stan_utils.R.  Same arguments as generate()
# (R/stan_utils.R:130) and the same `dat` list; every exported signature of the package stays as it is.

#' compile a model for the GPU engine: the stand-in for rstan::stan_model(model_code = generate(model, priors, ...))
#' @export
bp_cuda_model <- function(model, priors, simple_bd = FALSE, noise_model = FALSE) {
  if (is.null(attr(model, "class")) || attr(model, "class") != "bp_model") stop("model must be a bp_model object!")
  m <- unclass(model)
  m$func_deps <- vapply(model$func_deps, function(f) paste(deparse(f), collapse = ""), "")
  kind <- if (simple_bd) 2L else if (noise_model) 1L else 0L
  structure(list(ptr = .Call("bpc_R_model_create", m, lapply(priors, unclass), kind), simple_bd = simple_bd,
                 nparams = model$nparams, noise = noise_model), class = "bp_cuda_model")
}

#' @export
bp_cuda_data <- function(cm, dat) .Call("bpc_R_data_create", cm$ptr, dat, cm$simple_bd)

#' rstan::log_prob / rstan::grad_log_prob for one vector or a matrix with one column per chain
#' @export
bp_log_prob <- function(cm, cd, upars, adjust_transform = TRUE, gradient = FALSE) {
  r <- .Call("bpc_R_log_prob", cm$ptr, cd, as.matrix(upars), adjust_transform, gradient)
  lp <- r[[1]]
  lp[r[[3]] != 0L] <- -Inf
  if (gradient) attr(lp, "gradient") <- r[[2]]
  lp
}

#' rstan::sampling(stan_mod, data = dat, chains, iter, warmup, init, control = list(adapt_delta, max_treedepth))
#' @export
bp_sampling <- function(cm, cd, chains = 4, iter = 2000, warmup = floor(iter / 2), init = "random",
                        control = list(adapt_delta = 0.8, max_treedepth = 10), seed = sample.int(.Machine$integer.max, 1)) {
  inits <- if (is.list(init)) sapply(init, function(ch) unlist(ch[sprintf("Theta%d", seq_len(cm$nparams))])) else NULL
  r <- .Call("bpc_R_sample", cm$ptr, cd, as.integer(chains), as.integer(iter), as.integer(warmup),
             if (is.null(control$adapt_delta)) 0.8 else control$adapt_delta,
             if (is.null(control$max_treedepth)) 10L else as.integer(control$max_treedepth), as.numeric(seed), inits)
  names(r) <- c("draws", "lp__", "divergent__", "treedepth__", "n_leapfrog__", "accept_stat__", "stepsize__")
  dimnames(r$draws)[[1]] <- c(sprintf("Theta%d", seq_len(cm$nparams)), if (cm$noise) "sigma_obs")
  structure(r, class = "bp_cuda_fit")
}
