/* bpcuda_rcall.c — the `.Call` shim a bpinference maintainer adds under src/ of the R package.
 * It translates R objects to the plain structs of include/bpcuda.h and back; it contains no arithmetic.
 * Compiled only where R headers exist (not in the build image of libbpcuda):
 *     PKG_CPPFLAGS = -I<libbpcuda>/include      PKG_LIBS = -L<libbpcuda> -lbpcuda
 * NAMESPACE gains:  useDynLib(bpinference, .registration = TRUE)
 * Error policy: C handles are released before Rf_error is raised (no longjmp across live resources). */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <string.h>
#include "bpcuda.h"

static void model_finalizer(SEXP p) { bpc_model* m = (bpc_model*)R_ExternalPtrAddr(p); if (m) { bpc_model_free(m); R_ClearExternalPtr(p); } }
static void data_finalizer(SEXP p) { bpc_data* d = (bpc_data*)R_ExternalPtrAddr(p); if (d) { bpc_data_free(d); R_ClearExternalPtr(p); } }

static SEXP list_get(SEXP lst, const char* name) {
  SEXP names = Rf_getAttrib(lst, R_NamesSymbol);
  for (R_xlen_t i = 0; i < Rf_xlength(lst); ++i) if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) return VECTOR_ELT(lst, i);
  return R_NilValue;
}

/* .Call("bpc_R_model_create", model, priors, kind): model = bp_model list (R/bp_model.R:15-16) with func_deps already
 * deparsed to strings, priors = list of list(name, params, bounds) (R/stan_utils.R:467-469), kind 0/1/2 */
SEXP bpc_R_model_create(SEXP model, SEXP priors, SEXP kind) {
  SEXP e_mat = PROTECT(Rf_coerceVector(list_get(model, "e_mat"), REALSXP));
  SEXP p_vec = PROTECT(Rf_coerceVector(list_get(model, "p_vec"), INTSXP));
  SEXP fd = list_get(model, "func_deps");
  int E = Rf_nrows(e_mat), n = Rf_ncols(e_mat), P = Rf_asInteger(list_get(model, "nparams"));
  const char** fds = (const char**)R_alloc(E, sizeof(char*));
  for (int e = 0; e < E; ++e) fds[e] = CHAR(STRING_ELT(fd, e));
  bpc_prior* pr = (bpc_prior*)R_alloc(Rf_length(priors), sizeof(bpc_prior));
  for (int k = 0; k < Rf_length(priors); ++k) {
    SEXP pk = VECTOR_ELT(priors, k);
    SEXP par = PROTECT(Rf_coerceVector(list_get(pk, "params"), REALSXP));
    SEXP bnd = list_get(pk, "bounds");
    pr[k].name = CHAR(STRING_ELT(list_get(pk, "name"), 0));
    pr[k].nparams = Rf_length(par);
    for (int i = 0; i < 3; ++i) pr[k].params[i] = i < Rf_length(par) ? REAL(par)[i] : 0.0;
    pr[k].has_bounds = Rf_length(bnd) == 2 && !ISNA(Rf_asReal(bnd));
    if (pr[k].has_bounds) { SEXP b = PROTECT(Rf_coerceVector(bnd, REALSXP)); pr[k].lower = REAL(b)[0]; pr[k].upper = REAL(b)[1]; UNPROTECT(1); }
    UNPROTECT(1);
  }
  bpc_model_desc d = {n, E, P, Rf_asInteger(list_get(model, "ndep")), REAL(e_mat), INTEGER(p_vec), fds, pr, Rf_asInteger(kind)};
  bpc_model* m = NULL;
  int rc = Rf_length(priors) == P ? bpc_model_create(&d, &m) : -1;
  UNPROTECT(2);
  if (rc == -1) Rf_error("incorrect number of priors for model!");
  if (rc) Rf_error("%s", bpc_last_error());
  SEXP ptr = PROTECT(R_MakeExternalPtr(m, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(ptr, model_finalizer, TRUE);
  UNPROTECT(1);
  return ptr;
}

/* .Call("bpc_R_data_create", model_ptr, dat, simple_bd): dat = the list create_stan_data returns (R/stan_utils.R:57-62) */
SEXP bpc_R_data_create(SEXP mptr, SEXP dat, SEXP simple) {
  bpc_model* m = (bpc_model*)R_ExternalPtrAddr(mptr);
  if (!m) Rf_error("model handle is NULL (an external pointer does not survive saveRDS or a new session: create it again)");
  int sb = Rf_asLogical(simple);
  SEXP pop = PROTECT(Rf_coerceVector(list_get(dat, "pop_vec"), REALSXP)), ini = PROTECT(Rf_coerceVector(list_get(dat, "init_pop"), REALSXP));
  SEXP tim = PROTECT(Rf_coerceVector(list_get(dat, "times"), REALSXP)), fv = PROTECT(Rf_coerceVector(list_get(dat, "function_var"), REALSXP));
  SEXP vi = PROTECT(Rf_coerceVector(list_get(dat, "var_idx"), INTSXP));
  SEXP ti = sb ? R_NilValue : PROTECT(Rf_coerceVector(list_get(dat, "times_idx"), INTSXP));
  bpc_stan_data sd = {Rf_asInteger(list_get(dat, "ndatapts")), sb ? 0 : Rf_length(tim), Rf_asInteger(list_get(dat, "ndep_levels")),
                      Rf_asInteger(list_get(dat, "ndep")), REAL(pop), REAL(ini), REAL(tim), sb ? NULL : INTEGER(ti), REAL(fv), INTEGER(vi)};
  bpc_data* d = NULL;
  int rc = bpc_data_create(m, &sd, 0, 0, &d);
  UNPROTECT(sb ? 5 : 6);
  if (rc) Rf_error("%s", bpc_last_error());
  SEXP ptr = PROTECT(R_MakeExternalPtr(d, R_NilValue, mptr));   /* keeps the model alive */
  R_RegisterCFinalizerEx(ptr, data_finalizer, TRUE);
  UNPROTECT(1);
  return ptr;
}

/* .Call("bpc_R_log_prob", model_ptr, data_ptr, upars (dim x C matrix: one column per chain), adjust_transform, gradient) */
SEXP bpc_R_log_prob(SEXP mptr, SEXP dptr, SEXP upars, SEXP jac, SEXP want_grad) {
  bpc_model* m = (bpc_model*)R_ExternalPtrAddr(mptr);
  bpc_data* d = (bpc_data*)R_ExternalPtrAddr(dptr);
  if (!m || !d) Rf_error("model / data handle is NULL (an external pointer does not survive saveRDS or a new session: create it again)");
  int dim = bpc_model_dim(m), C = Rf_length(upars) / dim, g = Rf_asLogical(want_grad);
  SEXP u = PROTECT(Rf_coerceVector(upars, REALSXP));
  SEXP lp = PROTECT(Rf_allocVector(REALSXP, C)), grad = PROTECT(Rf_allocMatrix(REALSXP, dim, g ? C : 0)), st = PROTECT(Rf_allocVector(INTSXP, C));
  int rc = bpc_log_prob(m, d, REAL(u), C, Rf_asLogical(jac), REAL(lp), g ? REAL(grad) : NULL, INTEGER(st));
  if (rc) { UNPROTECT(4); Rf_error("%s", bpc_last_error()); }
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
  SET_VECTOR_ELT(out, 0, lp); SET_VECTOR_ELT(out, 1, grad); SET_VECTOR_ELT(out, 2, st);
  UNPROTECT(5);
  return out;
}

/* the data list of R/stan_utils.R:57-62 as a bpc_stan_data; PROTECTs *nprot objects that must stay alive while `sd` is used */
static bpc_stan_data stan_data_of(SEXP dat, int sb, int* nprot) {
  SEXP pop = PROTECT(Rf_coerceVector(list_get(dat, "pop_vec"), REALSXP)), ini = PROTECT(Rf_coerceVector(list_get(dat, "init_pop"), REALSXP));
  SEXP tim = PROTECT(Rf_coerceVector(list_get(dat, "times"), REALSXP)), fv = PROTECT(Rf_coerceVector(list_get(dat, "function_var"), REALSXP));
  SEXP vi = PROTECT(Rf_coerceVector(list_get(dat, "var_idx"), INTSXP));
  SEXP ti = sb ? R_NilValue : PROTECT(Rf_coerceVector(list_get(dat, "times_idx"), INTSXP));
  *nprot = sb ? 5 : 6;
  bpc_stan_data sd = {Rf_asInteger(list_get(dat, "ndatapts")), sb ? 0 : Rf_length(tim), Rf_asInteger(list_get(dat, "ndep_levels")),
                      Rf_asInteger(list_get(dat, "ndep")), REAL(pop), REAL(ini), REAL(tim), sb ? NULL : INTEGER(ti), REAL(fv), INTEGER(vi)};
  return sd;
}

static SEXP fit_list(SEXP draws, SEXP lp, SEXP div, SEXP td, SEXP nl, SEXP acc, SEXP eps, SEXP rhat) {
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 8));
  SEXP parts[8] = {draws, lp, div, td, nl, acc, eps, rhat};
  for (int i = 0; i < 8; ++i) SET_VECTOR_ELT(out, i, parts[i]);
  UNPROTECT(1);
  return out;
}

/* .Call("bpc_R_sample", model_ptr, data_ptr, chains, iter, warmup, adapt_delta, max_treedepth, seed, inits (nparams x chains or NULL),
 *       pooled_adaptation, init_opt_steps, metric (0 diag_e, 1 dense_e)): one device */
SEXP bpc_R_sample(SEXP mptr, SEXP dptr, SEXP chains, SEXP iter, SEXP warmup, SEXP delta, SEXP depth, SEXP seed, SEXP inits, SEXP pooled, SEXP optsteps, SEXP metric) {
  bpc_model* m = (bpc_model*)R_ExternalPtrAddr(mptr);
  bpc_data* d = (bpc_data*)R_ExternalPtrAddr(dptr);
  if (!m || !d) Rf_error("model / data handle is NULL (an external pointer does not survive saveRDS or a new session: create it again)");
  int C = Rf_asInteger(chains), it = Rf_asInteger(iter), wu = Rf_asInteger(warmup), dim = bpc_model_dim(m), ns = it - wu;
  if (C < 1 || ns < 1) Rf_error("need chains >= 1 and warmup < iter");
  bpc_sample_args a = {C, it, wu, Rf_asReal(delta), Rf_asInteger(depth), (unsigned long long)Rf_asReal(seed), 0,
                       Rf_isNull(inits) ? NULL : REAL(inits), Rf_asLogical(pooled), 0.0, Rf_asInteger(optsteps), Rf_asInteger(metric), NULL};
  SEXP draws = PROTECT(Rf_alloc3DArray(REALSXP, dim, ns, C)), lp = PROTECT(Rf_allocMatrix(REALSXP, ns, C));
  SEXP div = PROTECT(Rf_allocMatrix(INTSXP, ns, C)), td = PROTECT(Rf_allocMatrix(INTSXP, ns, C)), nl = PROTECT(Rf_allocMatrix(INTSXP, ns, C));
  SEXP acc = PROTECT(Rf_allocMatrix(REALSXP, ns, C)), eps = PROTECT(Rf_allocMatrix(REALSXP, ns, C)), rhat = PROTECT(Rf_allocVector(REALSXP, dim));
  /* (bpc_sample with R-hat: the sampler object is driven here so that the moments can be read before it is freed) */
  bpc_sampler* s = NULL;
  int rc = bpc_sampler_create(m, d, &a, &s), state = 0;
  while (!rc && state != 1) {
    rc = bpc_sampler_run(s, 1LL << 40, &state);
    if (!rc && state == 2) { rc = bpc_sampler_pool_stats(s, NULL, 0); if (!rc) rc = bpc_sampler_pool_apply(s, NULL); }
  }
  if (!rc) rc = bpc_sampler_draws(s, REAL(draws), REAL(lp), INTEGER(div), INTEGER(td), INTEGER(nl), REAL(acc), REAL(eps));
  if (!rc) {
    double* mom = (double*)R_alloc(4 * dim, sizeof(double));
    rc = bpc_sampler_chain_moments(s, NULL, 0, mom);
    if (!rc && ns >= 2) rc = bpc_rhat(mom, dim, ns, REAL(rhat));
  }
  if (s) bpc_sampler_free(s);
  if (rc) { UNPROTECT(8); Rf_error("%s", bpc_last_error()); }
  SEXP out = PROTECT(fit_list(draws, lp, div, td, nl, acc, eps, rhat));
  UNPROTECT(9);
  return out;
}

/* .Call("bpc_R_sample_multi", model_ptr, dat, simple_bd, devices (0-based integer vector), chains PER DEVICE, iter, warmup, adapt_delta,
 *       max_treedepth, seed, inits (nparams x total chains or NULL), pooled_adaptation, init_opt_steps, metric): bpc_sample_multi — the chains
 * are sharded over the devices inside libbpcuda (one host thread per device, ncclCommInitAll); R stays single-threaded. */
SEXP bpc_R_sample_multi(SEXP mptr, SEXP dat, SEXP simple, SEXP devices, SEXP chains, SEXP iter, SEXP warmup, SEXP delta, SEXP depth, SEXP seed,
                        SEXP inits, SEXP pooled, SEXP optsteps, SEXP metric) {
  bpc_model* m = (bpc_model*)R_ExternalPtrAddr(mptr);
  if (!m) Rf_error("model handle is NULL (an external pointer does not survive saveRDS or a new session: create it again)");
  SEXP dev = PROTECT(Rf_coerceVector(devices, INTSXP));
  int ndev = Rf_length(dev), C = Rf_asInteger(chains), it = Rf_asInteger(iter), wu = Rf_asInteger(warmup), dim = bpc_model_dim(m), ns = it - wu;
  if (ndev < 1 || C < 1 || ns < 1) { UNPROTECT(1); Rf_error("need at least one device, chains >= 1 and warmup < iter"); }
  int T = C * ndev, nprot = 0;
  bpc_stan_data sd = stan_data_of(dat, Rf_asLogical(simple), &nprot);
  bpc_sample_args a = {C, it, wu, Rf_asReal(delta), Rf_asInteger(depth), (unsigned long long)Rf_asReal(seed), 0,
                       Rf_isNull(inits) ? NULL : REAL(inits), Rf_asLogical(pooled), 0.0, Rf_asInteger(optsteps), Rf_asInteger(metric), NULL};
  SEXP draws = PROTECT(Rf_alloc3DArray(REALSXP, dim, ns, T)), lp = PROTECT(Rf_allocMatrix(REALSXP, ns, T));
  SEXP div = PROTECT(Rf_allocMatrix(INTSXP, ns, T)), td = PROTECT(Rf_allocMatrix(INTSXP, ns, T)), nl = PROTECT(Rf_allocMatrix(INTSXP, ns, T));
  SEXP acc = PROTECT(Rf_allocMatrix(REALSXP, ns, T)), eps = PROTECT(Rf_allocMatrix(REALSXP, ns, T)), rhat = PROTECT(Rf_allocVector(REALSXP, dim));
  int rc = bpc_sample_multi(m, &sd, INTEGER(dev), ndev, &a, REAL(draws), REAL(lp), INTEGER(div), INTEGER(td), INTEGER(nl), REAL(acc), REAL(eps),
                            REAL(rhat), NULL);
  if (rc) { UNPROTECT(9 + nprot); Rf_error("%s", bpc_last_error()); }
  SEXP out = PROTECT(fit_list(draws, lp, div, td, nl, acc, eps, rhat));
  UNPROTECT(10 + nprot);
  return out;
}

/* .Call("bpc_R_log_lik", model_ptr, data_ptr, upars (dim x C)) -> ndatapts x C matrix: the log_lik of generate(..., loo = T) */
SEXP bpc_R_log_lik(SEXP mptr, SEXP dptr, SEXP upars, SEXP ndatapts) {
  bpc_model* m = (bpc_model*)R_ExternalPtrAddr(mptr);
  bpc_data* d = (bpc_data*)R_ExternalPtrAddr(dptr);
  if (!m || !d) Rf_error("model / data handle is NULL");
  int dim = bpc_model_dim(m), C = Rf_length(upars) / dim, N = Rf_asInteger(ndatapts);
  SEXP u = PROTECT(Rf_coerceVector(upars, REALSXP)), ll = PROTECT(Rf_allocMatrix(REALSXP, N, C));
  int rc = bpc_log_lik(m, d, REAL(u), C, REAL(ll));
  UNPROTECT(2);
  if (rc) Rf_error("%s", bpc_last_error());
  return ll;
}

/* .Call("bpc_R_moments", model_ptr, data_ptr, upars (dim x C), ndatapts, ntypes) -> list(mu [ntypes x N x C], sigma [ntypes x ntypes x N x C])
 * (R/calculate_moments.R:79-86 for every chain and observation; sigma is symmetric, so R's column-major reading of the row-major block is the same matrix) */
SEXP bpc_R_moments(SEXP mptr, SEXP dptr, SEXP upars, SEXP ndatapts, SEXP ntypes) {
  bpc_model* m = (bpc_model*)R_ExternalPtrAddr(mptr);
  bpc_data* d = (bpc_data*)R_ExternalPtrAddr(dptr);
  if (!m || !d) Rf_error("model / data handle is NULL");
  int dim = bpc_model_dim(m), C = Rf_length(upars) / dim, N = Rf_asInteger(ndatapts), n = Rf_asInteger(ntypes);
  SEXP u = PROTECT(Rf_coerceVector(upars, REALSXP));
  SEXP mu = PROTECT(Rf_alloc3DArray(REALSXP, n, N, C)), sg = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)n * n * N * C));
  int rc = bpc_moments(m, d, REAL(u), C, REAL(mu), REAL(sg));
  if (rc) { UNPROTECT(3); Rf_error("%s", bpc_last_error()); }
  SEXP dims = PROTECT(Rf_allocVector(INTSXP, 4));
  INTEGER(dims)[0] = n; INTEGER(dims)[1] = n; INTEGER(dims)[2] = N; INTEGER(dims)[3] = C;
  Rf_setAttrib(sg, R_DimSymbol, dims);
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
  SET_VECTOR_ELT(out, 0, mu); SET_VECTOR_ELT(out, 1, sg);
  UNPROTECT(5);
  return out;
}

/* .Call("bpc_R_transformed", model_ptr, data_ptr, theta (dim x ndraws, CONSTRAINED), ndep_levels, nevents) -> nevents x ndep_levels x ndraws rates */
SEXP bpc_R_transformed(SEXP mptr, SEXP dptr, SEXP theta, SEXP nlev, SEXP nev) {
  bpc_model* m = (bpc_model*)R_ExternalPtrAddr(mptr);
  bpc_data* d = (bpc_data*)R_ExternalPtrAddr(dptr);
  if (!m || !d) Rf_error("model / data handle is NULL");
  int dim = bpc_model_dim(m), nd = Rf_length(theta) / dim;
  SEXP th = PROTECT(Rf_coerceVector(theta, REALSXP)), out = PROTECT(Rf_alloc3DArray(REALSXP, Rf_asInteger(nev), Rf_asInteger(nlev), nd));
  int rc = bpc_transformed_parameters(m, d, REAL(th), nd, REAL(out));
  UNPROTECT(2);
  if (rc) Rf_error("%s", bpc_last_error());
  return out;
}

/* .Call("bpc_R_simulate", e_mat, p_vec, r_vec, z0, times, nrep, seed, device) -> ntypes x ntimes x nrep populations (bp()/bpsims(), R/simulation.R:9-56) */
SEXP bpc_R_simulate(SEXP e_mat, SEXP p_vec, SEXP r_vec, SEXP z0, SEXP times, SEXP nrep, SEXP seed, SEXP device) {
  SEXP em = PROTECT(Rf_coerceVector(e_mat, REALSXP)), pv = PROTECT(Rf_coerceVector(p_vec, INTSXP)), rv = PROTECT(Rf_coerceVector(r_vec, REALSXP));
  SEXP z = PROTECT(Rf_coerceVector(z0, REALSXP)), tm = PROTECT(Rf_coerceVector(times, REALSXP));
  int E = Rf_nrows(em), n = Rf_ncols(em), nt = Rf_length(tm), R = Rf_asInteger(nrep);
  SEXP out = PROTECT(Rf_alloc3DArray(REALSXP, n, nt, R));
  int rc = bpc_simulate(n, E, REAL(em), INTEGER(pv), REAL(rv), REAL(z), REAL(tm), nt, R, (unsigned long long)Rf_asReal(seed), Rf_asInteger(device), REAL(out));
  UNPROTECT(6);
  if (rc) Rf_error("%s", bpc_last_error());
  return out;
}

SEXP bpc_R_device_count(void) { return Rf_ScalarInteger(bpc_device_count()); }

static const R_CallMethodDef call_methods[] = {
    {"bpc_R_model_create", (DL_FUNC)&bpc_R_model_create, 3}, {"bpc_R_data_create", (DL_FUNC)&bpc_R_data_create, 3},
    {"bpc_R_log_prob", (DL_FUNC)&bpc_R_log_prob, 5}, {"bpc_R_sample", (DL_FUNC)&bpc_R_sample, 12},
    {"bpc_R_sample_multi", (DL_FUNC)&bpc_R_sample_multi, 14}, {"bpc_R_log_lik", (DL_FUNC)&bpc_R_log_lik, 4},
    {"bpc_R_moments", (DL_FUNC)&bpc_R_moments, 5}, {"bpc_R_transformed", (DL_FUNC)&bpc_R_transformed, 5},
    {"bpc_R_simulate", (DL_FUNC)&bpc_R_simulate, 8}, {"bpc_R_device_count", (DL_FUNC)&bpc_R_device_count, 0}, {NULL, NULL, 0}};

void R_init_bpinference(DllInfo* dll) {
  R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
