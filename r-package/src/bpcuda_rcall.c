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

/* .Call("bpc_R_sample", model_ptr, data_ptr, chains, iter, warmup, adapt_delta, max_treedepth, seed, inits (nparams x chains or NULL)) */
SEXP bpc_R_sample(SEXP mptr, SEXP dptr, SEXP chains, SEXP iter, SEXP warmup, SEXP delta, SEXP depth, SEXP seed, SEXP inits) {
  bpc_model* m = (bpc_model*)R_ExternalPtrAddr(mptr);
  bpc_data* d = (bpc_data*)R_ExternalPtrAddr(dptr);
  int C = Rf_asInteger(chains), it = Rf_asInteger(iter), wu = Rf_asInteger(warmup), dim = bpc_model_dim(m), ns = it - wu;
  bpc_sample_args a = {C, it, wu, Rf_asReal(delta), Rf_asInteger(depth), (unsigned long long)Rf_asReal(seed), 0,
                       Rf_isNull(inits) ? NULL : REAL(inits), 0, 0.0};
  SEXP draws = PROTECT(Rf_alloc3DArray(REALSXP, dim, ns, C)), lp = PROTECT(Rf_allocMatrix(REALSXP, ns, C));
  SEXP div = PROTECT(Rf_allocMatrix(INTSXP, ns, C)), td = PROTECT(Rf_allocMatrix(INTSXP, ns, C)), nl = PROTECT(Rf_allocMatrix(INTSXP, ns, C));
  SEXP acc = PROTECT(Rf_allocMatrix(REALSXP, ns, C)), eps = PROTECT(Rf_allocMatrix(REALSXP, ns, C));
  int rc = bpc_sample(m, d, &a, REAL(draws), REAL(lp), INTEGER(div), INTEGER(td), INTEGER(nl), REAL(acc), REAL(eps));
  if (rc) { UNPROTECT(7); Rf_error("%s", bpc_last_error()); }
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 7));
  SEXP parts[7] = {draws, lp, div, td, nl, acc, eps};
  for (int i = 0; i < 7; ++i) SET_VECTOR_ELT(out, i, parts[i]);
  UNPROTECT(8);
  return out;
}

static const R_CallMethodDef call_methods[] = {
    {"bpc_R_model_create", (DL_FUNC)&bpc_R_model_create, 3}, {"bpc_R_data_create", (DL_FUNC)&bpc_R_data_create, 3},
    {"bpc_R_log_prob", (DL_FUNC)&bpc_R_log_prob, 5}, {"bpc_R_sample", (DL_FUNC)&bpc_R_sample, 9}, {NULL, NULL, 0}};

void R_init_bpinference(DllInfo* dll) {
  R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
