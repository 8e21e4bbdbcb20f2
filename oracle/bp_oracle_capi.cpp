// CPU ORACLE — C entry points for ctypes (TEST INFRASTRUCTURE ONLY, see bp_oracle.hpp header).
// Built by oracle/Makefile into oracle/libbporacle.so.  Used by tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs; never by the product path.
#include "bp_oracle.hpp"
#include <cstring>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace bpo;

namespace {

template <class T, int n>
int multitype_all(int E, const double* e_mat, const int* p_vec, int N, const double* pop_vec, const double* init_pop,
                  const int* var_idx, const int* times_idx, int L, int Tn, const double* times, int C,
                  const double* r, const double* sigma_obs, int grouped, double theta_step, int nthreads,
                  double* lp, double* rbar, double* sbar, int* status) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
  for (int c = 0; c < C; ++c) {
    std::vector<T> rc((size_t)L * E), rb((size_t)L * E, T(0.0));
    for (size_t i = 0; i < rc.size(); ++i) rc[i] = T(r[(size_t)c * L * E + i]);
    T lpc, sb;
    status[c] = multitype_chain<T, n>(E, e_mat, p_vec, N, pop_vec, init_pop, var_idx, times_idx, L, Tn, times,
                                       rc.data(), sigma_obs ? sigma_obs[c] : -1.0, grouped, theta_step, lpc, rb.data(), sb);
    lp[c] = val(lpc);
    for (size_t i = 0; i < rb.size(); ++i) rbar[(size_t)c * L * E + i] = val(rb[i]);
    if (sbar) sbar[c] = val(sb);
  }
  return 0;
}

template <class T>
int multitype_dispatch(int n, int E, const double* e_mat, const int* p_vec, int N, const double* pop_vec,
                       const double* init_pop, const int* var_idx, const int* times_idx, int L, int Tn,
                       const double* times, int C, const double* r, const double* sigma_obs, int grouped,
                       double theta_step, int nthreads, double* lp, double* rbar, double* sbar, int* status) {
#define BPO_CASE(NN) case NN: return multitype_all<T, NN>(E, e_mat, p_vec, N, pop_vec, init_pop, var_idx, times_idx, L, Tn, times, C, r, sigma_obs, grouped, theta_step, nthreads, lp, rbar, sbar, status)
  switch (n) { BPO_CASE(1); BPO_CASE(2); BPO_CASE(3); BPO_CASE(4); BPO_CASE(5); default: return -1; }
#undef BPO_CASE
}

}  // namespace

extern "C" {

// precision: 0 = double, 1 = long double.  Layouts: e_mat E x n col-major; pop_vec/init_pop N x n col-major;
// r, rbar [C][L][E]; sigma_obs [C] or NULL.  Returns 0 or -1 (unsupported ntypes).
int bpo_multitype_lp_grad(int precision, int n, int E, const double* e_mat, const int* p_vec, int N,
                          const double* pop_vec, const double* init_pop, const int* var_idx, const int* times_idx,
                          int L, int Tn, const double* times, int C, const double* r, const double* sigma_obs,
                          int grouped, double theta_step, int nthreads, double* lp, double* rbar, double* sbar,
                          int* status) {
  if (precision == 1)
    return multitype_dispatch<long double>(n, E, e_mat, p_vec, N, pop_vec, init_pop, var_idx, times_idx, L, Tn, times,
                                           C, r, sigma_obs, grouped, theta_step, nthreads, lp, rbar, sbar, status);
  return multitype_dispatch<double>(n, E, e_mat, p_vec, N, pop_vec, init_pop, var_idx, times_idx, L, Tn, times, C, r,
                                    sigma_obs, grouped, theta_step, nthreads, lp, rbar, sbar, status);
}

// Exact algorithmic operation count (SURVEY.md §8d convention) of ONE chain's evaluation of the
// likelihood part (generator assembly + moments + scores + adjoint), for the shipped algorithm.
long long bpo_multitype_flops(int n, int E, const double* e_mat, const int* p_vec, int N, const double* pop_vec,
                              const double* init_pop, const int* var_idx, const int* times_idx, int L, int Tn,
                              const double* times, const double* r, double sigma_obs, int grouped,
                              double theta_step) {
  std::vector<double> lp(1), rbar((size_t)L * E), sbar(1);
  std::vector<int> st(1);
  Counted::ops = 0;
  int rc = multitype_dispatch<Counted>(n, E, e_mat, p_vec, N, pop_vec, init_pop, var_idx, times_idx, L, Tn, times, 1, r,
                                       sigma_obs >= 0 ? &sigma_obs : nullptr, grouped, theta_step, 1, lp.data(),
                                       rbar.data(), sbar.data(), st.data());
  if (rc) return -1;
  return (long long)Counted::ops;
}

// one-type template.  r, rbar [C][L][2].
int bpo_onetype_lp_grad(int precision, int N, const double* pop_vec, const double* init_pop, const double* times,
                        const int* var_idx, int L, int C, const double* r, int nthreads, double* lp, double* rbar,
                        int* status) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
  for (int c = 0; c < C; ++c) {
    if (precision == 1) {
      std::vector<long double> rc(2 * (size_t)L), rb(2 * (size_t)L, 0.0L);
      for (size_t i = 0; i < rc.size(); ++i) rc[i] = r[(size_t)c * 2 * L + i];
      long double lpc;
      status[c] = onetype_chain<long double>(N, pop_vec, init_pop, times, var_idx, L, rc.data(), lpc, rb.data());
      lp[c] = (double)lpc;
      for (size_t i = 0; i < rb.size(); ++i) rbar[(size_t)c * 2 * L + i] = (double)rb[i];
    } else {
      std::vector<double> rb(2 * (size_t)L, 0.0);
      double lpc;
      status[c] = onetype_chain<double>(N, pop_vec, init_pop, times, var_idx, L, r + (size_t)c * 2 * L, lpc, rb.data());
      lp[c] = lpc;
      for (size_t i = 0; i < rb.size(); ++i) rbar[(size_t)c * 2 * L + i] = rb[i];
    }
  }
  return 0;
}

long long bpo_onetype_flops(int N, const double* pop_vec, const double* init_pop, const double* times,
                            const int* var_idx, int L, const double* r) {
  std::vector<Counted> rc(2 * (size_t)L), rb(2 * (size_t)L);
  for (size_t i = 0; i < rc.size(); ++i) rc[i] = Counted(r[i]);
  Counted lp;
  Counted::ops = 0;
  onetype_chain<Counted>(N, pop_vec, init_pop, times, var_idx, L, rc.data(), lp, rb.data());
  return (long long)Counted::ops;
}

// single-ancestor moments at one (rates, duration): M[n x n] row-major (M[l][j] = E[Z_j | one type-l
// ancestor]) and V[n][n][n] (V[l] = covariance matrix from one type-l ancestor).  For golden checks.
int bpo_moments(int n, int E, const double* e_mat, const int* p_vec, const double* r, double t, double theta_step,
                double* M, double* V) {
#define BPO_CASE(NN)                                                                       \
  case NN: {                                                                               \
    Gen<double, NN> g;                                                                     \
    assemble<double, NN>(e_mat, p_vec, E, r, g);                                           \
    Plan p = make_plan<double, NN>(g, t, theta_step);                                      \
    if (!p.ok) return -2;                                                                  \
    for (int l = 0; l < NN; ++l) {                                                         \
      State<double, NN> y;                                                                 \
      for (int i = 0; i < NN; ++i) { y.mu[i] = (i == l); for (int j = 0; j < NN; ++j) y.S[i][j] = 0; } \
      taylor_fwd<double, NN>(g, p, y, nullptr);                                            \
      for (int i = 0; i < NN; ++i) { M[l * NN + i] = y.mu[i]; for (int j = 0; j < NN; ++j) V[(l * NN + i) * NN + j] = y.S[i][j]; } \
    }                                                                                      \
    return 0;                                                                              \
  }
  switch (n) { BPO_CASE(1) BPO_CASE(2) BPO_CASE(3) BPO_CASE(4) BPO_CASE(5) default: return -1; }
#undef BPO_CASE
}

int bpo_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
