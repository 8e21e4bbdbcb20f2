// =============================================================================
// bp_oracle.hpp — CPU ORACLE (TEST INFRASTRUCTURE ONLY — NOT PRODUCT CODE).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may build, load or call anything under oracle/.  The product
// path (bpinference_b200/, libbpcuda.so) never links or imports this file.
//
// What it restates (reference = jproney/bpinference, paths under /root/reference):
//   * generator assembly           stanfiles/multitype_template.stan:23-43
//                                  (R twin: R/calculate_moments.R:38-55)
//   * first/second moment system   stanfiles/multitype_template.stan:50-58, 85-94
//                                  (R twin: R/calculate_moments.R:46, 57-77)
//   * per-observation mu / Sigma   stanfiles/multitype_template.stan:137-149
//   * MVN score                    stanfiles/multitype_template.stan:152
//   * noise variant                stanfiles/multitype_noise_template.stan:153
//   * one-type closed form         stanfiles/one_type_template.stan:37-46
//
// The reference hands the moment ODE to Stan's integrate_ode_rk45 (third-party,
// not vendored: rstan 2.19.2 / StanHeaders, banner in vignettes/*.pdf).  This
// oracle evaluates the SAME linear ODE in closed form: the state (mu, Sigma) of
// one observation obeys  mu' = A^T mu,  Sigma' = A^T Sigma + Sigma A + sum_i mu_i H_i
// (forward Kolmogorov form of template lines 55-58, with
//  H_l = C_l + diag(a_l) - a_l e_l^T - e_l a_l^T, a_l = row l of A), and the
// action of exp(t L) on (z0, 0) is summed as a Taylor series in steps of bounded
// norm.  oracle/stan_template_ref.py integrates the template's own ODE (RK45 and
// a tight-tolerance solver) and tests/ pin this file against it, against
// scipy.linalg.expm of the Van Loan block matrix and against the vignette
// values.  PARITY UNPINNED against rstan itself: no R / rstan exists in this
// image and the reference's tests hold no log_prob / grad_log_prob vectors.
// =============================================================================
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>
#include <limits>

namespace bpo {

// ---------------------------------------------------------------- flop counter
// Counted<double>: every +,-,*,/ and sqrt/exp/log counts 1 (an FMA therefore 2),
// the convention of SURVEY.md §8(d).
struct Counted {
  double v;
  static thread_local uint64_t ops;
  Counted() : v(0) {}
  Counted(double x) : v(x) {}
  explicit operator double() const { return v; }
};
inline thread_local uint64_t Counted::ops = 0;
inline Counted operator+(Counted a, Counted b) { ++Counted::ops; return Counted(a.v + b.v); }
inline Counted operator-(Counted a, Counted b) { ++Counted::ops; return Counted(a.v - b.v); }
inline Counted operator*(Counted a, Counted b) { ++Counted::ops; return Counted(a.v * b.v); }
inline Counted operator/(Counted a, Counted b) { ++Counted::ops; return Counted(a.v / b.v); }
inline Counted operator-(Counted a) { return Counted(-a.v); }
inline Counted& operator+=(Counted& a, Counted b) { a = a + b; return a; }
inline Counted& operator-=(Counted& a, Counted b) { a = a - b; return a; }
inline Counted& operator*=(Counted& a, Counted b) { a = a * b; return a; }
inline bool operator>(Counted a, Counted b) { return a.v > b.v; }
inline bool operator<(Counted a, Counted b) { return a.v < b.v; }
inline Counted sqrt(Counted a) { ++Counted::ops; return Counted(std::sqrt(a.v)); }
inline Counted exp(Counted a) { ++Counted::ops; return Counted(std::exp(a.v)); }
inline Counted log(Counted a) { ++Counted::ops; return Counted(std::log(a.v)); }
inline double val(Counted a) { return a.v; }
inline double val(double a) { return a; }
inline double val(long double a) { return (double)a; }
using std::sqrt; using std::exp; using std::log;

// status bits (per chain), mirrored by include/bpcuda.h
enum : int { ST_OK = 0, ST_NOT_PD = 1, ST_RATE_NONFINITE = 2, ST_ONETYPE_SINGULAR = 4, ST_PRIOR_SUPPORT = 8 };

// ---------------------------------------------------------------- Taylor plan
// Series degree for a step of scaled norm theta: smallest m with theta^m/m! <= 2^-55,
// plus one term for the (single) coupling block H.  theta_step bounds theta per step.
constexpr int M_CAP = 48;
struct DegreeTable {
  double th[M_CAP + 1];
  DegreeTable() {
    const double logtol = -55.0 * std::log(2.0);
    th[0] = 0.0;
    for (int m = 1; m <= M_CAP; ++m) th[m] = std::exp((logtol + std::lgamma(m + 1.0)) / m);
  }
  int degree(double theta) const {
    for (int m = 1; m < M_CAP; ++m) if (theta <= th[m]) return m + 1;
    return M_CAP;
  }
};
inline const DegreeTable& degree_table() { static DegreeTable t; return t; }

struct Plan { int s; int m; double h; bool ok; };

template <class T, int n>
struct Gen { T A[n][n]; T H[n][n][n]; };   // H[l] kept as a full symmetric n x n

// max column abs-sum of A (||A||_1); the Lyapunov part of L is bounded by twice that
template <class T, int n>
inline Plan make_plan(const Gen<T, n>& g, double t, double theta_step) {
  double a1 = 0;
  for (int j = 0; j < n; ++j) { double s = 0; for (int i = 0; i < n; ++i) s += std::fabs(val(g.A[i][j])); if (s > a1) a1 = s; }
  double x = 2.0 * a1 * t;
  Plan p; p.ok = std::isfinite(x) && t >= 0;
  if (!p.ok) { p.s = 1; p.m = 1; p.h = 0; return p; }
  p.s = (int)std::ceil(x / theta_step); if (p.s < 1) p.s = 1;
  p.h = t / p.s;
  p.m = degree_table().degree(x / p.s);
  return p;
}

// ---------------------------------------------------------------- generator
// stanfiles/multitype_template.stan:23-43.  e_mat is E x n column-major (R layout),
// p_vec 1-based.  A = B - diag(lambda); C_i[a,b] = sum_{e:p_e=i} r_e (E[e,a]E[e,b] - d_ab E[e,a]).
template <class T, int n>
inline void assemble(const double* e_mat, const int* p_vec, int E, const T* r, Gen<T, n>& g) {
  T C[n][n][n];
  for (int i = 0; i < n; ++i) for (int a = 0; a < n; ++a) { g.A[i][a] = T(0.0); for (int b = 0; b < n; ++b) C[i][a][b] = T(0.0); }
  for (int e = 0; e < E; ++e) {
    const int i = p_vec[e] - 1;
    for (int a = 0; a < n; ++a) {
      const double ea = e_mat[e + E * a];
      if (ea != 0.0) g.A[i][a] += r[e] * T(ea);
      for (int b = 0; b < n; ++b) {
        const double w = ea * e_mat[e + E * b] - (a == b ? ea : 0.0);
        if (w != 0.0) C[i][a][b] += r[e] * T(w);
      }
    }
    g.A[i][i] -= r[e];
  }
  for (int l = 0; l < n; ++l)
    for (int a = 0; a < n; ++a)
      for (int b = 0; b < n; ++b) {
        T h = C[l][a][b];
        if (a == b) h += g.A[l][a];
        if (b == l) h -= g.A[l][a];
        if (a == l) h -= g.A[l][b];
        g.H[l][a][b] = h;
      }
}

// cotangent of assemble: rbar_e += <Abar, dA/dr_e> + <Hbar, dH/dr_e>  (full-matrix inner products)
template <class T, int n>
inline void assemble_vjp(const double* e_mat, const int* p_vec, int E, const Gen<T, n>& gb, T* rbar) {
  T At[n][n];
  for (int l = 0; l < n; ++l) for (int a = 0; a < n; ++a)
    At[l][a] = gb.A[l][a] + gb.H[l][a][a] - gb.H[l][a][l] - gb.H[l][l][a];
  for (int e = 0; e < E; ++e) {
    const int i = p_vec[e] - 1;
    T acc = -At[i][i];
    for (int a = 0; a < n; ++a) {
      const double ea = e_mat[e + E * a];
      if (ea != 0.0) acc += At[i][a] * T(ea);
      for (int b = 0; b < n; ++b) {
        const double w = ea * e_mat[e + E * b] - (a == b ? ea : 0.0);
        if (w != 0.0) acc += gb.H[i][a][b] * T(w);
      }
    }
    rbar[e] += acc;
  }
}

// ---------------------------------------------------------------- moment state
template <class T, int n> struct State { T mu[n]; T S[n][n]; };

// L(w) = (A^T mu, A^T S + S A + sum_i mu_i H_i)
template <class T, int n>
inline void L_apply(const Gen<T, n>& g, const State<T, n>& w, State<T, n>& out) {
  for (int j = 0; j < n; ++j) { T s = T(0.0); for (int i = 0; i < n; ++i) s += g.A[i][j] * w.mu[i]; out.mu[j] = s; }
  T W[n][n];
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { T s = T(0.0); for (int k = 0; k < n; ++k) s += g.A[k][i] * w.S[k][j]; W[i][j] = s; }
  for (int i = 0; i < n; ++i) for (int j = i; j < n; ++j) {
    T s = W[i][j] + W[j][i];
    for (int l = 0; l < n; ++l) s += w.mu[l] * g.H[l][i][j];
    out.S[i][j] = s; out.S[j][i] = s;
  }
}
// L^T(wbar) = (A a + [<Sbar,H_i>]_i, A Sbar + Sbar A^T)
template <class T, int n>
inline void LT_apply(const Gen<T, n>& g, const State<T, n>& w, State<T, n>& out) {
  for (int i = 0; i < n; ++i) {
    T s = T(0.0);
    for (int j = 0; j < n; ++j) s += g.A[i][j] * w.mu[j];
    for (int a = 0; a < n; ++a) for (int b = 0; b < n; ++b) s += g.H[i][a][b] * w.S[a][b];
    out.mu[i] = s;
  }
  T W[n][n];
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { T s = T(0.0); for (int k = 0; k < n; ++k) s += g.A[i][k] * w.S[k][j]; W[i][j] = s; }
  for (int i = 0; i < n; ++i) for (int j = i; j < n; ++j) { T s = W[i][j] + W[j][i]; out.S[i][j] = s; out.S[j][i] = s; }
}

// y <- exp(t L) y ; when `store` is given every Taylor term of every step is kept for the reverse sweep
template <class T, int n>
inline void taylor_fwd(const Gen<T, n>& g, const Plan& p, State<T, n>& y, std::vector<State<T, n>>* store) {
  State<T, n> w, lw;
  for (int st = 0; st < p.s; ++st) {
    w = y;
    for (int k = 0; k < p.m; ++k) {
      if (store) store->push_back(w);
      L_apply(g, w, lw);
      const T al = T(p.h / (k + 1));
      for (int i = 0; i < n; ++i) { w.mu[i] = al * lw.mu[i]; y.mu[i] += w.mu[i]; }
      for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { w.S[i][j] = al * lw.S[i][j]; y.S[i][j] += w.S[i][j]; }
    }
  }
}

// reverse sweep: ybar is the cotangent of the output state; accumulates Abar, Hbar into gb.
template <class T, int n>
inline void taylor_bwd(const Gen<T, n>& g, const Plan& p, const std::vector<State<T, n>>& store,
                       State<T, n> ybar, Gen<T, n>& gb) {
  State<T, n> wb, lt;
  for (int st = p.s - 1; st >= 0; --st) {
    wb = ybar;
    for (int k = p.m - 1; k >= 0; --k) {
      const State<T, n>& w = store[(size_t)st * p.m + k];
      const T al = T(p.h / (k + 1));
      // Abar += al (mu a^T + 2 S Sbar) ; Hbar_l += al mu_l Sbar
      for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) {
        T s = w.mu[i] * wb.mu[j];
        T q = T(0.0);
        for (int k2 = 0; k2 < n; ++k2) q += w.S[i][k2] * wb.S[k2][j];
        gb.A[i][j] += al * (s + T(2.0) * q);
      }
      for (int l = 0; l < n; ++l) { const T c = al * w.mu[l]; for (int a = 0; a < n; ++a) for (int b = 0; b < n; ++b) gb.H[l][a][b] += c * wb.S[a][b]; }
      LT_apply(g, wb, lt);
      for (int i = 0; i < n; ++i) wb.mu[i] = ybar.mu[i] + al * lt.mu[i];
      for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) wb.S[i][j] = ybar.S[i][j] + al * lt.S[i][j];
    }
    ybar = wb;
  }
}

// ---------------------------------------------------------------- MVN score
// pop_vec[k] ~ multi_normal(mu, Sigma): -1/2 log det Sigma - 1/2 r^T Sigma^-1 r  (the -(n/2) log 2pi
// constant is dropped, as Stan's `~` does).  Returns false when Sigma is not positive definite /
// finite.  mub, Sb are d lp / d mu and d lp / d Sigma (full symmetric).
template <class T, int n>
inline bool mvn_score(const T* x, const State<T, n>& y, T& lp, State<T, n>& yb) {
  T Lc[n][n];
  for (int j = 0; j < n; ++j) {
    T d = y.S[j][j];
    for (int k = 0; k < j; ++k) d -= Lc[j][k] * Lc[j][k];
    if (!(val(d) > 0.0) || !std::isfinite(val(d))) return false;
    Lc[j][j] = sqrt(d);
    for (int i = j + 1; i < n; ++i) {
      T s = y.S[i][j];
      for (int k = 0; k < j; ++k) s -= Lc[i][k] * Lc[j][k];
      Lc[i][j] = s / Lc[j][j];
    }
  }
  // Linv (lower triangular inverse)
  T Li[n][n];
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) Li[i][j] = T(0.0);
  for (int j = 0; j < n; ++j) {
    Li[j][j] = T(1.0) / Lc[j][j];
    for (int i = j + 1; i < n; ++i) {
      T s = T(0.0);
      for (int k = j; k < i; ++k) s -= Lc[i][k] * Li[k][j];
      Li[i][j] = s / Lc[i][i];
    }
  }
  T r[n], a[n];
  for (int i = 0; i < n; ++i) { r[i] = x[i] - y.mu[i]; if (!std::isfinite(val(r[i]))) return false; }
  T q = T(0.0), ld = T(0.0);
  for (int i = 0; i < n; ++i) { T s = T(0.0); for (int k = 0; k <= i; ++k) s += Li[i][k] * r[k]; a[i] = s; q += s * s; ld += log(Lc[i][i]); }
  lp = -ld - T(0.5) * q;
  for (int i = 0; i < n; ++i) { T s = T(0.0); for (int k = i; k < n; ++k) s += Li[k][i] * a[k]; yb.mu[i] = s; }
  for (int i = 0; i < n; ++i) for (int j = i; j < n; ++j) {
    T si = T(0.0);
    for (int k = j; k < n; ++k) si += Li[k][i] * Li[k][j];
    T v = T(0.5) * (yb.mu[i] * yb.mu[j] - si);
    yb.S[i][j] = v; yb.S[j][i] = v;
  }
  return true;
}

// ---------------------------------------------------------------- multitype likelihood, one chain
// Data follow the Stan data block (multitype_template.stan:64-80): pop_vec / init_pop are N x n
// column-major, var_idx / times_idx 1-based, times[T].  r is [L][E] for this chain; rbar [L][E] is
// accumulated.  sigma_obs < 0 selects the noise-free template.  grouped != 0 computes single-ancestor
// moments once per (env level, duration) as the template does (lines 132-134); grouped == 0 runs the
// series per observation from (z0, 0).  Both give the same numbers; they differ in cost.
template <class T, int n>
inline int multitype_chain(int E, const double* e_mat, const int* p_vec, int N, const double* pop_vec,
                           const double* init_pop, const int* var_idx, const int* times_idx, int L, int Tn,
                           const double* times, const T* r, double sigma_obs, int grouped, double theta_step,
                           T& lp_out, T* rbar, T& sbar) {
  int status = ST_OK;
  const bool noise = sigma_obs >= 0.0;
  lp_out = T(0.0); sbar = T(0.0);
  std::vector<Gen<T, n>> gens(L), gbs(L);
  for (int v = 0; v < L; ++v) {
    assemble<T, n>(e_mat, p_vec, E, r + (size_t)v * E, gens[v]);
    for (int i = 0; i < n; ++i) for (int a = 0; a < n; ++a) { gbs[v].A[i][a] = T(0.0); for (int b = 0; b < n; ++b) gbs[v].H[i][a][b] = T(0.0); }
    for (int e = 0; e < E; ++e) if (!std::isfinite(val(r[(size_t)v * E + e]))) status |= ST_RATE_NONFINITE;
  }
  if (status) return status;
  std::vector<State<T, n>> store;
  auto score_one = [&](int k, State<T, n>& y, State<T, n>& yb) -> bool {
    T x[n];
    for (int i = 0; i < n; ++i) x[i] = T(pop_vec[k + (size_t)N * i]);
    if (noise) for (int i = 0; i < n; ++i) y.S[i][i] += T(sigma_obs) * y.mu[i];
    T lp;
    if (!mvn_score<T, n>(x, y, lp, yb)) return false;
    lp_out += lp;
    if (noise) for (int i = 0; i < n; ++i) { sbar += yb.S[i][i] * y.mu[i]; yb.mu[i] += T(sigma_obs) * yb.S[i][i]; }
    return true;
  };
  if (!grouped) {
    for (int k = 0; k < N; ++k) {
      const int v = var_idx[k] - 1; const double t = times[times_idx[k] - 1];
      Plan p = make_plan<T, n>(gens[v], t, theta_step);
      if (!p.ok) { status |= ST_RATE_NONFINITE; continue; }
      State<T, n> y, yb;
      for (int i = 0; i < n; ++i) { y.mu[i] = T(init_pop[k + (size_t)N * i]); for (int j = 0; j < n; ++j) y.S[i][j] = T(0.0); }
      store.clear();
      taylor_fwd<T, n>(gens[v], p, y, &store);
      if (!score_one(k, y, yb)) { status |= ST_NOT_PD; continue; }
      taylor_bwd<T, n>(gens[v], p, store, yb, gbs[v]);
    }
  } else {
    std::vector<State<T, n>> mom(n), momb(n);
    std::vector<std::vector<State<T, n>>> stores(n);
    for (int v = 0; v < L; ++v) for (int ti = 0; ti < Tn; ++ti) {
      bool any = false;
      for (int k = 0; k < N && !any; ++k) any = (var_idx[k] - 1 == v && times_idx[k] - 1 == ti);
      if (!any) continue;
      Plan p = make_plan<T, n>(gens[v], times[ti], theta_step);
      if (!p.ok) { status |= ST_RATE_NONFINITE; continue; }
      for (int l = 0; l < n; ++l) {
        for (int i = 0; i < n; ++i) { mom[l].mu[i] = T(i == l ? 1.0 : 0.0); momb[l].mu[i] = T(0.0); for (int j = 0; j < n; ++j) { mom[l].S[i][j] = T(0.0); momb[l].S[i][j] = T(0.0); } }
        stores[l].clear();
        taylor_fwd<T, n>(gens[v], p, mom[l], &stores[l]);
      }
      for (int k = 0; k < N; ++k) {
        if (var_idx[k] - 1 != v || times_idx[k] - 1 != ti) continue;
        State<T, n> y, yb;
        for (int i = 0; i < n; ++i) { y.mu[i] = T(0.0); for (int j = 0; j < n; ++j) y.S[i][j] = T(0.0); }
        for (int l = 0; l < n; ++l) {
          const T z = T(init_pop[k + (size_t)N * l]);
          for (int i = 0; i < n; ++i) { y.mu[i] += z * mom[l].mu[i]; for (int j = 0; j < n; ++j) y.S[i][j] += z * mom[l].S[i][j]; }
        }
        if (!score_one(k, y, yb)) { status |= ST_NOT_PD; continue; }
        for (int l = 0; l < n; ++l) {
          const T z = T(init_pop[k + (size_t)N * l]);
          for (int i = 0; i < n; ++i) { momb[l].mu[i] += z * yb.mu[i]; for (int j = 0; j < n; ++j) momb[l].S[i][j] += z * yb.S[i][j]; }
        }
      }
      for (int l = 0; l < n; ++l) taylor_bwd<T, n>(gens[v], p, stores[l], momb[l], gbs[v]);
    }
  }
  for (int v = 0; v < L; ++v) assemble_vjp<T, n>(e_mat, p_vec, E, gbs[v], rbar + (size_t)v * E);
  return status;
}

// ---------------------------------------------------------------- one-type closed form, one chain
// stanfiles/one_type_template.stan:37-46.  r is [L][2] = (birth, death) per env level.
template <class T>
inline int onetype_chain(int N, const double* pop_vec, const double* init_pop, const double* times,
                         const int* var_idx, int L, const T* r, T& lp_out, T* rbar) {
  int status = ST_OK;
  lp_out = T(0.0);
  for (int k = 0; k < N; ++k) {
    const int v = var_idx[k] - 1;
    const T b = r[2 * v], d = r[2 * v + 1], t = T(times[k]), z0 = T(init_pop[k]), x = T(pop_vec[k]);
    const T g = b - d;
    if (val(g) == 0.0 || !std::isfinite(val(g))) { status |= ST_ONETYPE_SINGULAR; continue; }
    const T e = exp(t * g);
    const T mu = e * z0;
    const T Nn = b * (T(2.0) * e - T(1.0)) - d;
    const T q = e * Nn / g - e * e;
    const T var = q * z0;
    if (!(val(var) > 0.0) || !std::isfinite(val(var))) { status |= ST_ONETYPE_SINGULAR; continue; }
    const T sig = sqrt(var);
    const T res = x - mu;
    lp_out += -log(sig) - T(0.5) * res * res / var;
    const T dmu = res / var;
    const T dvar = T(0.5) * (res * res / var - T(1.0)) / var;
    const T te = t * e;
    const T dq_db = (te * Nn + e * (T(2.0) * e - T(1.0) + T(2.0) * b * te)) / g - e * Nn / (g * g) - T(2.0) * te * e;
    const T dq_dd = (-te * Nn - e * (T(2.0) * b * te + T(1.0))) / g + e * Nn / (g * g) + T(2.0) * te * e;
    rbar[2 * v] += dmu * te * z0 + dvar * z0 * dq_db;
    rbar[2 * v + 1] += -dmu * te * z0 + dvar * z0 * dq_dd;
  }
  return status;
}

}  // namespace bpo
