// hostsim.cpp — TEST INFRASTRUCTURE: compiles the per-thread core of bp_kernels.cuh (the BP_HD device
// functions, not the kernels) with g++ for one generated model, so that tests can (a) check the kernel
// arithmetic against the oracle without a GPU and (b) recount its operations with a counting scalar.
// Build: g++ -DBP_HOSTSIM [-DBP_COUNTED] -include <generated model source> hostsim.cpp   (tests/hostsim_util.py)
// The generated source must be passed with -DBP_MODEL_SRC="\"path\"".
#include <cmath>
#include <cstddef>
#include <cstdint>

#ifdef BP_COUNTED
struct real {
  double v;
  real() : v(0) {}
  real(double x) : v(x) {}
};
static uint64_t g_ops = 0;
inline real operator+(real a, real b) { ++g_ops; return real(a.v + b.v); }
inline real operator-(real a, real b) { ++g_ops; return real(a.v - b.v); }
inline real operator*(real a, real b) { ++g_ops; return real(a.v * b.v); }
inline real operator/(real a, real b) { ++g_ops; return real(a.v / b.v); }
inline real operator-(real a) { return real(-a.v); }
inline real& operator+=(real& a, real b) { a = a + b; return a; }
inline real& operator-=(real& a, real b) { a = a - b; return a; }
inline bool operator>(real a, real b) { return a.v > b.v; }
inline bool operator<(real a, real b) { return a.v < b.v; }
inline bool operator==(real a, real b) { return a.v == b.v; }
inline bool operator>=(real a, real b) { return a.v >= b.v; }
inline bool operator<=(real a, real b) { return a.v <= b.v; }
inline bool operator!=(real a, real b) { return a.v != b.v; }
inline real sqrt(real a) { ++g_ops; return real(std::sqrt(a.v)); }
inline real exp(real a) { ++g_ops; return real(std::exp(a.v)); }
inline real log(real a) { ++g_ops; return real(std::log(a.v)); }
inline real pow(real a, real b) { ++g_ops; return real(std::pow(a.v, b.v)); }
inline real fabs(real a) { return real(std::fabs(a.v)); }
inline real ceil(real a) { return real(std::ceil(a.v)); }
inline double bp_val(real a) { return a.v; }
#else
typedef double real;
using std::sqrt; using std::exp; using std::log; using std::pow; using std::fabs; using std::ceil;
inline double bp_val(double a) { return a; }
#endif

static real bp_host_store[4096];   // the term ring of the one simulated thread

#include BP_MODEL_SRC

extern "C" {

int hs_dims(int* out) {
  out[0] = BP_N; out[1] = BP_E; out[2] = BP_P; out[3] = BP_DIM; out[4] = BP_KDEP; out[5] = BP_KIND; out[6] = BP_MCAP; out[7] = BP_XDEP;
  return 0;
}

// one observation: theta [BP_DIM] (constrained, sigma_obs last), x [BP_KDEP]; returns lp and accumulates cb [BP_P], sb
int hs_obs(const double* theta, const double* x, const double* z0, const double* xo, double t, double* lp_out, double* cb_out,
           double* sb_out, int* apps_out, unsigned long long* ops_out) {
  real th[BP_DIM], xk[BP_KDEP], r[BP_E], cb[BP_P], rb[BP_E];
  for (int k = 0; k < BP_DIM; ++k) th[k] = real(theta[k]);
  for (int k = 0; k < BP_KDEP; ++k) xk[k] = real(x[k]);
  for (int k = 0; k < BP_P; ++k) cb[k] = real(0.0);
  int status = 0, apps = 0, dhint = 1;
  real lp = real(0.0), sb = real(0.0);
#ifdef BP_COUNTED
  g_ops = 0;
#endif
  bp_rates(th, xk, r);
#if BP_KIND != 2
  real A[BP_N * BP_N], H[BP_N * BP_S], gA[BP_N * BP_N], gH[BP_N * BP_S], z[BP_N], xx[BP_N];
  for (int i = 0; i < BP_N * BP_N; ++i) gA[i] = real(0.0);
  for (int i = 0; i < BP_N * BP_S; ++i) gH[i] = real(0.0);
  for (int i = 0; i < BP_N; ++i) { z[i] = real(z0[i]); xx[i] = real(xo[i]); }
  bp_assemble(r, A, H);
  const real a1 = bp_norm1(A);
#if BP_NOISE
  const real sigma = th[BP_P];
#else
  const real sigma = real(0.0);
#endif
#ifdef BP_COUNTED
  const uint64_t before = g_ops;
#endif
  bp_obs_lpgrad(A, H, a1, z, xx, real(t), sigma, 0, 1, lp, gA, gH, sb, status, apps, dhint);
#ifdef BP_COUNTED
  if (ops_out) ops_out[1] = g_ops - before;   // the per-observation core alone
#endif
  bp_assemble_vjp(gA, gH, rb);
#else
  real rb2[2] = {real(0.0), real(0.0)};
#ifdef BP_COUNTED
  const uint64_t before = g_ops;
#endif
  bp_onetype_obs(r[0], r[1], real(t), real(z0[0]), real(xo[0]), lp, rb2, status, false);
#ifdef BP_COUNTED
  if (ops_out) ops_out[1] = g_ops - before;
#endif
  rb[0] = rb2[0]; rb[1] = rb2[1];
#endif
  bp_rates_vjp(th, xk, rb, cb);
  *lp_out = bp_val(lp);
  for (int k = 0; k < BP_P; ++k) cb_out[k] = bp_val(cb[k]);
  *sb_out = bp_val(sb);
  *apps_out = apps;
#ifdef BP_COUNTED
  if (ops_out) ops_out[0] = g_ops;
#endif
  return status;
}

#if BP_KIND != 2 && BP_N <= 3
// The centred / octet formula for ONE (chain, observation) in scalar form, with exactly the arithmetic the octet path performs per
// observation (bp_fusedo: stage, forward product, score through the adjugate, packed cotangent, S' update), so that its operation
// count can be recounted with the counting scalar and its numbers compared with the direct series:
//   X[(j, l)] = c_j(h) z0_l;  y[d] = sum_(j, l) N_j[d][l] X[(j, l)];  (det, q, b) = score(y);  S'[(j, l)][d] += X[(j, l)] b[d]
// Nj: [J + 1][BP_D][BP_N].  out: y [BP_D], det, q, b [BP_D] (packed coordinates: off-diagonals doubled), S' [J + 1][BP_N][BP_D].
int hs_centred_obs(int J, const double* Nj, const double* z0, const double* xo, double h, double sigma, double* y_out, double* det_out,
                   double* q_out, double* b_out, double* sp_out, unsigned long long* ops_out) {
  real X[8][BP_N], y[BP_D], xx[BP_N], ymu[BP_N], yS[BP_S], bmu[BP_N], bS[BP_S], b[BP_D];
  if (J < 0 || J > 7) return -1;
#ifdef BP_COUNTED
  g_ops = 0;
#endif
  real w = real(1.0);
  const real hh = real(h);                                     // (t_k - tau_g: one subtraction, counted below)
  for (int d = 0; d < BP_D; ++d) y[d] = real(0.0);
  for (int j = 0; j <= J; ++j) {
    if (j > 0) w = w * (hh * real(bp_rk[j - 1]));
    for (int l = 0; l < BP_N; ++l) X[j][l] = w * real(z0[l]);
    for (int d = 0; d < BP_D; ++d)
      for (int l = 0; l < BP_N; ++l) y[d] += real(Nj[(j * BP_D + d) * BP_N + l]) * X[j][l];
  }
  for (int i = 0; i < BP_N; ++i) { ymu[i] = y[i]; xx[i] = real(xo[i]); }
  for (int q = 0; q < BP_S; ++q) yS[q] = y[BP_N + q];
#if BP_NOISE
  for (int i = 0; i < BP_N; ++i) yS[bp_sidx(i, i)] += real(sigma) * ymu[i];
#endif
  real det, q;
  const bool ok = bp_mvn_score_small<true>(xx, ymu, yS, det, q, bmu, bS);   // bS in packed coordinates
  real qsum = real(0.0), dprod = real(1.0);
  if (ok) { qsum += q; dprod = dprod * det; }                 // the running sums of bp_fusedo (one logarithm per thread at the very end)
#if BP_NOISE
  real sb = real(0.0);
  for (int i = 0; i < BP_N; ++i) { sb += bS[bp_sidx(i, i)] * ymu[i]; bmu[i] += real(sigma) * bS[bp_sidx(i, i)]; }
#endif
  for (int i = 0; i < BP_N; ++i) b[i] = bmu[i];
  for (int qq = 0; qq < BP_S; ++qq) b[BP_N + qq] = bS[qq];
  for (int j = 0; j <= J; ++j)
    for (int l = 0; l < BP_N; ++l)
      for (int d = 0; d < BP_D; ++d) {
        real acc = real(sp_out[(j * BP_N + l) * BP_D + d]);
        acc += X[j][l] * b[d];
        sp_out[(j * BP_N + l) * BP_D + d] = bp_val(acc);
      }
#ifdef BP_COUNTED
  if (ops_out) ops_out[0] = g_ops + 1;                        // + the subtraction h = t - tau
#endif
  for (int d = 0; d < BP_D; ++d) { y_out[d] = bp_val(y[d]); b_out[d] = bp_val(b[d]); }
  *det_out = bp_val(det);
  *q_out = bp_val(q);
  return ok ? 0 : 1;
}

// both scores on the same input: the LDL^T one (lp, bmu, bS) and the adjugate one (det, q, bmu, bS)
int hs_score_both(const double* x, const double* mu, const double* S, double* lp_out, double* b1, double* det_out, double* q_out, double* b2) {
  real xx[BP_N], m[BP_N], s[BP_S], bmu[BP_N], bS[BP_S], lp, det, q;
  for (int i = 0; i < BP_N; ++i) { xx[i] = real(x[i]); m[i] = real(mu[i]); }
  for (int k = 0; k < BP_S; ++k) s[k] = real(S[k]);
  const bool ok1 = bp_mvn_score(xx, m, s, lp, bmu, bS);
  *lp_out = bp_val(lp);
  for (int i = 0; i < BP_N; ++i) b1[i] = bp_val(bmu[i]);
  for (int k = 0; k < BP_S; ++k) b1[BP_N + k] = bp_val(bS[k]);
  const bool ok2 = bp_mvn_score_small<false>(xx, m, s, det, q, bmu, bS);
  *det_out = bp_val(det); *q_out = bp_val(q);
  for (int i = 0; i < BP_N; ++i) b2[i] = bp_val(bmu[i]);
  for (int k = 0; k < BP_S; ++k) b2[BP_N + k] = bp_val(bS[k]);
  return (ok1 ? 0 : 1) | (ok2 ? 0 : 2);
}
#endif

}  // extern "C"
