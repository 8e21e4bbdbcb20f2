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

}  // extern "C"
