// sampler.cu — device-resident many-chain NUTS driver of libbpcuda (sm_100a).
//
// Replaces what `rstan::sampling` does around the log-density path (sampler settings of
// the reference's examples/*.R): multinomial NUTS with a diagonal metric, dual-averaging step-size
// adaptation to adapt_delta, Stan's windowed metric adaptation (init buffer / doubling windows / term
// buffer) and the step-size search that follows every metric update.  The reference runs one chain per R
// process; here every chain is a small state machine living in HBM, one thread per chain, advanced by ONE
// leapfrog step per "pass":
//     pass = bp_prologue + bp_fused/bp_onetype + bp_epilogue   (log density + gradient of all chains, NVRTC module)
//          + bp_nuts_advance                                   (this file: finish the leapfrog, tree bookkeeping,
//                                                               U-turn checks, adaptation, start the next leapfrog)
// Chains do not wait for each other: a chain that ends its trajectory starts the next transition in the same
// pass.  The tree is built iteratively (leaf by leaf with checkpointed partial momentum sums for the sub-tree
// U-turn checks), which is the same tree and the same multinomial sampling as the recursive formulation.
// Random numbers: Philox4x32-10, key = seed, counter = (global chain id, draw index) — reproducible and
// independent of how chains are sharded over GPUs.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

#include "bpcuda.h"
#include "bpcuda_internal.hpp"

using namespace bpc;

namespace {

constexpr int MAXD = 12;        // max tree depth supported
constexpr int MAXDIM = 32;      // max unconstrained dimension

enum Phase { PH_INIT = 0, PH_FIND_EPS_FIRST = 1, PH_FIND_EPS = 2, PH_NUTS = 3, PH_WAIT = 4, PH_DONE = 5, PH_FAILED = 6, PH_OPT = 7, PH_HESS = 8 };

// scalar double fields
enum { S_EPS, S_LP_CUR, S_H0, S_LSW, S_SUB_LSW, S_SUM_ACC, S_DA_MU, S_DA_SBAR, S_DA_XBAR, S_DA_CNT, S_LP_PROP, S_LP_SP, S_W_N, S_NLEAP_TOT, S_OPT_LR, S_OPT_GD, S_OPT_FLAT, NSCAL };
// vector double fields (dim each)
enum { V_Q_CUR, V_G_CUR, V_QL, V_PL, V_GL, V_QR, V_PR, V_GR, V_Q_PROP, V_G_PROP, V_Q_SP, V_G_SP, V_RHO, V_RHO_SUB, V_MINV, V_W_MEAN, V_W_M2, V_W_EVAL, NVEC };
// int fields
enum { I_PHASE, I_ITER, I_DEPTH, I_NLEAP, I_DIR, I_LEAF, I_DIVERGENT, I_FE_DIR, I_WIN_NEXT, I_WIN_SIZE, I_WIN_COUNTER, I_TRIES, I_RNG, I_PSEUDO, I_STAGE, NINT };

struct Params {
  int C, dim, P, noise, max_depth, iter, warmup, chain_offset, pooled, have_inits, opt_steps, dense, nsum, debug, hess, two_stage;
  int init_buffer, term_buffer, base_window;
  double delta, eps0;
  unsigned long long seed;
  double* sd;          // [(NSCAL + NVEC*dim + 2*MAXD*dim + nsum)][C]
  double* metric;      // dense metric shared by the chains of this device: mu[dim], L[dim][dim] (Sigma = L L^T, lower); then the previous (mu, L)
  int* si;             // [NINT][C]
  double* q_eval;      // [C][dim]  position at which the log density is evaluated in the next pass
  double* lp_eval;     // [C]
  double* g_eval;      // [C][dim]
  int* st_eval;        // [C]
  double* draws;       // [C][nsave][dim]
  double* d_lp; double* d_acc; double* d_eps; int* d_depth; int* d_nleap; int* d_div;
  int* counters;       // [0] finished chains, [1] waiting chains, [2] failed chains
  int has_bounds[MAXDIM];
  double lo[MAXDIM], hi[MAXDIM];
};

// ---- Philox4x32-10 ---------------------------------------------------------------------------------
__device__ __forceinline__ void philox_round(unsigned& c0, unsigned& c1, unsigned& c2, unsigned& c3, unsigned k0, unsigned k1) {
  const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
  const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
  const unsigned n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
  c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}
__device__ void philox(unsigned long long seed, unsigned chain, unsigned ctr, unsigned (&out)[4]) {
  unsigned c0 = ctr, c1 = chain, c2 = 0x2545F491u, c3 = 0u;
  unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) { philox_round(c0, c1, c2, c3, k0, k1); k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct Chain {
  const Params& P;
  int c;
  __device__ Chain(const Params& p, int chain) : P(p), c(chain) {}
  __device__ double& s(int f) const { return P.sd[(size_t)f * P.C + c]; }
  __device__ double& v(int f, int k) const { return P.sd[(size_t)(NSCAL + f * P.dim + k) * P.C + c]; }
  __device__ double& rck(int lvl, int k) const { return P.sd[(size_t)(NSCAL + NVEC * P.dim + lvl * P.dim + k) * P.C + c]; }
  __device__ double& rsck(int lvl, int k) const { return P.sd[(size_t)(NSCAL + NVEC * P.dim + (MAXD + lvl) * P.dim + k) * P.C + c]; }
  __device__ double& cross(int idx) const { return P.sd[(size_t)(NSCAL + NVEC * P.dim + 2 * MAXD * P.dim + idx) * P.C + c]; }
  __device__ int& i(int f) const { return P.si[(size_t)f * P.C + c]; }
  // two uniforms in (0,1) per call
  __device__ void uniforms(double& u0, double& u1) const {
    unsigned r[4];
    philox(P.seed, (unsigned)(P.chain_offset + c), (unsigned)i(I_RNG), r);
    i(I_RNG) += 1;
    u0 = (((unsigned long long)r[0] << 21) ^ (r[1] >> 11)) * (1.0 / 9007199254740992.0) + (0.5 / 9007199254740992.0);
    u1 = (((unsigned long long)r[2] << 21) ^ (r[3] >> 11)) * (1.0 / 9007199254740992.0) + (0.5 / 9007199254740992.0);
  }
  __device__ double uniform() const { double a, b; uniforms(a, b); return a; }
  __device__ void normals(double& z0, double& z1) const {
    double a, b;
    uniforms(a, b);
    const double r = sqrt(-2.0 * log(a));
    double sn, cs;
    sincospi(2.0 * b, &sn, &cs);
    z0 = r * cs; z1 = r * sn;
  }
};

__device__ __forceinline__ double logaddexp(double a, double b) {
  if (a == -INFINITY) return b;
  if (b == -INFINITY) return a;
  const double m = fmax(a, b);
  return m + log1p(exp(-fabs(a - b)));
}

// ---- the metric as a change of coordinates --------------------------------------------------------------------------
// The sampler works on w with q = mu + L w.  Diagonal metric (rstan's default diag_e): L = I, mu = 0 and the per-chain inverse
// mass V_MINV, exactly as before.  Dense metric (rstan's dense_e; here with the covariance pooled over all chains): Sigma = L L^T is
// shared by the chains of a device, the NUTS arithmetic runs in the whitened coordinates with a unit metric -- which IS HMC with
// M^-1 = Sigma -- and only the two hand-offs with the log-density kernels change: q_eval = mu + L w_eval, grad_w = L^T grad_q.
__device__ __forceinline__ void metric_to_q(const Params& P, const double* w, double* q) {
  const int dim = P.dim;
  if (!P.dense) { for (int k = 0; k < dim; ++k) q[k] = w[k]; return; }
  const double* mu = P.metric;
  const double* L = P.metric + dim;
  for (int i = 0; i < dim; ++i) {
    double a = mu[i];
    for (int j = 0; j <= i; ++j) a += L[i * dim + j] * w[j];
    q[i] = a;
  }
}

// the point w_eval of this chain -> q_eval, where the next pass evaluates the log density
__device__ void publish_eval(const Chain& ch) {
  const int dim = ch.P.dim;
  double w[MAXDIM], q[MAXDIM];
  for (int k = 0; k < dim; ++k) w[k] = ch.v(V_W_EVAL, k);
  metric_to_q(ch.P, w, q);
  for (int k = 0; k < dim; ++k) ch.P.q_eval[(size_t)ch.c * dim + k] = q[k];
}

// gradient of the last pass in the sampler's coordinates: L^T grad_q
__device__ __forceinline__ void load_grad(const Chain& ch, double* gw) {
  const Params& P = ch.P;
  const int dim = P.dim;
  double g[MAXDIM];
  for (int k = 0; k < dim; ++k) g[k] = P.g_eval[(size_t)ch.c * dim + k];
  if (!P.dense) { for (int k = 0; k < dim; ++k) gw[k] = g[k]; return; }
  const double* L = P.metric + dim;
  for (int j = 0; j < dim; ++j) {
    double a = 0.0;
    for (int i = j; i < dim; ++i) a += L[i * dim + j] * g[i];
    gw[j] = a;
  }
}

// p ~ N(0, M), M = diag(1/minv); writes into vector field f
__device__ void sample_momentum(const Chain& ch, int f) {
  const int dim = ch.P.dim;
  for (int k = 0; k < dim; k += 2) {
    double z0, z1;
    ch.normals(z0, z1);
    ch.v(f, k) = z0 / sqrt(ch.v(V_MINV, k));
    if (k + 1 < dim) ch.v(f, k + 1) = z1 / sqrt(ch.v(V_MINV, k + 1));
  }
}

__device__ double kinetic(const Chain& ch, int f) {
  double ke = 0.0;
  for (int k = 0; k < ch.P.dim; ++k) ke += ch.v(V_MINV, k) * ch.v(f, k) * ch.v(f, k);
  return 0.5 * ke;
}

// half kick + drift from the edge (fq, fp, fg) in direction dir; leaves the half-kicked momentum in fp and writes q_eval
__device__ void start_leapfrog(const Chain& ch, int fq, int fp, int fg, int dir) {
  const double e = ch.s(S_EPS) * dir;
  for (int k = 0; k < ch.P.dim; ++k) {
    const double ph = ch.v(fp, k) + 0.5 * e * ch.v(fg, k);
    ch.v(fp, k) = ph;
    ch.v(V_W_EVAL, k) = ch.v(fq, k) + e * ch.v(V_MINV, k) * ph;
  }
  publish_eval(ch);
}

// second half kick with the freshly evaluated gradient; the edge becomes the new leaf.  Returns its Hamiltonian.
__device__ double finish_leapfrog(const Chain& ch, int fq, int fp, int fg, int dir, double& lp_new) {
  const double e = ch.s(S_EPS) * dir;
  const int dim = ch.P.dim;
  lp_new = ch.P.st_eval[ch.c] ? -INFINITY : ch.P.lp_eval[ch.c];
  double gw[MAXDIM];
  load_grad(ch, gw);
  for (int k = 0; k < dim; ++k) {
    const double g = gw[k];
    ch.v(fq, k) = ch.v(V_W_EVAL, k);
    ch.v(fg, k) = g;
    ch.v(fp, k) += 0.5 * e * g;
  }
  double H = -lp_new + kinetic(ch, fp);
  if (isnan(H)) H = INFINITY;
  return H;
}

__device__ void begin_subtree(const Chain& ch) {
  ch.i(I_DIR) = ch.uniform() < 0.5 ? -1 : 1;
  ch.i(I_LEAF) = 0;
  ch.s(S_SUB_LSW) = -INFINITY;
  for (int k = 0; k < ch.P.dim; ++k) ch.v(V_RHO_SUB, k) = 0.0;
  const int dir = ch.i(I_DIR);
  start_leapfrog(ch, dir > 0 ? V_QR : V_QL, dir > 0 ? V_PR : V_PL, dir > 0 ? V_GR : V_GL, dir);
}

__device__ void begin_transition(const Chain& ch) {
  const int dim = ch.P.dim;
  sample_momentum(ch, V_PL);
  for (int k = 0; k < dim; ++k) {
    const double q = ch.v(V_Q_CUR, k), g = ch.v(V_G_CUR, k), p = ch.v(V_PL, k);
    ch.v(V_QL, k) = q; ch.v(V_QR, k) = q; ch.v(V_GL, k) = g; ch.v(V_GR, k) = g; ch.v(V_PR, k) = p;
    ch.v(V_Q_PROP, k) = q; ch.v(V_G_PROP, k) = g; ch.v(V_RHO, k) = p;
  }
  ch.s(S_LP_PROP) = ch.s(S_LP_CUR);
  ch.s(S_H0) = -ch.s(S_LP_CUR) + kinetic(ch, V_PL);
  ch.s(S_LSW) = 0.0;
  ch.s(S_SUM_ACC) = 0.0;
  ch.i(I_DEPTH) = 0;
  ch.i(I_NLEAP) = 0;
  ch.i(I_DIVERGENT) = 0;
  ch.i(I_PHASE) = PH_NUTS;
  begin_subtree(ch);
}

// Stan's init_stepsize: one trial leapfrog from the current point with a fresh momentum
__device__ void begin_eps_trial(const Chain& ch, int phase) {
  const int dim = ch.P.dim;
  sample_momentum(ch, V_PR);
  for (int k = 0; k < dim; ++k) { ch.v(V_QR, k) = ch.v(V_Q_CUR, k); ch.v(V_GR, k) = ch.v(V_G_CUR, k); }
  ch.s(S_H0) = -ch.s(S_LP_CUR) + kinetic(ch, V_PR);
  ch.i(I_PHASE) = phase;
  start_leapfrog(ch, V_QR, V_PR, V_GR, 1);
}

__device__ void restart_dual_averaging(const Chain& ch) {
  ch.s(S_DA_MU) = log(10.0 * ch.s(S_EPS));
  ch.s(S_DA_SBAR) = 0.0;
  ch.s(S_DA_XBAR) = 0.0;
  ch.s(S_DA_CNT) = 0.0;
}

__device__ bool uturn_ok(const Chain& ch, int fpa, int fpb, int frho) {   // true = keep going
  double a = 0.0, b = 0.0;
  for (int k = 0; k < ch.P.dim; ++k) {
    const double r = ch.v(frho, k), mi = ch.v(V_MINV, k);
    a += mi * ch.v(fpa, k) * r;
    b += mi * ch.v(fpb, k) * r;
  }
  return a > 0.0 && b > 0.0;
}

// Stan's windowed_adaptation::compute_next_window
__device__ void compute_next_window(const Chain& ch) {
  const int end = ch.P.warmup - ch.P.term_buffer - 1;
  if (ch.i(I_WIN_NEXT) == end) return;
  ch.i(I_WIN_SIZE) *= 2;
  ch.i(I_WIN_NEXT) = ch.i(I_WIN_COUNTER) + ch.i(I_WIN_SIZE);
  if (ch.i(I_WIN_NEXT) == end) return;
  const int boundary = ch.i(I_WIN_NEXT) + 2 * ch.i(I_WIN_SIZE);
  if (boundary >= ch.P.warmup - ch.P.term_buffer) ch.i(I_WIN_NEXT) = end;
}

__device__ void record_draw(const Chain& ch, int it, double accept) {
  const Params& P = ch.P;
  const int sidx = it - P.warmup;
  if (sidx < 0) return;
  const int nsave = P.iter - P.warmup;
  const size_t row = (size_t)ch.c * nsave + sidx;
  double wv[MAXDIM], qv[MAXDIM];
  for (int k = 0; k < P.dim; ++k) wv[k] = ch.v(V_Q_CUR, k);
  metric_to_q(P, wv, qv);
  for (int k = 0; k < P.dim; ++k) {
    const double u = qv[k];
    double th = u;
    if (k < P.P) {
      if (P.has_bounds[k]) {
        const double sg = (u >= 0.0) ? 1.0 / (1.0 + exp(-u)) : exp(u) / (1.0 + exp(u));
        th = P.lo[k] + (P.hi[k] - P.lo[k]) * sg;
      }
    } else {
      th = exp(u);   // sigma_obs
    }
    P.draws[row * P.dim + k] = th;
  }
  P.d_lp[row] = ch.s(S_LP_CUR);
  P.d_acc[row] = accept;
  P.d_eps[row] = ch.s(S_EPS);
  P.d_depth[row] = ch.i(I_DEPTH);
  P.d_nleap[row] = ch.i(I_NLEAP);
  P.d_div[row] = ch.i(I_DIVERGENT);
}

// end of a transition: move to the proposal, record, adapt, and start whatever comes next
__device__ void end_transition(const Chain& ch) {
  const Params& P = ch.P;
  const int dim = P.dim;
  for (int k = 0; k < dim; ++k) { ch.v(V_Q_CUR, k) = ch.v(V_Q_PROP, k); ch.v(V_G_CUR, k) = ch.v(V_G_PROP, k); }
  ch.s(S_LP_CUR) = ch.s(S_LP_PROP);
  const int nl = max(1, ch.i(I_NLEAP));
  const double accept = ch.s(S_SUM_ACC) / nl;
  ch.s(S_NLEAP_TOT) += ch.i(I_NLEAP);
  const int it = ch.i(I_ITER);
  record_draw(ch, it, accept);
  ch.i(I_ITER) = it + 1;
  bool metric_updated = false, wait = false, eps_wait = false;
  if (it < P.warmup) {
    // dual averaging (Stan's stepsize_adaptation::learn_stepsize: gamma 0.05, t0 10, kappa 0.75)
    const double cnt = ch.s(S_DA_CNT) + 1.0;
    ch.s(S_DA_CNT) = cnt;
    const double stat = fmin(1.0, accept);
    const double eta = 1.0 / (cnt + 10.0);
    const double sbar = (1.0 - eta) * ch.s(S_DA_SBAR) + eta * (P.delta - stat);
    ch.s(S_DA_SBAR) = sbar;
    const double x = ch.s(S_DA_MU) - sbar * sqrt(cnt) / 0.05;
    const double xeta = pow(cnt, -0.75);
    ch.s(S_DA_XBAR) = (1.0 - xeta) * ch.s(S_DA_XBAR) + xeta * x;
    ch.s(S_EPS) = exp(x);
    // windowed metric adaptation (Stan's var_adaptation::learn_variance)
    const int wc = ch.i(I_WIN_COUNTER);
    const bool in_window = wc >= P.init_buffer && wc < P.warmup - P.term_buffer && wc != P.warmup;
    if (in_window && P.dense) {
      // dense metric: raw sums of dq = q - mu_cur = L w and dq dq^T (the centre is shared by all chains, so the sums of all chains add)
      ch.s(S_W_N) += 1.0;
      double dq[MAXDIM];
      const double* Lm = P.metric + dim;
      for (int i = 0; i < dim; ++i) {
        double a = 0.0;
        for (int j = 0; j <= i; ++j) a += Lm[i * dim + j] * ch.v(V_Q_CUR, j);
        dq[i] = a;
        ch.v(V_W_MEAN, i) += a;
      }
      int idx = 0;
      for (int i = 0; i < dim; ++i)
        for (int j = 0; j <= i; ++j, ++idx) ch.cross(idx) += dq[i] * dq[j];
    } else if (in_window) {
      const double n = ch.s(S_W_N) + 1.0;
      ch.s(S_W_N) = n;
      for (int k = 0; k < dim; ++k) {
        const double q = ch.v(V_Q_CUR, k);
        const double d = q - ch.v(V_W_MEAN, k);
        const double mean = ch.v(V_W_MEAN, k) + d / n;
        ch.v(V_W_MEAN, k) = mean;
        ch.v(V_W_M2, k) += d * (q - mean);
      }
    }
    if (wc == ch.i(I_WIN_NEXT) && wc != P.warmup && in_window) {
      if (P.pooled) {
        wait = true;   // the pooled metric is set by bp_pool_apply once every chain of every rank is here
      } else {
        compute_next_window(ch);
        const double n = ch.s(S_W_N);
        for (int k = 0; k < dim; ++k) {
          const double var = n > 1.0 ? ch.v(V_W_M2, k) / (n - 1.0) : 1.0;
          ch.v(V_MINV, k) = (n / (n + 5.0)) * var + 1e-3 * (5.0 / (n + 5.0));
          ch.v(V_W_MEAN, k) = 0.0;
          ch.v(V_W_M2, k) = 0.0;
        }
        ch.s(S_W_N) = 0.0;
        metric_updated = true;
      }
    }
    ch.i(I_WIN_COUNTER) = wc + 1;
    if (it + 1 == P.warmup) {
      ch.s(S_EPS) = exp(ch.s(S_DA_XBAR));   // complete_adaptation
      // pooled adaptation: one more exchange -- every chain samples with the geometric mean of the adapted step sizes.  With
      // one metric for all chains there is one right step size; a chain whose own dual averaging ended a factor of two low would
      // build trees one level deeper for the whole sampling phase and hold up the end of the run.
      if (P.pooled) eps_wait = true;
    }
  }
  if (ch.i(I_ITER) >= P.iter) {
    ch.i(I_PHASE) = PH_DONE;
    atomicAdd(P.counters + 0, 1);
    return;
  }
  if (wait || eps_wait) {
    if (eps_wait) ch.i(I_PSEUDO) = 2;
    ch.i(I_PHASE) = PH_WAIT;
    atomicAdd(P.counters + 1, 1);
    return;
  }
  if (metric_updated) {
    ch.i(I_FE_DIR) = 0;
    begin_eps_trial(ch, PH_FIND_EPS_FIRST);
    return;
  }
  begin_transition(ch);
}

// ---- optional climb from the initial point (bpc_sample_args.init_opt_steps) ---------------------------------------------
// rstan starts NUTS at the initial values.  With 10^5 observations the posterior is three or four orders of magnitude narrower
// than the range initial values are drawn from (uniform_initialize(0, 1), R/stan_utils.R:74-87, or rstan's uniform(-2, 2)); far
// out there the generator norm is large, every evaluation needs sub-steps (the slow path), the first NUTS iterations are long
// trees of tiny steps, and chains that start on the wrong end of the posterior's ridge (only birth - death is sharply identified)
// crawl along it for the whole warm-up.  So, when asked to, every chain first maximises the unconstrained log density with at
// most init_opt_steps evaluations of L-BFGS (6 curvature pairs, backtracking Armijo line search), each evaluation being the same
// fused log-density + gradient pass NUTS uses -- what rstan's `optimizing` would give as initial values.
// State: x = V_Q_CUR, ascent gradient = V_G_CUR, f = S_LP_CUR; direction V_PL, step S_OPT_LR, g.d S_OPT_GD; pairs s_i / y_i in the
// (still unused) U-turn checkpoint rows rck / rsck; I_LEAF evaluations, I_NLEAP stored pairs, I_DIR newest pair, I_DEPTH failed trials.
constexpr int LBFGS_M = 10;

__device__ void opt_direction(const Chain& ch) {
  const int dim = ch.P.dim;
  const int np = ch.i(I_NLEAP), head = ch.i(I_DIR);
  double q[MAXDIM], al[LBFGS_M], rho[LBFGS_M];
  for (int k = 0; k < dim; ++k) q[k] = ch.v(V_G_CUR, k);
  double gamma = 0.0;
  for (int t = 0; t < np; ++t) {          // newest -> oldest
    const int i = (head - t + LBFGS_M) % LBFGS_M;
    double sy = 0.0, sq = 0.0, yy = 0.0;
    for (int k = 0; k < dim; ++k) { const double sv = ch.rck(i, k), yv = ch.rsck(i, k); sy += sv * yv; sq += sv * q[k]; yy += yv * yv; }
    rho[t] = 1.0 / sy;
    al[t] = rho[t] * sq;
    for (int k = 0; k < dim; ++k) q[k] -= al[t] * ch.rsck(i, k);
    if (t == 0) gamma = sy / yy;
  }
  double gmax = 0.0;
  for (int k = 0; k < dim; ++k) gmax = fmax(gmax, fabs(ch.v(V_G_CUR, k)));
  if (np == 0) gamma = gmax > 0.0 ? 0.25 / gmax : 1.0;    // first step: at most 0.25 in any coordinate
  for (int k = 0; k < dim; ++k) q[k] *= gamma;
  for (int t = np - 1; t >= 0; --t) {     // oldest -> newest
    const int i = (head - t + LBFGS_M) % LBFGS_M;
    double yr = 0.0;
    for (int k = 0; k < dim; ++k) yr += ch.rsck(i, k) * q[k];
    const double b = rho[t] * yr;
    for (int k = 0; k < dim; ++k) q[k] += ch.rck(i, k) * (al[t] - b);
  }
  double gd = 0.0, dmax = 0.0;
  for (int k = 0; k < dim; ++k) { gd += ch.v(V_G_CUR, k) * q[k]; dmax = fmax(dmax, fabs(q[k])); }
  if (!(gd > 0.0) || !isfinite(gd)) {     // not an ascent direction: forget the pairs, steepest ascent
    ch.i(I_NLEAP) = 0;
    const double g0 = gmax > 0.0 ? 0.25 / gmax : 1.0;
    gd = 0.0; dmax = 0.0;
    for (int k = 0; k < dim; ++k) { q[k] = g0 * ch.v(V_G_CUR, k); gd += ch.v(V_G_CUR, k) * q[k]; dmax = fmax(dmax, fabs(q[k])); }
  }
  for (int k = 0; k < dim; ++k) ch.v(V_PL, k) = q[k];
  ch.s(S_OPT_GD) = gd;
  ch.s(S_OPT_LR) = dmax > 1.0 ? 1.0 / dmax : 1.0;          // never more than 1 in any coordinate per step
  ch.i(I_DEPTH) = 0;
}

__device__ void opt_propose(const Chain& ch) {
  const double a = ch.s(S_OPT_LR);
  for (int k = 0; k < ch.P.dim; ++k) ch.v(V_W_EVAL, k) = ch.v(V_Q_CUR, k) + a * ch.v(V_PL, k);
  publish_eval(ch);
}

__device__ void begin_opt(const Chain& ch) {
  ch.i(I_LEAF) = 0;
  ch.i(I_NLEAP) = 0;
  ch.i(I_DIR) = 0;
  ch.s(S_OPT_FLAT) = 0.0;
  ch.i(I_PHASE) = PH_OPT;
  opt_direction(ch);
  opt_propose(ch);
}

// ---- the metric of the first warm-up iterations from the curvature at the optimum (pooled adaptation + init_opt_steps) ----------
// NUTS with a unit metric on a posterior of scale 1e-3 spends its first window (the most expensive one: the deepest trees) learning
// what the optimiser already knows.  After the ascent every chain takes 2 dim more evaluations -- central differences of the gradient
// along the coordinate axes -- for the Hessian at its optimum, inverts it (Laplace approximation Sigma_c = (-H)^-1), and hands
// (x_c, Sigma_c) to the pooled-window exchange as if it were a window of HESS_N draws with that mean and covariance: the chains wait
// (PH_WAIT, I_PSEUDO), the statistics are added over all chains of all ranks like any other window, and the first metric (diagonal or
// dense) is the average Laplace covariance.  A chain whose Hessian is not negative definite contributes nothing.  The window
// schedule does not advance; the real windows refine the metric afterwards.  g(x +- delta e_k) are kept in rck / rsck.
constexpr double HESS_N = 4.0;
__device__ __forceinline__ double hess_delta(double x) { return 2e-6 * (1.0 + fabs(x)); }

__device__ void hess_propose(const Chain& ch) {
  const int idx = ch.i(I_LEAF), k = idx >> 1;
  for (int j = 0; j < ch.P.dim; ++j) ch.v(V_W_EVAL, j) = ch.v(V_Q_CUR, j);
  ch.v(V_W_EVAL, k) += ((idx & 1) ? -1.0 : 1.0) * hess_delta(ch.v(V_Q_CUR, k));
  publish_eval(ch);
}

__device__ void begin_hess(const Chain& ch) {
  ch.i(I_LEAF) = 0;
  ch.i(I_DEPTH) = 0;     // 1: an evaluation failed
  ch.i(I_PHASE) = PH_HESS;
  hess_propose(ch);
}

__device__ void hess_step(const Chain& ch) {
  const Params& P = ch.P;
  const int dim = P.dim, c = ch.c;
  const int idx = ch.i(I_LEAF), k = idx >> 1;
  double gw[MAXDIM];
  load_grad(ch, gw);
  bool fin = P.st_eval[c] == 0;
  for (int j = 0; j < dim; ++j) fin = fin && isfinite(gw[j]);
  if (!fin) ch.i(I_DEPTH) = 1;
  for (int j = 0; j < dim; ++j) { if (idx & 1) ch.rsck(k, j) = gw[j]; else ch.rck(k, j) = gw[j]; }
  ch.i(I_LEAF) = idx + 1;
  if (idx + 1 < 2 * dim) { hess_propose(ch); return; }
  // A = -(H + H^T) / 2,  H[j][k] = (g_j(x + d e_k) - g_j(x - d e_k)) / (2 d)
  double A[MAXD * MAXD], R[MAXD * MAXD], Sg[MAXD * MAXD];
  bool ok = ch.i(I_DEPTH) == 0;
  for (int a = 0; a < dim; ++a)
    for (int b = 0; b <= a; ++b) {
      const double hab = (ch.rck(b, a) - ch.rsck(b, a)) / (2.0 * hess_delta(ch.v(V_Q_CUR, b)));
      const double hba = (ch.rck(a, b) - ch.rsck(a, b)) / (2.0 * hess_delta(ch.v(V_Q_CUR, a)));
      A[a * dim + b] = A[b * dim + a] = -0.5 * (hab + hba);
    }
  for (int i = 0; i < dim && ok; ++i)
    for (int j = 0; j <= i; ++j) {
      double a = A[i * dim + j];
      for (int q = 0; q < j; ++q) a -= R[i * dim + q] * R[j * dim + q];
      if (i == j) { if (!(a > 0.0) || !isfinite(a)) { ok = false; break; } R[i * dim + i] = sqrt(a); }
      else R[i * dim + j] = a / R[j * dim + j];
    }
  if (ok) {   // Sigma = A^-1, column by column: R y = e_col, R^T x = y
    for (int col = 0; col < dim; ++col) {
      double y[MAXD];
      for (int i = 0; i < dim; ++i) { double a = (i == col) ? 1.0 : 0.0; for (int q = 0; q < i; ++q) a -= R[i * dim + q] * y[q]; y[i] = a / R[i * dim + i]; }
      for (int i = dim - 1; i >= 0; --i) { double a = y[i]; for (int q = i + 1; q < dim; ++q) a -= R[q * dim + i] * Sg[q * dim + col]; Sg[i * dim + col] = a / R[i * dim + i]; }
    }
    for (int i = 0; i < dim * dim; ++i) ok = ok && isfinite(Sg[i]);
  }
  if (P.debug) printf("[hess] chain %d ok %d sd0 %.3e sd1 %.3e\n", c, (int)ok, ok ? sqrt(Sg[0]) : 0.0, ok && dim > 1 ? sqrt(Sg[dim + 1]) : 0.0);
  // the pseudo-window (coordinates: the metric is still the identity, w = q)
  ch.s(S_W_N) = ok ? HESS_N : 0.0;
  if (P.dense) {
    double dq[MAXDIM];
    const double* mu = P.metric;
    for (int i = 0; i < dim; ++i) { dq[i] = ch.v(V_Q_CUR, i) - mu[i]; ch.v(V_W_MEAN, i) = ok ? HESS_N * dq[i] : 0.0; }
    int t = 0;
    for (int i = 0; i < dim; ++i)
      for (int j = 0; j <= i; ++j, ++t) ch.cross(t) = ok ? HESS_N * (Sg[i * dim + j] + dq[i] * dq[j]) : 0.0;
  } else {
    for (int i = 0; i < dim; ++i) { ch.v(V_W_MEAN, i) = ok ? ch.v(V_Q_CUR, i) : 0.0; ch.v(V_W_M2, i) = ok ? HESS_N * Sg[i * dim + i] : 0.0; }
  }
  ch.i(I_LEAF) = 0; ch.i(I_DEPTH) = 0;
  ch.i(I_PSEUDO) = 1;
  ch.i(I_PHASE) = PH_WAIT;
  atomicAdd(P.counters + 1, 1);
}

__device__ void end_opt(const Chain& ch) {
  if (ch.P.debug) {
    double gmax = 0.0;
    for (int k = 0; k < ch.P.dim; ++k) gmax = fmax(gmax, fabs(ch.v(V_G_CUR, k)));
    printf("[opt] chain %d evals %d lp %.6f gmax %.3e alpha %.3e fails %d pairs %d flat %.0f q0 %.4f q2 %.4f\n", ch.c, ch.i(I_LEAF), ch.s(S_LP_CUR), gmax,
           ch.s(S_OPT_LR), ch.i(I_DEPTH), ch.i(I_NLEAP), ch.s(S_OPT_FLAT), ch.v(V_Q_CUR, 0), ch.P.dim > 2 ? ch.v(V_Q_CUR, 2) : 0.0);
  }
  ch.i(I_NLEAP) = 0; ch.i(I_DEPTH) = 0; ch.i(I_LEAF) = 0; ch.i(I_DIR) = 0;
  if (ch.P.two_stage && ch.i(I_STAGE) == 0) {
    // end of the first stage (ascent on the subsample, bpc_sample_args.init_opt_data): wait until every chain of every rank is here;
    // the exchange that follows restarts the chains that ended in a wrong basin and switches the evaluation to the full data
    ch.i(I_STAGE) = 1;
    ch.i(I_PSEUDO) = 3;
    ch.s(S_W_N) = 1.0;
    ch.i(I_PHASE) = PH_WAIT;
    atomicAdd(ch.P.counters + 1, 1);
    return;
  }
  if (ch.P.hess) { begin_hess(ch); return; }
  ch.i(I_FE_DIR) = 0;
  begin_eps_trial(ch, PH_FIND_EPS_FIRST);
}

__device__ void opt_step(const Chain& ch) {
  const Params& P = ch.P;
  const int dim = P.dim, c = ch.c;
  const double ft = P.st_eval[c] ? -INFINITY : P.lp_eval[c];
  double gt[MAXDIM];
  load_grad(ch, gt);
  const double f = ch.s(S_LP_CUR), a = ch.s(S_OPT_LR);
  bool ok = isfinite(ft) && ft >= f + 1e-4 * a * ch.s(S_OPT_GD);
  for (int k = 0; k < dim; ++k) ok = ok && isfinite(gt[k]);
  const int evals = ch.i(I_LEAF) + 1;
  ch.i(I_LEAF) = evals;
  if (ok) {
    // curvature pair of the accepted step: s = x_t - x, y = g - g_t (gradient of -lp); kept when s.y > 0
    double sy = 0.0, ss = 0.0, yy = 0.0;
    for (int k = 0; k < dim; ++k) { const double sv = a * ch.v(V_PL, k), yv = ch.v(V_G_CUR, k) - gt[k]; sy += sv * yv; ss += sv * sv; yy += yv * yv; }
    if (sy > 1e-10 * sqrt(ss * yy) && sy > 0.0) {
      const int np = ch.i(I_NLEAP), head = np == 0 ? 0 : (ch.i(I_DIR) + 1) % LBFGS_M;
      for (int k = 0; k < dim; ++k) { ch.rck(head, k) = a * ch.v(V_PL, k); ch.rsck(head, k) = ch.v(V_G_CUR, k) - gt[k]; }
      ch.i(I_DIR) = head;
      ch.i(I_NLEAP) = np < LBFGS_M ? np + 1 : LBFGS_M;
    }
    for (int k = 0; k < dim; ++k) { ch.v(V_Q_CUR, k) = ch.v(V_W_EVAL, k); ch.v(V_G_CUR, k) = gt[k]; }
    ch.s(S_LP_CUR) = ft;
    // converged: three accepted steps in a row that change lp by less than 1e-3 (a thousandth of a nat is far inside the posterior)
    ch.s(S_OPT_FLAT) = (ft - f < 1e-3) ? ch.s(S_OPT_FLAT) + 1.0 : 0.0;
    if (evals >= P.opt_steps || ch.s(S_OPT_FLAT) >= 3.0) { end_opt(ch); return; }
    opt_direction(ch);
  } else {
    ch.s(S_OPT_LR) = 0.5 * a;
    const int fails = ch.i(I_DEPTH) + 1;
    ch.i(I_DEPTH) = fails;
    if (evals >= P.opt_steps || fails > 40) { end_opt(ch); return; }
    if (fails == 12 && ch.i(I_NLEAP) > 0) { ch.i(I_NLEAP) = 0; opt_direction(ch); }   // a bad model of the curvature: start again from steepest ascent
  }
  opt_propose(ch);
}

// One chain per WARP (lane 0 works): the chains of a pass are at different points of their state machines, and 32 of them in one
// warp execute the union of all the paths taken, one after the other; alone in its warp a chain executes its own path only, and
// the 8 x more blocks spread over the SMs.
__global__ void bp_nuts_advance(Params P) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= P.C || (threadIdx.x & 31)) return;
  Chain ch(P, c);
  const int phase = ch.i(I_PHASE);
  const int dim = P.dim;
  if (phase == PH_DONE || phase == PH_WAIT || phase == PH_FAILED) return;

  if (phase == PH_INIT) {
    const double lp = P.st_eval[c] ? -INFINITY : P.lp_eval[c];
    bool ok = isfinite(lp);
    for (int k = 0; k < dim; ++k) ok = ok && isfinite(P.g_eval[(size_t)c * dim + k]);
    if (!ok) {
      // rstan retries random initial values (up to 100 attempts); user-supplied inits that fail are an error there too
      if (ch.i(I_TRIES) >= 100) { ch.i(I_PHASE) = PH_FAILED; atomicAdd(P.counters + 0, 1); atomicAdd(P.counters + 2, 1); return; }
      ch.i(I_TRIES) += 1;
      for (int k = 0; k < dim; ++k) ch.v(V_W_EVAL, k) = 4.0 * ch.uniform() - 2.0;
      publish_eval(ch);
      return;
    }
    {
      double gw[MAXDIM];
      load_grad(ch, gw);
      for (int k = 0; k < dim; ++k) { ch.v(V_Q_CUR, k) = ch.v(V_W_EVAL, k); ch.v(V_G_CUR, k) = gw[k]; }
    }
    ch.s(S_LP_CUR) = lp;
    if (P.opt_steps > 0) { begin_opt(ch); return; }
    ch.i(I_FE_DIR) = 0;
    begin_eps_trial(ch, PH_FIND_EPS_FIRST);
    return;
  }

  if (phase == PH_OPT) { opt_step(ch); return; }
  if (phase == PH_HESS) { hess_step(ch); return; }

  if (phase == PH_FIND_EPS_FIRST || phase == PH_FIND_EPS) {
    double lp_new;
    const double H = finish_leapfrog(ch, V_QR, V_PR, V_GR, 1, lp_new);
    const double dH = ch.s(S_H0) - H;           // -inf when the trial point is rejected
    const double l08 = -0.22314355131420976;    // log(0.8)
    if (phase == PH_FIND_EPS_FIRST) {
      ch.i(I_FE_DIR) = dH > l08 ? 1 : -1;
      begin_eps_trial(ch, PH_FIND_EPS);
      return;
    }
    const int d = ch.i(I_FE_DIR);
    const bool stop = (d == 1 && !(dH > l08)) || (d == -1 && !(dH < l08));
    const double e = ch.s(S_EPS);
    if (stop || e > 1e7 || e < 1e-300) {
      restart_dual_averaging(ch);
      begin_transition(ch);
      return;
    }
    ch.s(S_EPS) = d == 1 ? 2.0 * e : 0.5 * e;
    begin_eps_trial(ch, PH_FIND_EPS);
    return;
  }

  // ---- PH_NUTS: a leapfrog of the sub-tree under construction has just been evaluated ----------
  // This is the path of almost every pass.  The vectors it needs are read once, in batches of independent loads, into
  // local arrays (a loop over a run-time dim that reads and writes the state in HBM pays one L2 round trip per element:
  // the compiler cannot move the loads of element k + 1 above the stores of element k), and written back once.
  // (the scalars too: one batch of loads up front, the stores where the original sequence has them)
  const int dir = ch.i(I_DIR), leaf = ch.i(I_LEAF), depth = ch.i(I_DEPTH), nleap = ch.i(I_NLEAP), rng = ch.i(I_RNG), st_e = P.st_eval[c];
  const double eps = ch.s(S_EPS), h0 = ch.s(S_H0), sub0 = ch.s(S_SUB_LSW), sacc0 = ch.s(S_SUM_ACC), lp_e = P.lp_eval[c];
  const int fq = dir > 0 ? V_QR : V_QL, fp = dir > 0 ? V_PR : V_PL, fg = dir > 0 ? V_GR : V_GL;
  double pl[MAXDIM], ql[MAXDIM], gl[MAXDIM], ml[MAXDIM], rl[MAXDIM];
  load_grad(ch, gl);
  for (int k = 0; k < dim; ++k) {
    ql[k] = ch.v(V_W_EVAL, k);
    pl[k] = ch.v(fp, k);
    ml[k] = ch.v(V_MINV, k);
    rl[k] = ch.v(V_RHO_SUB, k);
  }
  // second half kick with the freshly evaluated gradient (finish_leapfrog): the edge becomes the new leaf
  const double e = eps * dir;
  const double lp_new = st_e ? -INFINITY : lp_e;
  double ke = 0.0;
  for (int k = 0; k < dim; ++k) {
    pl[k] += 0.5 * e * gl[k];
    ke += ml[k] * pl[k] * pl[k];
  }
  for (int k = 0; k < dim; ++k) { ch.v(fq, k) = ql[k]; ch.v(fg, k) = gl[k]; ch.v(fp, k) = pl[k]; }
  double H = -lp_new + 0.5 * ke;
  if (isnan(H)) H = INFINITY;
  ch.i(I_NLEAP) = nleap + 1;
  const double dH = h0 - H;
  if (!(H - h0 <= 1000.0)) {   // divergent (also NaN / rejected points): the sub-tree is discarded
    ch.i(I_DIVERGENT) = 1;
    end_transition(ch);
    return;
  }
  // multinomial sampling inside the sub-tree: the new leaf replaces the sub-tree's proposal w.p. w_leaf / w_subtree
  const double sub_lsw = logaddexp(sub0, dH);
  ch.s(S_SUB_LSW) = sub_lsw;
  ch.s(S_SUM_ACC) = sacc0 + fmin(1.0, exp(dH));
  double u_take;
  {   // ch.uniform() with the counter already in a register
    unsigned r[4];
    philox(P.seed, (unsigned)(P.chain_offset + c), (unsigned)rng, r);
    ch.i(I_RNG) = rng + 1;
    u_take = (((unsigned long long)r[0] << 21) ^ (r[1] >> 11)) * (1.0 / 9007199254740992.0) + (0.5 / 9007199254740992.0);
  }
  const bool take = u_take < exp(dH - sub_lsw);
  for (int k = 0; k < dim; ++k) {
    rl[k] += pl[k];
    ch.v(V_RHO_SUB, k) = rl[k];
    if (take) { ch.v(V_Q_SP, k) = ql[k]; ch.v(V_G_SP, k) = gl[k]; }
  }
  if (take) ch.s(S_LP_SP) = lp_new;
  // sub-tree U-turn checks with checkpointed partial sums (leaf index bits select the balanced sub-trees that end here)
  int idx_max = __popc((unsigned)leaf >> 1);
  bool turning = false;
  if ((leaf & 1) == 0) {
    for (int k = 0; k < dim; ++k) { ch.rck(idx_max, k) = pl[k]; ch.rsck(idx_max, k) = rl[k]; }
  } else {
    int ntrail = 0;
    for (unsigned b = (unsigned)leaf; b & 1u; b >>= 1) ++ntrail;
    const int idx_min = idx_max - ntrail + 1;
    for (int lvl = idx_max; lvl >= idx_min && !turning; --lvl) {
      double ck[MAXDIM], sk[MAXDIM];
      for (int k = 0; k < dim; ++k) { ck[k] = ch.rck(lvl, k); sk[k] = ch.rsck(lvl, k); }
      double a = 0.0, b = 0.0;
      for (int k = 0; k < dim; ++k) {
        const double r = rl[k] - sk[k] + ck[k];
        a += ml[k] * ck[k] * r;
        b += ml[k] * pl[k] * r;
      }
      turning = !(a > 0.0 && b > 0.0);
    }
  }
  ch.i(I_LEAF) = leaf + 1;
  if (turning) {   // invalid sub-tree: the transition ends with the current tree-level proposal
    end_transition(ch);
    return;
  }
  if (leaf + 1 < (1 << depth)) {   // keep extending the same sub-tree (start_leapfrog from the local copies)
    double wn[MAXDIM], qn[MAXDIM];
    for (int k = 0; k < dim; ++k) {
      const double ph = pl[k] + 0.5 * e * gl[k];
      ch.v(fp, k) = ph;
      wn[k] = ql[k] + e * ml[k] * ph;
      ch.v(V_W_EVAL, k) = wn[k];
    }
    metric_to_q(P, wn, qn);
    for (int k = 0; k < dim; ++k) P.q_eval[(size_t)c * dim + k] = qn[k];
    return;
  }
  // sub-tree complete: biased progressive sampling at the top level, merge, tree-level U-turn check
  const double lsw = ch.s(S_LSW);
  if (sub_lsw > lsw || ch.uniform() < exp(sub_lsw - lsw)) {
    for (int k = 0; k < dim; ++k) { ch.v(V_Q_PROP, k) = ch.v(V_Q_SP, k); ch.v(V_G_PROP, k) = ch.v(V_G_SP, k); }
    ch.s(S_LP_PROP) = ch.s(S_LP_SP);
  }
  ch.s(S_LSW) = logaddexp(lsw, sub_lsw);
  for (int k = 0; k < dim; ++k) ch.v(V_RHO, k) += ch.v(V_RHO_SUB, k);
  ch.i(I_DEPTH) = depth + 1;
  if (!uturn_ok(ch, V_PL, V_PR, V_RHO) || depth + 1 >= P.max_depth) {
    end_transition(ch);
    return;
  }
  begin_subtree(ch);
}

// initial positions and per-chain adaptation state
__global__ void bp_nuts_init(Params P, const double* __restrict__ inits_u) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= P.C) return;
  Chain ch(P, c);
  for (int f = 0; f < NINT; ++f) ch.i(f) = 0;
  for (int f = 0; f < NSCAL; ++f) ch.s(f) = 0.0;
  ch.i(I_PHASE) = PH_INIT;
  ch.s(S_EPS) = P.eps0;
  ch.i(I_WIN_SIZE) = P.base_window;
  ch.i(I_WIN_NEXT) = P.init_buffer + P.base_window - 1;
  for (int k = 0; k < P.dim; ++k) {
    ch.v(V_MINV, k) = 1.0;
    ch.v(V_W_MEAN, k) = 0.0;
    ch.v(V_W_M2, k) = 0.0;
    const double q0 = P.have_inits ? inits_u[(size_t)c * P.dim + k] : 4.0 * ch.uniform() - 2.0;
    P.q_eval[(size_t)c * P.dim + k] = q0;     // (the metric starts as the identity: w = q)
    ch.v(V_W_EVAL, k) = q0;
  }
  for (int idx = 0; idx < P.nsum; ++idx) ch.cross(idx) = 0.0;
  if (P.have_inits) ch.i(I_TRIES) = 100;   // user-supplied initial values are not retried
}

// ---- selection after the ascent -------------------------------------------------------------------------------------------
// From starting values as wide as uniform(0, 1) some chains end their ascent in the wrong place: at a rate bound (a local optimum
// where the logistic transform saturates), or still crawling along the posterior's ridge when the evaluations run out.  Their
// log density is thousands of nats below the others'.  Left alone they would sit there for the whole run (NUTS cannot cross
// such a gap) and, worse, poison the pooled metric.  So the Laplace pseudo-window only counts the chains of this device whose
// log density is within PSEUDO_TOL nats of the device's best ("good" chains), and the others restart from the point of a good
// chain (the (c mod n_good)-th one): a multi-start optimisation whose winners seed the sampler.  Chains that restart from the
// same point separate at once -- their random-number streams differ.
__device__ __forceinline__ double pseudo_tol(const Params& P) { return 10.0 + P.dim; }
__device__ __forceinline__ bool pseudo_cand(const Params& P, int c) {
  const int ps = P.si[(size_t)I_PSEUDO * P.C + c];
  return P.si[(size_t)I_PHASE * P.C + c] == PH_WAIT && (ps == 1 || ps == 3);
}
__device__ double pseudo_lp_max(const Params& P) {
  double m = -INFINITY;
  for (int c = 0; c < P.C; ++c)
    if (pseudo_cand(P, c) && P.sd[(size_t)S_W_N * P.C + c] > 0.0) m = fmax(m, P.sd[(size_t)S_LP_CUR * P.C + c]);
  return m;
}
__device__ __forceinline__ bool pseudo_good(const Params& P, int c, double lpmax) {
  return P.sd[(size_t)S_W_N * P.C + c] > 0.0 && P.sd[(size_t)S_LP_CUR * P.C + c] >= lpmax - pseudo_tol(P);
}

__global__ void bp_pseudo_rescue(Params P) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= P.C || !pseudo_cand(P, c)) return;
  const double lpmax = pseudo_lp_max(P);
  if (!(lpmax > -INFINITY) || pseudo_good(P, c, lpmax)) return;
  int ngood = 0;
  for (int k = 0; k < P.C; ++k) ngood += pseudo_cand(P, k) && pseudo_good(P, k, lpmax);
  if (ngood == 0) return;
  int want = c % ngood, donor = -1;
  for (int k = 0; k < P.C && donor < 0; ++k)
    if (pseudo_cand(P, k) && pseudo_good(P, k, lpmax) && want-- == 0) donor = k;
  Chain me(P, c), dn(P, donor);
  for (int k = 0; k < P.dim; ++k) { me.v(V_Q_CUR, k) = dn.v(V_Q_CUR, k); me.v(V_G_CUR, k) = dn.v(V_G_CUR, k); }
  // (S_LP_CUR is copied last by design of the good test: a rescued chain has S_W_N = 0 and is never a donor)
  me.s(S_W_N) = 0.0;
  me.s(S_LP_CUR) = dn.s(S_LP_CUR);
  if (P.debug) printf("[rescue] chain %d <- chain %d (lp %.3f)\n", c, donor, dn.s(S_LP_CUR));
}

// pooled adaptation: this rank's window statistics over all its waiting chains, fixed order.  Layout (npool = nstat + 2 doubles):
//   [0] count;  diagonal metric: [1 .. dim] sum q, [dim + 1 .. 2 dim] sum q^2;  dense metric: [1 .. dim] sum dq, then sum dq dq^T
//   (lower triangle, row-major), dq = q - mu of the metric in force;  [nstat + 1] 1.0 when every chain of this rank has finished.
__global__ void bp_pool_stats(Params P, double* __restrict__ out, int with_flag) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int nstat = P.dense ? P.dim + P.nsum : 2 * P.dim;
  if (k == nstat + 1 && with_flag) { out[k] = P.counters[0] >= P.C ? 1.0 : 0.0; return; }
  if (k > nstat) return;
  double acc = 0.0;
  const double lpmax = pseudo_lp_max(P);     // (-inf unless this is the Laplace pseudo-window)
  for (int c = 0; c < P.C; ++c) {
    Chain ch(P, c);
    if (ch.i(I_PHASE) != PH_WAIT) continue;
    if (ch.i(I_PSEUDO) == 2) {           // the exchange at the end of the warm-up: count and sum of log step sizes
      if (k == 0) acc += 1.0; else if (k == 1) acc += log(ch.s(S_EPS));
      continue;
    }
    if (ch.i(I_PSEUDO) == 3) continue;    // the stage switch of the two-stage ascent: a barrier, no statistics
    if (ch.i(I_PSEUDO) == 1 && !pseudo_good(P, c, lpmax)) continue;   // the ascent of this chain ended in the wrong place
    const double n = ch.s(S_W_N);
    if (k == 0) acc += n;
    else if (P.dense) acc += k <= P.dim ? ch.v(V_W_MEAN, k - 1) : ch.cross(k - 1 - P.dim);
    else if (k <= P.dim) acc += n * ch.v(V_W_MEAN, k - 1);
    else { const double m = ch.v(V_W_MEAN, k - 1 - P.dim); acc += ch.v(V_W_M2, k - 1 - P.dim) + n * m * m; }
  }
  out[k] = acc;
}

// dense metric: the new (mu, L) from the reduced sums, one thread.  The metric in force moves to the "previous" slot (bp_pool_apply
// needs both to carry every chain's point and gradient over).  A covariance that is not positive
// definite leaves the metric as it is.
__global__ void bp_pool_metric(Params P, const double* __restrict__ red) {
  if (threadIdx.x || blockIdx.x) return;
  const int dim = P.dim, M = dim + dim * dim;
  for (int c = 0; c < P.C; ++c)          // the step-size exchange at the end of the warm-up leaves the metric alone
    if (P.si[(size_t)I_PHASE * P.C + c] == PH_WAIT) { if (P.si[(size_t)I_PSEUDO * P.C + c] >= 2) return; break; }
  // (pooled over thousands of draws the estimate needs no shrinkage towards the unit scale: Stan's 1e-3 * 5 / (n + 5) floor would
  // swamp variances of 1e-6; n / (n + 5) stays, plus a relative jitter on the diagonal)
  double* cur = P.metric;
  double* prev = P.metric + M;
  for (int i = 0; i < M; ++i) prev[i] = cur[i];
  const double n = red[0];
  if (!(n > dim + 1.0)) return;
  double m[MAXDIM];
  for (int i = 0; i < dim; ++i) m[i] = red[1 + i] / n;
  // Cholesky of the regularised covariance into a scratch copy (the third slot)
  double* Ln = P.metric + 2 * M + dim;
  int idx = 0;
  bool ok = true;
  for (int i = 0; i < dim; ++i)
    for (int j = 0; j <= i; ++j, ++idx) {
      double cov = (red[1 + dim + idx] - n * m[i] * m[j]) / (n - 1.0);
      cov = (n / (n + 5.0)) * cov * (i == j ? 1.0 + 1e-9 : 1.0);
      double a = cov;
      for (int k = 0; k < j; ++k) a -= Ln[i * dim + k] * Ln[j * dim + k];
      if (i == j) {
        if (!(a > 0.0) || !isfinite(a)) ok = false;
        Ln[i * dim + i] = ok ? sqrt(a) : 1.0;
      } else {
        Ln[i * dim + j] = a / Ln[j * dim + j];
        Ln[j * dim + i] = 0.0;
      }
    }
  if (!ok) return;
  for (int i = 0; i < dim; ++i) cur[i] += m[i];
  for (int i = 0; i < dim * dim; ++i) cur[dim + i] = Ln[i];
}

__global__ void bp_pool_apply(Params P, const double* __restrict__ red) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= P.C) return;
  Chain ch(P, c);
  if (ch.i(I_PHASE) != PH_WAIT) return;
  const double n = red[0];
  const int dim = P.dim;
  if (ch.i(I_PSEUDO) == 3) {           // stage switch: evaluate the point the first stage reached on the full data, then the second ascent
    ch.i(I_PSEUDO) = 0;
    ch.s(S_W_N) = 0.0;
    ch.i(I_TRIES) = 100;               // (no random restarts from here)
    for (int k = 0; k < dim; ++k) ch.v(V_W_EVAL, k) = ch.v(V_Q_CUR, k);
    publish_eval(ch);
    ch.i(I_PHASE) = PH_INIT;
    return;
  }
  if (ch.i(I_PSEUDO) == 2) {           // end of the warm-up: the pooled step size
    ch.i(I_PSEUDO) = 0;
    if (n > 0.0) ch.s(S_EPS) = exp(red[1] / n);
    begin_transition(ch);
    return;
  }
  if (ch.i(I_PSEUDO)) {                // the Laplace pseudo-window of hess_step: the window schedule does not advance
    ch.i(I_PSEUDO) = 0;
  } else {
    ch.i(I_WIN_COUNTER) -= 1;          // compute_next_window is defined on the counter of the window's last iteration
    compute_next_window(ch);
    ch.i(I_WIN_COUNTER) += 1;
  }
  if (P.dense) {
    // carry the current point and gradient over to the new coordinates: q = mu0 + L0 w, g_q = L0^-T g_w; w' = L1^-1 (q - mu1), g_w' = L1^T g_q
    const int M = dim + dim * dim;
    const double *mu1 = P.metric, *L1 = P.metric + dim, *mu0 = P.metric + M, *L0 = P.metric + M + dim;
    double w[MAXDIM], g[MAXDIM], q[MAXDIM], gq[MAXDIM];
    for (int k = 0; k < dim; ++k) { w[k] = ch.v(V_Q_CUR, k); g[k] = ch.v(V_G_CUR, k); }
    for (int i = 0; i < dim; ++i) { double a = mu0[i]; for (int j = 0; j <= i; ++j) a += L0[i * dim + j] * w[j]; q[i] = a; }
    for (int i = dim - 1; i >= 0; --i) { double a = g[i]; for (int j = i + 1; j < dim; ++j) a -= L0[j * dim + i] * gq[j]; gq[i] = a / L0[i * dim + i]; }
    for (int i = 0; i < dim; ++i) { double a = q[i] - mu1[i]; for (int j = 0; j < i; ++j) a -= L1[i * dim + j] * w[j]; w[i] = a / L1[i * dim + i]; }
    for (int j = 0; j < dim; ++j) { double a = 0.0; for (int i = j; i < dim; ++i) a += L1[i * dim + j] * gq[i]; g[j] = a; }
    for (int k = 0; k < dim; ++k) { ch.v(V_Q_CUR, k) = w[k]; ch.v(V_G_CUR, k) = g[k]; ch.v(V_W_MEAN, k) = 0.0; ch.v(V_MINV, k) = 1.0; }
    for (int idx = 0; idx < P.nsum; ++idx) ch.cross(idx) = 0.0;
  } else {
    for (int k = 0; k < dim; ++k) {
      const double mean = red[1 + k] / n;
      const double var = n > 1.0 ? (red[1 + dim + k] - n * mean * mean) / (n - 1.0) : 1.0;
      // (pooled over thousands of draws: Stan's shrinkage n / (n + 5) with the pooled count; its absolute 1e-3 * 5 / (n + 5) floor, meant
      // for one chain's 25 draws of unit-scale parameters, would swamp variances of 1e-6 and is replaced by a relative one)
      if (n > 1.0 && var > 0.0 && isfinite(var)) ch.v(V_MINV, k) = (n / (n + 5.0)) * var * (1.0 + 1e-9) + 1e-300;   // (else: keep the metric)
      ch.v(V_W_MEAN, k) = 0.0;
      ch.v(V_W_M2, k) = 0.0;
    }
  }
  ch.s(S_W_N) = 0.0;
  ch.i(I_FE_DIR) = 0;
  begin_eps_trial(ch, PH_FIND_EPS_FIRST);
}

// per-dimension sums over this rank's chains of (1, mean, mean^2, variance) of the saved draws of every HALF chain: rstan's summary
// Rhat (the number check_rhat reads, R/stan_utils.R:516-533) is split R-hat -- each chain is cut in two halves of floor(ns / 2)
// draws (the middle draw of an odd count is dropped) and the halves are treated as chains [recalled from Stan's chains.hpp /
// rstan 2.19 `split_potential_scale_reduction`; not checkable here].  With fewer than 4 saved draws the chain is not split.
__global__ void bp_chain_moments(Params P, double* __restrict__ out) {
  const int k = threadIdx.x;
  if (k >= P.dim) return;
  const int ns = P.iter - P.warmup;
  const int nh = ns >= 4 ? ns / 2 : ns, nhalves = ns >= 4 ? 2 : 1;
  double cnt = 0.0, sm = 0.0, sm2 = 0.0, sv = 0.0;
  for (int c = 0; c < P.C; ++c) {
    if (P.si[(size_t)I_PHASE * P.C + c] == PH_FAILED) continue;
    for (int hf = 0; hf < nhalves; ++hf) {
      const int i0 = hf == 0 ? 0 : ns - nh;
      double mean = 0.0, m2 = 0.0;
      for (int i = 0; i < nh; ++i) {
        const double x = P.draws[((size_t)c * ns + i0 + i) * P.dim + k];
        const double d = x - mean;
        mean += d / (i + 1);
        m2 += d * (x - mean);
      }
      cnt += 1.0; sm += mean; sm2 += mean * mean; sv += nh > 1 ? m2 / (nh - 1) : 0.0;
    }
  }
  out[4 * k + 0] = cnt; out[4 * k + 1] = sm; out[4 * k + 2] = sm2; out[4 * k + 3] = sv;
}

}  // namespace

struct bpc_sampler {
  bpc_model* m = nullptr;
  bpc_data* d = nullptr;
  bpc_data* d_opt = nullptr;            // two-stage ascent: the subsample the first stage is evaluated on, while stage1 is set
  bool stage1 = false;
  Params P{};
  DevBuf sd, si, q_eval, lp_eval, g_eval, st_eval, draws, d_lp, d_acc, d_eps, d_depth, d_nleap, d_div, counters, pool, moments, inits, metric;
  int npool = 0;
  long long passes = 0;
  int* h_counters = nullptr;   // pinned
  cudaStream_t stream = nullptr;
  cudaGraph_t graph = nullptr;          // POLL passes (4 kernels each) captured once and replayed: the per-pass cost of
  cudaGraphExec_t graph_exec = nullptr; // small problems is launch latency, not arithmetic
  bool graph_tried = false;
  int graph_launches = 0;               // kernels of one replay
  unsigned long long graph_ws_gen = 0;  // generation of the data handle's workspace the graph was captured under
  bool trace = std::getenv("BPC_SAMPLER_TRACE") != nullptr;
  bool trace_fine = std::getenv("BPC_SAMPLER_TRACE") != nullptr && std::atoi(std::getenv("BPC_SAMPLER_TRACE")) >= 2;
  std::chrono::steady_clock::time_point trace_t0 = std::chrono::steady_clock::now(), trace_t1 = std::chrono::steady_clock::now();
};

extern "C" {

int bpc_sampler_create(bpc_model* m, bpc_data* d, const bpc_sample_args* a, bpc_sampler** out) {
  if (!m || !d || !a || !out) return fail(BPC_E_INVALID, "null argument");
  *out = nullptr;
  if (d->model != m) return fail(BPC_E_INVALID, "data handle belongs to another model");
  const ModelSpec& sp = m->spec;
  if (sp.dim > MAXDIM) return fail(BPC_E_UNSUPPORTED, "too many parameters for the sampler");
  if (a->chains < 1 || a->iter < 1 || a->warmup < 0 || a->warmup >= a->iter) return fail(BPC_E_INVALID, "need chains >= 1 and 0 <= warmup < iter");
  if (a->max_treedepth < 1 || a->max_treedepth > MAXD) return fail(BPC_E_INVALID, "max_treedepth must be in [1, 12]");
  if (!(a->adapt_delta > 0.0 && a->adapt_delta < 1.0)) return fail(BPC_E_INVALID, "adapt_delta must be in (0, 1)");
  if (a->metric != 0 && a->metric != 1) return fail(BPC_E_INVALID, "metric must be 0 (diag_e) or 1 (dense_e)");
  if (a->metric == 1 && !a->pooled_adaptation)
    return fail(BPC_E_UNSUPPORTED, "the dense metric is estimated from the warm-up draws of all chains: it needs pooled_adaptation = 1");
  cudaError_t e;
  if ((e = cudaSetDevice(d->device)) != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
  auto s = std::make_unique<bpc_sampler>();
  s->m = m; s->d = d;
  Params& P = s->P;
  P.C = a->chains; P.dim = sp.dim; P.P = sp.nparams; P.noise = sp.kind == BPC_MULTITYPE_NOISE;
  P.max_depth = a->max_treedepth; P.iter = a->iter; P.warmup = a->warmup; P.chain_offset = a->chain_id_offset;
  P.pooled = a->pooled_adaptation ? 1 : 0; P.have_inits = a->inits ? 1 : 0;
  P.delta = a->adapt_delta; P.eps0 = a->stepsize > 0 ? a->stepsize : 1.0; P.seed = a->seed;
  P.opt_steps = a->init_opt_steps > 0 ? a->init_opt_steps : 0;
  P.dense = a->metric == 1 ? 1 : 0;
  P.debug = std::getenv("BPC_SAMPLER_DEBUG") ? 1 : 0;
  P.two_stage = (a->init_opt_data != nullptr && P.opt_steps > 0 && a->pooled_adaptation) ? 1 : 0;
  if (a->init_opt_data && (a->init_opt_data->model != m || a->init_opt_data->device != d->device))
    return fail(BPC_E_INVALID, "init_opt_data must be a data handle of the same model on the same device");
  s->d_opt = P.two_stage ? a->init_opt_data : nullptr;
  s->stage1 = P.two_stage != 0;
  P.hess = (P.opt_steps > 0 && P.pooled && sp.dim <= MAXD && !std::getenv("BPC_NO_HESS_INIT")) ? 1 : 0;
  P.nsum = P.dense ? sp.dim * (sp.dim + 1) / 2 : 0;
  // Stan's window schedule (windowed_adaptation): 75 / 25 / 50, rescaled to 15% / 75% / 10% for short warm-ups
  P.init_buffer = 75; P.term_buffer = 50; P.base_window = 25;
  if (P.warmup < 20) { P.init_buffer = P.warmup; P.term_buffer = 0; P.base_window = 0; }   // no metric adaptation
  else if (P.init_buffer + P.base_window + P.term_buffer > P.warmup) {
    P.init_buffer = (int)(0.15 * P.warmup); P.term_buffer = (int)(0.1 * P.warmup); P.base_window = P.warmup - (P.init_buffer + P.term_buffer);
  }
  for (int k = 0; k < MAXDIM; ++k) { P.has_bounds[k] = 0; P.lo[k] = 0; P.hi[k] = 0; }
  for (int k = 0; k < sp.nparams; ++k) { P.has_bounds[k] = sp.priors[k].has_bounds; P.lo[k] = sp.priors[k].lower; P.hi[k] = sp.priors[k].upper; }
  const size_t C = P.C, dim = P.dim, ns = P.iter - P.warmup;
  int rc;
#define AL(buf, bytes) if ((rc = s->buf.alloc(bytes))) return rc
  AL(sd, (NSCAL + NVEC * dim + 2 * MAXD * dim + (size_t)P.nsum) * C * 8);
  AL(si, (size_t)NINT * C * 4);
  AL(q_eval, C * dim * 8); AL(lp_eval, C * 8); AL(g_eval, C * dim * 8); AL(st_eval, C * 4);
  AL(draws, C * ns * dim * 8); AL(d_lp, C * ns * 8); AL(d_acc, C * ns * 8); AL(d_eps, C * ns * 8);
  AL(d_depth, C * ns * 4); AL(d_nleap, C * ns * 4); AL(d_div, C * ns * 4);
  s->npool = 2 + (P.dense ? (int)dim + P.nsum : 2 * (int)dim);
  AL(counters, 16); AL(pool, (size_t)s->npool * 8); AL(moments, 4 * dim * 8); AL(metric, 3 * (dim + dim * dim) * 8);
#undef AL
  P.sd = (double*)s->sd.p; P.si = (int*)s->si.p; P.q_eval = (double*)s->q_eval.p; P.lp_eval = (double*)s->lp_eval.p;
  P.g_eval = (double*)s->g_eval.p; P.st_eval = (int*)s->st_eval.p; P.draws = (double*)s->draws.p; P.d_lp = (double*)s->d_lp.p;
  P.d_acc = (double*)s->d_acc.p; P.d_eps = (double*)s->d_eps.p; P.d_depth = (int*)s->d_depth.p; P.d_nleap = (int*)s->d_nleap.p;
  P.d_div = (int*)s->d_div.p; P.counters = (int*)s->counters.p; P.metric = (double*)s->metric.p;
  if ((e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking)) != cudaSuccess) return cuda_fail(e, "cudaStreamCreate");
  if ((e = cudaMallocHost(&s->h_counters, 16)) != cudaSuccess) return cuda_fail(e, "cudaMallocHost");
  for (int i = 0; i < 4; ++i) s->h_counters[i] = 0;
  cudaMemsetAsync(s->sd.p, 0, s->sd.cap, s->stream);
  cudaMemsetAsync(s->counters.p, 0, 16, s->stream);
  {   // the metric starts as the identity (mu = 0, L = I) in all three slots
    std::vector<double> id(3 * (dim + dim * dim), 0.0);
    for (int slot = 0; slot < 3; ++slot)
      for (size_t k = 0; k < dim; ++k) id[slot * (dim + dim * dim) + dim + k * dim + k] = 1.0;
    if ((e = cudaMemcpyAsync(s->metric.p, id.data(), id.size() * 8, cudaMemcpyHostToDevice, s->stream)) != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
    cudaStreamSynchronize(s->stream);
  }
  cudaMemsetAsync(s->draws.p, 0, C * ns * dim * 8, s->stream);
  const double* inits_dev = nullptr;
  if (a->inits) {
    // constrained initial values -> unconstrained (sigma_obs starts at 0.05)
    std::vector<double> th(C * dim), u(C * dim);
    for (size_t c = 0; c < C; ++c) {
      for (int k = 0; k < sp.nparams; ++k) th[c * dim + k] = a->inits[c * sp.nparams + k];
      if (P.noise) th[c * dim + sp.nparams] = 0.05;
    }
    if ((rc = bpc_unconstrain(m, th.data(), (int)C, u.data()))) return rc;
    if ((rc = s->inits.alloc(C * dim * 8))) return rc;
    if ((e = cudaMemcpyAsync(s->inits.p, u.data(), C * dim * 8, cudaMemcpyHostToDevice, s->stream)) != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
    cudaStreamSynchronize(s->stream);
    inits_dev = (const double*)s->inits.p;
  }
  bp_nuts_init<<<(P.C + 127) / 128, 128, 0, s->stream>>>(P, inits_dev);
  g_launches += 1;
  if ((e = cudaGetLastError()) != cudaSuccess) return cuda_fail(e, "bp_nuts_init");
  *out = s.release();
  return BPC_OK;
}

void bpc_sampler_free(bpc_sampler* s) {
  if (!s) return;
  cudaSetDevice(s->d->device);
  if (s->stream) cudaStreamSynchronize(s->stream);
  if (s->graph_exec) cudaGraphExecDestroy(s->graph_exec);
  if (s->graph) cudaGraphDestroy(s->graph);
  if (s->stream) cudaStreamDestroy(s->stream);
  if (s->h_counters) cudaFreeHost(s->h_counters);
  delete s;
}

static int sampler_passes(bpc_sampler* s, int nb) {
  const Params& P = s->P;
  for (int b = 0; b < nb; ++b) {
    int rc = log_prob_dev(s->m, s->stage1 ? s->d_opt : s->d, P.q_eval, P.C, 1, P.lp_eval, P.g_eval, P.st_eval, s->stream, P.si + (size_t)I_PHASE * P.C);
    if (rc) return rc;
    bp_nuts_advance<<<(P.C + 3) / 4, 128, 0, s->stream>>>(P);   // (ordinary launch: see the measurement at bpc::pdl_launch's use in log_prob_dev)
    g_launches += 1;
  }
  return BPC_OK;
}

int bpc_sampler_run(bpc_sampler* s, long long max_passes, int* state) {
  if (!s || !state) return fail(BPC_E_INVALID, "null argument");
  cudaError_t e;
  if ((e = cudaSetDevice(s->d->device)) != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
  const Params& P = s->P;
  bpc_data* cur = s->stage1 ? s->d_opt : s->d;   // the data handle the passes evaluate (two-stage ascent: the subsample first)
  *state = 0;
  if (s->h_counters[0] >= P.C) { *state = 1; return BPC_OK; }   // (last poll: every chain has finished; nothing left to run)
  long long done_passes = 0;
  // poll the finished / waiting counters every POLL passes: cheap for large problems, and for small ones the
  // polling interval bounds the tail of idle passes
  const int POLL = 16;
  // The data handle's workspace is shared with every other entry point that takes the handle (bpc_log_prob, bpc_log_lik,
  // bpc_moments, ...): if one of them re-tiled or reallocated it since the graph was captured, the graph's baked-in pointers and
  // tiling are stale -- drop it, run a block eagerly (which rebuilds the workspace for this chain count) and capture again.  And if
  // the workspace was last used on another stream, wait for that work before touching it.
  if (s->graph_exec && s->graph_ws_gen != cur->ws_gen) {
    cudaGraphExecDestroy(s->graph_exec); s->graph_exec = nullptr;
    if (s->graph) { cudaGraphDestroy(s->graph); s->graph = nullptr; }
    s->graph_tried = false;
  }
  { const int rc = order_workspace(cur, s->stream); if (rc) return rc; }
  bool eager_block_done = s->graph_exec != nullptr;   // a block must run eagerly under the current workspace before a capture
  while (done_passes < max_passes) {
    const int nb = (int)std::min<long long>(POLL, max_passes - done_passes);
    if (nb == POLL && s->passes >= POLL && !s->graph_tried && eager_block_done && cur->ws_chains == P.C) {
      // the first block ran eagerly (module load, workspace allocation); capture the next one and replay it from now on
      s->graph_tried = true;
      if (!getenv("BPC_NO_GRAPH") && cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        const long long before = g_launches;
        const int rc = sampler_passes(s, POLL);
        cudaGraph_t g = nullptr;
        const cudaError_t ec = cudaStreamEndCapture(s->stream, &g);
        s->graph_launches = (int)(g_launches - before);
        g_launches = before;   // nothing ran during capture
        if (rc == BPC_OK && ec == cudaSuccess && g && cudaGraphInstantiate(&s->graph_exec, g, 0) == cudaSuccess) { s->graph = g; s->graph_ws_gen = cur->ws_gen; }
        else { if (g) cudaGraphDestroy(g); s->graph_exec = nullptr; cudaGetLastError(); }
      }
    }
    if (nb == POLL && s->graph_exec) {
      if ((e = cudaGraphLaunch(s->graph_exec, s->stream)) != cudaSuccess) return cuda_fail(e, "cudaGraphLaunch");
      g_launches += s->graph_launches;
    } else {
      const int rc = sampler_passes(s, nb);
      if (rc) return rc;
      eager_block_done = true;
    }
    done_passes += nb;
    s->passes += nb;
    if (s->trace_fine && s->passes <= 4096) {   // BPC_SAMPLER_TRACE=2: every block of 16 passes of the start of the run
      cudaStreamSynchronize(s->stream);
      const auto now = std::chrono::steady_clock::now();
      std::fprintf(stderr, "[sampler] passes %lld  %.1f us per pass (block of %d)%s\n", s->passes,
                   std::chrono::duration<double, std::micro>(now - s->trace_t1).count() / nb, nb, s->stage1 ? "  stage 1 (subsample)" : "");
      s->trace_t1 = std::chrono::steady_clock::now();
    }
    if (s->trace && s->passes / 2048 != (s->passes - nb) / 2048) {   // BPC_SAMPLER_TRACE: wall time per block of passes and where the chains are
      cudaStreamSynchronize(s->stream);
      const auto now = std::chrono::steady_clock::now();
      std::fprintf(stderr, "[sampler] passes %lld  %.1f us per pass  finished %d waiting %d of %d\n", s->passes,
                   std::chrono::duration<double, std::micro>(now - s->trace_t0).count() / 2048.0, s->h_counters[0], s->h_counters[1], P.C);
      s->trace_t0 = now;
    }
    if ((e = cudaMemcpyAsync(s->h_counters, P.counters, 12, cudaMemcpyDeviceToHost, s->stream)) != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync");
    if ((e = cudaStreamSynchronize(s->stream)) != cudaSuccess) return cuda_fail(e, "sampler pass");
    if (s->h_counters[0] >= P.C) { *state = 1; break; }
    if (P.pooled && s->h_counters[0] + s->h_counters[1] >= P.C && s->h_counters[1] > 0) { *state = 2; break; }
  }
  return release_workspace(cur, s->stream);   // (graph replays do not pass through log_prob_dev's own record)
}

int bpc_sampler_pool_stats(bpc_sampler* s, double* stats_dev, int count) {
  if (!s) return fail(BPC_E_INVALID, "null argument");
  if (stats_dev && count < s->npool - 1) return fail(BPC_E_INVALID, "stats buffer too small: see bpc_sampler_pool_len");
  cudaSetDevice(s->d->device);
  bp_pool_stats<<<(s->npool + 127) / 128, 128, 0, s->stream>>>(s->P, stats_dev ? stats_dev : (double*)s->pool.p, (!stats_dev || count >= s->npool) ? 1 : 0);
  g_launches += 1;
  cudaError_t e = cudaStreamSynchronize(s->stream);
  if (e != cudaSuccess) return cuda_fail(e, "bp_pool_stats");
  return BPC_OK;
}

int bpc_sampler_pool_len(const bpc_sampler* s) { return s ? s->npool : -1; }

int bpc_sampler_pool_apply(bpc_sampler* s, const double* reduced_dev) {
  if (!s) return fail(BPC_E_INVALID, "null argument");
  cudaSetDevice(s->d->device);
  const double* red = reduced_dev ? reduced_dev : (const double*)s->pool.p;
  if (s->stage1) {
    // the first exchange of a two-stage run is the stage switch: from here on the passes evaluate the full data (the graph captured
    // on the subsample's kernels and workspace is dropped)
    s->stage1 = false;
    if (s->graph_exec) { cudaGraphExecDestroy(s->graph_exec); s->graph_exec = nullptr; }
    if (s->graph) { cudaGraphDestroy(s->graph); s->graph = nullptr; }
    s->graph_tried = false;
  }
  if (s->P.hess || s->P.two_stage) { bp_pseudo_rescue<<<(s->P.C + 127) / 128, 128, 0, s->stream>>>(s->P); g_launches += 1; }
  if (s->P.dense) {
    bp_pool_metric<<<1, 32, 0, s->stream>>>(s->P, red);
    g_launches += 1;
  }
  bp_pool_apply<<<(s->P.C + 127) / 128, 128, 0, s->stream>>>(s->P, red);
  cudaMemsetAsync(s->P.counters + 1, 0, 4, s->stream);
  g_launches += 1;
  cudaError_t e = cudaStreamSynchronize(s->stream);
  if (e != cudaSuccess) return cuda_fail(e, "bp_pool_apply");
  return BPC_OK;
}

int bpc_sampler_draws(bpc_sampler* s, double* draws, double* lp, int* divergent, int* treedepth, int* n_leapfrog,
                      double* accept_stat, double* stepsize) {
  if (!s) return fail(BPC_E_INVALID, "null argument");
  cudaSetDevice(s->d->device);
  cudaStreamSynchronize(s->stream);
  const size_t C = s->P.C, ns = s->P.iter - s->P.warmup, dim = s->P.dim;
  cudaError_t e = cudaSuccess;
  auto cp = [&](void* dst, const void* src, size_t bytes) { if (dst && e == cudaSuccess) e = cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost); };
  cp(draws, s->draws.p, C * ns * dim * 8);
  cp(lp, s->d_lp.p, C * ns * 8);
  cp(divergent, s->d_div.p, C * ns * 4);
  cp(treedepth, s->d_depth.p, C * ns * 4);
  cp(n_leapfrog, s->d_nleap.p, C * ns * 4);
  cp(accept_stat, s->d_acc.p, C * ns * 8);
  cp(stepsize, s->d_eps.p, C * ns * 8);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
  return BPC_OK;
}

int bpc_sampler_chain_moments(bpc_sampler* s, double* stats_dev, int count, double* stats_host) {
  if (!s) return fail(BPC_E_INVALID, "null argument");
  if (stats_dev && count < 4 * s->P.dim) return fail(BPC_E_INVALID, "stats buffer too small: need 4*dim doubles");
  cudaSetDevice(s->d->device);
  double* dst = stats_dev ? stats_dev : (double*)s->moments.p;
  bp_chain_moments<<<1, MAXDIM, 0, s->stream>>>(s->P, dst);
  g_launches += 1;
  cudaError_t e = cudaStreamSynchronize(s->stream);
  if (e != cudaSuccess) return cuda_fail(e, "bp_chain_moments");
  if (stats_host && (e = cudaMemcpy(stats_host, dst, (size_t)4 * s->P.dim * 8, cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
  return BPC_OK;
}

int bpc_rhat(const double* red, int dim, int draws_per_chain, double* rhat) {
  if (!red || !rhat || dim < 1 || draws_per_chain < 2) return fail(BPC_E_INVALID, "bad argument");
  const int n = draws_per_chain >= 4 ? draws_per_chain / 2 : draws_per_chain;   // length of the half chains the moments are of
  for (int k = 0; k < dim; ++k) {
    const double m = red[4 * k], sm = red[4 * k + 1], sm2 = red[4 * k + 2], sv = red[4 * k + 3];
    if (m < 2) { rhat[k] = std::nan(""); continue; }
    const double B_over_n = (sm2 - sm * sm / m) / (m - 1.0);   // variance of the chain means
    const double W = sv / m;
    const double var_plus = W * (n - 1.0) / n + B_over_n;
    rhat[k] = std::sqrt(var_plus / W);
  }
  return BPC_OK;
}

long long bpc_sampler_total_leapfrogs(bpc_sampler* s) {
  if (!s) return 0;
  cudaSetDevice(s->d->device);
  cudaStreamSynchronize(s->stream);
  std::vector<double> v(s->P.C);
  if (cudaMemcpy(v.data(), s->P.sd + (size_t)S_NLEAP_TOT * s->P.C, (size_t)s->P.C * 8, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  double t = 0;
  for (double x : v) t += x;
  return (long long)t;
}

long long bpc_sampler_passes(const bpc_sampler* s) { return s ? s->passes : 0; }

int bpc_sampler_progress(bpc_sampler* s, int* iter, double* stepsize, double* lp, int* phase) {
  if (!s) return fail(BPC_E_INVALID, "null argument");
  cudaSetDevice(s->d->device);
  cudaStreamSynchronize(s->stream);
  const size_t C = s->P.C;
  cudaError_t e = cudaSuccess;
  auto cp = [&](void* dst, const void* src, size_t bytes) { if (dst && e == cudaSuccess) e = cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost); };
  cp(iter, s->P.si + (size_t)I_ITER * C, C * 4);
  cp(phase, s->P.si + (size_t)I_PHASE * C, C * 4);
  cp(stepsize, s->P.sd + (size_t)S_EPS * C, C * 8);
  cp(lp, s->P.sd + (size_t)S_LP_CUR * C, C * 8);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
  return BPC_OK;
}

int bpc_sample(bpc_model* m, bpc_data* d, const bpc_sample_args* a, double* draws, double* lp, int* divergent, int* treedepth,
               int* n_leapfrog, double* accept_stat, double* stepsize) {
  bpc_sampler* s = nullptr;
  int rc = bpc_sampler_create(m, d, a, &s);
  if (rc) return rc;
  int state = 0;
  while (state != 1) {
    if ((rc = bpc_sampler_run(s, 1 << 20, &state))) break;
    if (state == 2) {
      if ((rc = bpc_sampler_pool_stats(s, nullptr, 0))) break;
      if ((rc = bpc_sampler_pool_apply(s, nullptr))) break;
    }
  }
  if (!rc) rc = bpc_sampler_draws(s, draws, lp, divergent, treedepth, n_leapfrog, accept_stat, stepsize);
  if (!rc && s->h_counters[2] > 0) { bpc_sampler_free(s); return fail(BPC_E_INVALID, "initialization failed for some chains: log density or gradient not finite at 100 random initial values"); }
  bpc_sampler_free(s);
  return rc;
}

}  // extern "C"
