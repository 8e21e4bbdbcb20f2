// =============================================================================
// bp_kernels.cuh — hand-written fp64 kernels of libbpcuda for sm_100a.
//
// This header is compiled ONCE PER MODEL: codegen.cpp emits a prelude with the model's
// compile-time shape (BP_N types, BP_E events, BP_P parameters, ...), the user's rate
// expressions as device functions (bp_rates / bp_rates_vjp), the generator assembly as
// straight-line code (bp_assemble / bp_assemble_vjp), the prior table, and then includes this
// file; NVRTC compiles the translation unit for sm_100a.  The same header is compiled by nvcc
// into libbpcuda.so with a default prelude (the build()/ptxas -v/SASS check) and by g++ with
// BP_HOSTSIM for the per-thread core (tests only).
//
// What it computes (reference: jproney/bpinference, paths under /root/reference):
//   transform + Jacobian      declarations emitted by parse_prior, R/stan_utils.R:397-410;
//                             sigma_obs: stanfiles/multitype_noise_template.stan:103
//   generator A, C            stanfiles/multitype_template.stan:23-43
//   moments                   the ODE of multitype_template.stan:50-58 solved in closed form
//                             (see "moment system" below)
//   mu, Sigma per observation multitype_template.stan:137-149
//   multi_normal score        multitype_template.stan:152 ; noise variant multitype_noise_template.stan:153
//   one-type closed form      stanfiles/one_type_template.stan:37-46
//   priors                    `ThetaK ~ name(params)` emitted at R/stan_utils.R:395
//   log_lik                   stanfiles/multitype_loo_block.stan:31, simple_birth_death_loo_block.stan:11
// and the reverse-mode gradient of all of it w.r.t. the unconstrained parameters.
//
// Moment system.  The template integrates single-ancestor moments M (n x n) and D (n x n^2) and then
// contracts them with init_pop[k].  Because that contraction is linear, the state of ONE observation
//     y = (mu, Sigma),  mu = M^T z0,  Sigma = sum_l z0_l (D_l - m_l m_l^T)
// obeys its own closed linear ODE of dimension n + n(n+1)/2:
//     mu'    = A^T mu
//     Sigma' = A^T Sigma + Sigma A + sum_l mu_l H_l ,   H_l = C_l + diag(a_l) - a_l e_l^T - e_l a_l^T
// (a_l = row l of A), with y(0) = (z0, 0).  y(t) = exp(t L) y(0) is evaluated as the action of the
// matrix exponential on one vector: a Taylor series whose degree is chosen from ||tL|| for a 2^-55
// truncation bound, in sub-steps when ||tL|| exceeds what a series of BP_MCAP + 1 applications can cover.  The reverse
// sweep is the exact adjoint of that series (Horner form), which needs the series terms in reverse
// order: they are kept in shared memory, one column per thread.
// =============================================================================
#ifndef BP_KERNELS_CUH
#define BP_KERNELS_CUH

#ifndef BP_HD   // normally supplied by the generated prelude
#ifdef BP_HOSTSIM
#define BP_HD inline
#define BP_GLOBAL_CONST static const
#else
#define BP_HD __device__ __forceinline__
#define BP_GLOBAL_CONST __constant__ const
typedef double real;
#endif
#endif

#define BP_S (BP_N * (BP_N + 1) / 2)
#define BP_D (BP_N + BP_S)
// partial sums one CTA hands to the epilogue: lp, d/dTheta[BP_P], d/dsigma_obs, Taylor applications
#define BP_NPART (BP_P + 3)

// status bits — same values as include/bpcuda.h
#define BP_ST_NOT_PD 1
#define BP_ST_RATE_NONFINITE 2
#define BP_ST_ONETYPE_SINGULAR 4
#define BP_ST_PRIOR_SUPPORT 8

#define BP_SMAX 4096   // more sub-steps than this: the rates are treated as non-finite (chain rejected)

// index of (i,j) in the packed upper triangle of a symmetric n x n matrix
BP_HD constexpr int bp_sidx(int i, int j) {
  return i <= j ? i * BP_N - i * (i - 1) / 2 + (j - i) : j * BP_N - j * (j - 1) / 2 + (i - j);
}

BP_HD bool bp_finite(real v) {
#ifdef BP_HOSTSIM
  return std::isfinite(bp_val(v));
#else
  return isfinite(v);
#endif
}

// ---------------------------------------------------------------------------------------------
// structural sparsity of the generator: codegen sets bit (i*n+j) of BP_AMASK when A[i][j] can be non-zero for this
// model's events, bit (l*s+q) of BP_HMASK for H_l[q].  After unrolling the tests below are compile-time constants, so
// products with structural zeros (and gradient accumulators nobody reads) are never emitted.
#define BP_ANZ(i, j) (((BP_AMASK) >> ((i) * BP_N + (j))) & 1ull)
#define BP_HNZ(l, q) (((BP_HMASK) >> ((l) * BP_S + (q))) & 1ull)

// s (+)= a*b, the first product initialises s (a multiply, not a multiply-add)
BP_HD void bp_acc(real& s, bool& have, real a, real b) {
  if (!have) { s = a * b; have = true; } else { s += a * b; }
}

// L(w) = (A^T mu, A^T S + S A + sum_l mu_l H_l)        (forward moment operator)
BP_HD void bp_L_apply(const real (&A)[BP_N * BP_N], const real (&H)[BP_N * BP_S], const real (&wmu)[BP_N],
                      const real (&wS)[BP_S], real (&omu)[BP_N], real (&oS)[BP_S]) {
#pragma unroll
  for (int j = 0; j < BP_N; ++j) {
    real s = real(0.0);
    bool have = false;
#pragma unroll
    for (int i = 0; i < BP_N; ++i) if (BP_ANZ(i, j)) bp_acc(s, have, A[i * BP_N + j], wmu[i]);
    omu[j] = s;
  }
#pragma unroll
  for (int i = 0; i < BP_N; ++i) {
#pragma unroll
    for (int j = i; j < BP_N; ++j) {
      real s = real(0.0);
      bool have = false;
#pragma unroll
      for (int l = 0; l < BP_N; ++l) if (BP_HNZ(l, bp_sidx(i, j))) bp_acc(s, have, wmu[l], H[l * BP_S + bp_sidx(i, j)]);
      if (i == j) {
        real d = real(0.0);
        bool haved = false;
#pragma unroll
        for (int k = 0; k < BP_N; ++k) if (BP_ANZ(k, i)) bp_acc(d, haved, A[k * BP_N + i], wS[bp_sidx(k, i)]);
        if (haved) { if (have) s = s + real(2.0) * d; else s = d + d; }
      } else {
#pragma unroll
        for (int k = 0; k < BP_N; ++k) {
          if (BP_ANZ(k, i)) bp_acc(s, have, A[k * BP_N + i], wS[bp_sidx(k, j)]);
          if (BP_ANZ(k, j)) bp_acc(s, have, A[k * BP_N + j], wS[bp_sidx(k, i)]);
        }
      }
      oS[bp_sidx(i, j)] = s;
    }
  }
}

// L^T(z) = (A a + [<Sb, H_i>]_i , A Sb + Sb A^T)       (adjoint operator; Sb is a full symmetric
// cotangent stored packed, inner products count off-diagonals twice)
BP_HD void bp_LT_apply(const real (&A)[BP_N * BP_N], const real (&H)[BP_N * BP_S], const real (&bmu)[BP_N],
                       const real (&bS)[BP_S], real (&omu)[BP_N], real (&oS)[BP_S]) {
  real b2[BP_S];
#pragma unroll
  for (int i = 0; i < BP_N; ++i)
#pragma unroll
    for (int j = i; j < BP_N; ++j) {
      bool used = false;
#pragma unroll
      for (int l = 0; l < BP_N; ++l) used = used || BP_HNZ(l, bp_sidx(i, j));
      b2[bp_sidx(i, j)] = (i == j || !used) ? bS[bp_sidx(i, j)] : bS[bp_sidx(i, j)] + bS[bp_sidx(i, j)];
    }
#pragma unroll
  for (int i = 0; i < BP_N; ++i) {
    real s = real(0.0);
    bool have = false;
#pragma unroll
    for (int j = 0; j < BP_N; ++j) if (BP_ANZ(i, j)) bp_acc(s, have, A[i * BP_N + j], bmu[j]);
#pragma unroll
    for (int q = 0; q < BP_S; ++q) if (BP_HNZ(i, q)) bp_acc(s, have, H[i * BP_S + q], b2[q]);
    omu[i] = s;
  }
#pragma unroll
  for (int i = 0; i < BP_N; ++i) {
#pragma unroll
    for (int j = i; j < BP_N; ++j) {
      real s = real(0.0);
      bool have = false;
      if (i == j) {
#pragma unroll
        for (int k = 0; k < BP_N; ++k) if (BP_ANZ(i, k)) bp_acc(s, have, A[i * BP_N + k], bS[bp_sidx(k, i)]);
        if (have) s = s + s;
      } else {
#pragma unroll
        for (int k = 0; k < BP_N; ++k) {
          if (BP_ANZ(i, k)) bp_acc(s, have, A[i * BP_N + k], bS[bp_sidx(k, j)]);
          if (BP_ANZ(j, k)) bp_acc(s, have, A[j * BP_N + k], bS[bp_sidx(k, i)]);
        }
      }
      oS[bp_sidx(i, j)] = s;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// The term store: BP_MSLOT slots of BP_D doubles per thread in shared memory, laid out [slot][component][thread]
// (conflict-free), used as a RING: term j lives in slot j mod BP_MSLOT, so a forward pass over m terms leaves the
// last BP_MSLOT of them behind without any test in the loop.  Indexed through BP_SMEM so that no generic-to-shared
// address conversion is needed inside the loops.
#ifdef BP_HOSTSIM
#define BP_SMEM(idx) bp_host_store[(idx)]
#else
extern __shared__ double bp_smem[];
#define BP_SMEM(idx) bp_smem[(idx)]
#endif

// where the ring lives: shared memory (bp_fused: one column per thread) or a global scratch buffer (bp_grouped: the
// few series of a chain, L2-resident)
struct BpRingShared { BP_HD real& at(int i) const { return BP_SMEM(i); } };
struct BpRingGlobal { real* base; BP_HD real& at(int i) const { return base[i]; } };

template <class ST>
BP_HD void bp_store_term(ST st, int addr, int stride, const real (&wmu)[BP_N], const real (&wS)[BP_S]) {
#pragma unroll
  for (int i = 0; i < BP_N; ++i) st.at(addr + i * stride) = wmu[i];
#pragma unroll
  for (int q = 0; q < BP_S; ++q) st.at(addr + (BP_N + q) * stride) = wS[q];
}

// Taylor series of one sub-step, y <- y + sum_{j>=1} c_j u_j with the UNSCALED Krylov vectors u_0 = (smu,sS),
// u_{j+1} = L u_j and c_j = h^j / j!  (keeping u unscaled and scaling inside the accumulation makes every scaling part
// of an FMA).  Runs `napps` applications of L; `acc` = 0 leaves y untouched ((s) may alias (y)).  With STORE, u_k is
// written to ring slot k mod BP_MSLOT for k = 0 .. napps-1 (and u_napps too when store_tail).  Returns c_napps in c.
// The loop is unrolled by two so that the two term buffers alternate without register copies.
template <bool STORE, class ST>
BP_HD void bp_series_fwd(const real (&A)[BP_N * BP_N], const real (&H)[BP_N * BP_S], real h, int napps, const real (&smu)[BP_N],
                         const real (&sS)[BP_S], bool acc, real (&ymu)[BP_N], real (&yS)[BP_S], bool store_tail, ST st, int soff, int stride,
                         real& c) {
  real wmu[BP_N], wS[BP_S], lmu[BP_N], lS[BP_S];
#pragma unroll
  for (int i = 0; i < BP_N; ++i) { wmu[i] = smu[i]; lmu[i] = smu[i]; }
#pragma unroll
  for (int q = 0; q < BP_S; ++q) { wS[q] = sS[q]; lS[q] = sS[q]; }
  c = real(1.0);
  const real on = acc ? real(1.0) : real(0.0);
  const int span = BP_MSLOT * BP_D * stride;
  int addr = soff;
  for (int k = 0; k < napps; k += 2) {
    if (STORE) { bp_store_term(st, addr, stride, wmu, wS); addr += BP_D * stride; if (addr == soff + span) addr = soff; }
    bp_L_apply(A, H, wmu, wS, lmu, lS);
    c = c * (h * bp_rk[k]);
    {
      const real ca = on * c;
#pragma unroll
      for (int i = 0; i < BP_N; ++i) ymu[i] += ca * lmu[i];
#pragma unroll
      for (int q = 0; q < BP_S; ++q) yS[q] += ca * lS[q];
    }
    if (k + 1 < napps) {
      if (STORE) { bp_store_term(st, addr, stride, lmu, lS); addr += BP_D * stride; if (addr == soff + span) addr = soff; }
      bp_L_apply(A, H, lmu, lS, wmu, wS);
      c = c * (h * bp_rk[k + 1]);
      const real ca = on * c;
#pragma unroll
      for (int i = 0; i < BP_N; ++i) ymu[i] += ca * wmu[i];
#pragma unroll
      for (int q = 0; q < BP_S; ++q) yS[q] += ca * wS[q];
    }
  }
  if (STORE && store_tail) {
    if (napps & 1) bp_store_term(st, addr, stride, lmu, lS); else bp_store_term(st, addr, stride, wmu, wS);
  }
}

// Horner steps k = k_hi .. k_lo of the exact adjoint of the series, on the SCALED accumulator z_k = c_k v_k:
//     z_k = c_k b + L^T z_{k+1},     gA += u_k.mu (x) z_{k+1}.mu + 2 u_k.S z_{k+1}.S,     gH_l += u_k.mu_l z_{k+1}.S
// (b = cotangent of the sub-step's result, constant; z_m = c_m b; z_0 = cotangent of the sub-step's input).  c enters as
// c_{k_hi+1} and leaves as c_{k_lo}; hinv = 1/h.  u_k is read from ring slot k mod BP_MSLOT.
template <class ST>
BP_HD void bp_series_bwd(const real (&A)[BP_N * BP_N], const real (&H)[BP_N * BP_S], real hinv, int k_hi, int k_lo, ST st, int soff, int stride,
                         const real (&bmu)[BP_N], const real (&bS)[BP_S], real (&zmu)[BP_N], real (&zS)[BP_S], real& c,
                         real (&gA)[BP_N * BP_N], real (&gH)[BP_N * BP_S]) {
  const int step = BP_D * stride;
  int addr = soff + (k_hi % BP_MSLOT) * step;
  for (int k = k_hi; k >= k_lo; --k) {
    real umu[BP_N], uS[BP_S];
#pragma unroll
    for (int i = 0; i < BP_N; ++i) umu[i] = st.at(addr + i * stride);
#pragma unroll
    for (int q = 0; q < BP_S; ++q) uS[q] = st.at(addr + (BP_N + q) * stride);
    addr = (addr == soff) ? soff + (BP_MSLOT - 1) * step : addr - step;
#pragma unroll
    for (int i = 0; i < BP_N; ++i)
#pragma unroll
      for (int j = 0; j < BP_N; ++j)
        if (BP_ANZ(i, j)) {
          real s2 = uS[bp_sidx(i, 0)] * zS[bp_sidx(0, j)];
#pragma unroll
          for (int q = 1; q < BP_N; ++q) s2 += uS[bp_sidx(i, q)] * zS[bp_sidx(q, j)];
          gA[i * BP_N + j] += umu[i] * zmu[j];
          gA[i * BP_N + j] += real(2.0) * s2;
        }
#pragma unroll
    for (int l = 0; l < BP_N; ++l)
#pragma unroll
      for (int q = 0; q < BP_S; ++q)
        if (BP_HNZ(l, q)) gH[l * BP_S + q] += umu[l] * zS[q];
    real tmu[BP_N], tS[BP_S];
    bp_LT_apply(A, H, zmu, zS, tmu, tS);
    c = c * (bp_kp1[k] * hinv);
#pragma unroll
    for (int i = 0; i < BP_N; ++i) zmu[i] = c * bmu[i] + tmu[i];
#pragma unroll
    for (int q = 0; q < BP_S; ++q) zS[q] = c * bS[q] + tS[q];
  }
}

// ---------------------------------------------------------------------------------------------
// pop_vec[k] ~ multi_normal(mu, Sigma): lp = -1/2 log det Sigma - 1/2 r^T Sigma^-1 r  (Stan's `~` drops
// -(n/2) log 2 pi; add_const restores it for log_lik).  Writes bmu = dlp/dmu, bS = dlp/dSigma (full
// symmetric, packed).  Returns false when Sigma is not positive definite / not finite.
BP_HD bool bp_mvn_score(const real (&x)[BP_N], const real (&mu)[BP_N], const real (&S)[BP_S], real& lp, real (&bmu)[BP_N],
                        real (&bS)[BP_S]) {
  // square-root-free factorisation Sigma = L D L^T (L unit lower triangular): positive definite <=> all d_j > 0.
  // Costs n reciprocals and one logarithm (of the product of the pivots) instead of n square roots, n reciprocals and
  // n logarithms; these are long dependent chains in fp64, so they matter more than their flop count suggests.
  real Lc[BP_N][BP_N], Li[BP_N][BP_N], dj[BP_N], id[BP_N];
  bool ok = true;
#pragma unroll
  for (int j = 0; j < BP_N; ++j) {
    real v[BP_N];
    real d = S[bp_sidx(j, j)];
#pragma unroll
    for (int k = 0; k < j; ++k) { v[k] = Lc[j][k] * dj[k]; d -= Lc[j][k] * v[k]; }
    if (!(d > real(0.0)) || !bp_finite(d)) ok = false;
    dj[j] = d;
    id[j] = real(1.0) / d;
#pragma unroll
    for (int i = j + 1; i < BP_N; ++i) {
      real s = S[bp_sidx(i, j)];
#pragma unroll
      for (int k = 0; k < j; ++k) s -= Lc[i][k] * v[k];
      Lc[i][j] = s * id[j];
    }
  }
  if (!ok) return false;
  // log det = log prod d_j when the product cannot over/underflow, the sum of logs otherwise
  real prod = dj[0];
  bool safe = dj[0] > real(1e-60) && dj[0] < real(1e60);
#pragma unroll
  for (int j = 1; j < BP_N; ++j) { prod = prod * dj[j]; safe = safe && dj[j] > real(1e-60) && dj[j] < real(1e60); }
  real ld;
  if (safe) {
    ld = log(prod);
  } else {
    ld = log(dj[0]);
#pragma unroll
    for (int j = 1; j < BP_N; ++j) ld += log(dj[j]);
  }
  // inverse of the unit lower-triangular factor
#pragma unroll
  for (int j = 0; j < BP_N; ++j) {
#pragma unroll
    for (int i = j + 1; i < BP_N; ++i) {
      real s = -Lc[i][j];
#pragma unroll
      for (int k = j + 1; k < i; ++k) s -= Lc[i][k] * Li[k][j];
      Li[i][j] = s;
    }
  }
  real z[BP_N], a[BP_N];
  real q = real(0.0);
  bool fin = true;
#pragma unroll
  for (int i = 0; i < BP_N; ++i) {
    real s = x[i] - mu[i];
#pragma unroll
    for (int k = 0; k < i; ++k) s += Li[i][k] * (x[k] - mu[k]);
    z[i] = s;
    a[i] = s * id[i];
    q += s * a[i];
    if (!bp_finite(s)) fin = false;
  }
  if (!fin) return false;
  lp = real(-0.5) * (ld + q);
#pragma unroll
  for (int i = 0; i < BP_N; ++i) {
    real s = a[i];
#pragma unroll
    for (int k = i + 1; k < BP_N; ++k) s += Li[k][i] * a[k];
    bmu[i] = s;
  }
  // Sigma^-1 = Li^T D^-1 Li ; W[k][i] = Li[k][i] / d_k
  real W[BP_N][BP_N];
#pragma unroll
  for (int k = 0; k < BP_N; ++k)
#pragma unroll
    for (int i = 0; i < k; ++i) W[k][i] = Li[k][i] * id[k];
#pragma unroll
  for (int i = 0; i < BP_N; ++i)
#pragma unroll
    for (int j = i; j < BP_N; ++j) {
      real si = (i == j) ? id[j] : W[j][i];
#pragma unroll
      for (int k = j + 1; k < BP_N; ++k) si += W[k][i] * Li[k][j];
      bS[bp_sidx(i, j)] = real(0.5) * (bmu[i] * bmu[j] - si);
    }
  return true;
}

#if BP_N <= 3
// The same score for n <= 3 through the adjugate: Sigma^-1 = adj(Sigma) / det with ONE reciprocal, positive definiteness from the
// leading principal minors (which are cofactors), and the log-determinant NOT taken here: the caller multiplies the determinants of
// many observations together (bp_detprod below) and takes one logarithm at the end -- the reciprocals and the logarithm are half of
// the instructions of bp_mvn_score and its longest dependent chains.  q = r^T Sigma^-1 r; lp of the observation = -1/2 (log det + q).
// PACKED: bS is returned in packed coordinates (an off-diagonal entry stands for both symmetric positions: twice the value), which is
// what the W / S' accumulations want -- the factors 1/2 and 2 are folded into the constants instead of being applied afterwards.
template <bool PACKED = false>
BP_HD bool bp_mvn_score_small(const real (&x)[BP_N], const real (&mu)[BP_N], const real (&S)[BP_S], real& det, real& q, real (&bmu)[BP_N],
                              real (&bS)[BP_S]) {
  real inv[BP_S];
  bool ok;
#if BP_N == 1
  det = S[0];
  ok = det > real(0.0);
  inv[0] = real(1.0);
#elif BP_N == 2
  det = S[0] * S[2] - S[1] * S[1];
  ok = S[0] > real(0.0) && det > real(0.0);
  inv[0] = S[2]; inv[1] = -S[1]; inv[2] = S[0];
#else
  const real a = S[0], b = S[1], c = S[2], d = S[3], e = S[4], f = S[5];
  inv[0] = d * f - e * e;
  inv[1] = c * e - b * f;
  inv[2] = b * e - c * d;
  inv[3] = a * f - c * c;
  inv[4] = b * c - a * e;
  inv[5] = a * d - b * b;
  det = a * inv[0] + b * inv[1] + c * inv[2];
  ok = a > real(0.0) && inv[5] > real(0.0) && det > real(0.0);
#endif
  if (!ok || !bp_finite(det)) return false;
  const real idet = real(1.0) / det;
  real r[BP_N];
#pragma unroll
  for (int i = 0; i < BP_N; ++i) r[i] = x[i] - mu[i];
  q = real(0.0);
#pragma unroll
  for (int i = 0; i < BP_N; ++i) {
    real s = real(0.0);
#pragma unroll
    for (int j = 0; j < BP_N; ++j) s += inv[bp_sidx(i, j)] * r[j];
    bmu[i] = s * idet;
    q += r[i] * bmu[i];
  }
  if (!bp_finite(q)) return false;
  // dlp/dSigma = 1/2 (bmu bmu^T - Sigma^-1): one multiply and one multiply-add per entry
  const real mh = real(-0.5) * idet;
  real hb[BP_N];
#pragma unroll
  for (int i = 0; i < BP_N; ++i) hb[i] = real(0.5) * bmu[i];
#pragma unroll
  for (int i = 0; i < BP_N; ++i)
#pragma unroll
    for (int j = i; j < BP_N; ++j) {
      if (PACKED && i != j) bS[bp_sidx(i, j)] = bmu[i] * bmu[j] + (mh + mh) * inv[bp_sidx(i, j)];
      else bS[bp_sidx(i, j)] = hb[i] * bmu[j] + mh * inv[bp_sidx(i, j)];
    }
  return true;
}

#ifndef BP_HOSTSIM
// running product of determinants as (mantissa in [1, 2), exponent): prod <- prod * det, renormalised with integer operations
struct BpDetProd {
  double m;
  int e;
  double extra;   // log-determinants taken directly (determinants outside the range a product can carry)
  __device__ __forceinline__ void init() { m = 1.0; e = 0; extra = 0.0; }
  __device__ __forceinline__ void mul(double det) {
    if (det > 1e-280 && det < 1e280) {
      m *= det;
      const int hi = __double2hiint(m), ex = ((hi >> 20) & 0x7ff) - 1023;
      e += ex;
      m = __hiloint2double(hi - (ex << 20), __double2loint(m));
    } else {
      extra += log(det);
    }
  }
  __device__ __forceinline__ double logdet() const { return log(m) + (double)e * 0.69314718055994530942 + extra; }
};
#endif
#endif

// ||A||_1 (max column abs sum); the Lyapunov part of L is bounded by twice that
BP_HD real bp_norm1(const real (&A)[BP_N * BP_N]) {
  real a1 = real(0.0);
#pragma unroll
  for (int j = 0; j < BP_N; ++j) {
    real s = real(0.0);
#pragma unroll
    for (int i = 0; i < BP_N; ++i) s += fabs(A[i * BP_N + j]);
    a1 = (s > a1) ? s : a1;   // NaN in s never wins; callers test the rates for finiteness separately
  }
  return a1;
}

struct BpPlan { int s; int m; real h; };

// sub-steps and series degree for duration t: scaled norm th = 2 ||A||_1 t / s <= BP_THETA_CAP,
// applications m = 1 + min{ d : th <= bp_th[d] }  (bp_th[d]^d / d! = 2^-55).  `dhint` carries the degree found for
// the previous observation of this thread: observations are sorted by duration, so the search usually moves by 0 or 1.
BP_HD bool bp_make_plan(real a1, real t, BpPlan& p, int& dhint, int dcap = BP_MCAP, double theta_cap = BP_THETA_CAP) {
  const real xx = (a1 + a1) * t;
  if (!(xx >= real(0.0)) || !bp_finite(xx)) return false;
  real th = xx;
  p.s = 1;
  p.h = t;
  if (xx > real(theta_cap)) {
    const real sr = ceil(xx / real(theta_cap));
    if (sr > real((double)BP_SMAX)) return false;
#ifdef BP_HOSTSIM
    p.s = (int)bp_val(sr);
#else
    p.s = (int)sr;
#endif
    p.h = t / sr;
    th = xx / sr;
  }
  int d = dhint;
  while (d > 1 && !(th > real(bp_th[d - 1]))) --d;
  while (d < dcap && th > real(bp_th[d])) ++d;
  dhint = d;
  p.m = d + 1;
  return true;
}

// ---------------------------------------------------------------------------------------------
// one observation of the multitype templates: forward series, score, reverse series.
// Accumulates lp, gA, gH, sb (d/dsigma_obs), apps; ORs status.  store/stride: this thread's term column.
#if BP_KIND != 2
// Forward series from (imu, iS) over p.s sub-steps, `score` turns the final state into its cotangent (b), then the
// exact reverse sweep accumulates gA, gH.  Shared by bp_fused (state of one observation, score = multi_normal) and
// bp_grouped (single-ancestor moments, score = the cotangent reduced over the group's observations).
//
// The reverse sweep needs the series terms u_0 .. u_{m-1} of a sub-step in reverse order.  The ring keeps the last
// BP_MSLOT terms a forward pass produced: the top segment m-BP_MSLOT .. m-1 is reversed first, then for every lower
// segment lo .. hi the terms 0 .. hi are recomputed into the ring (hi extra applications of L) and reversed.  A small
// ring means few kilobytes of shared memory per thread, i.e. more warps per SM to hide the fp64 latencies, at the price
// of recomputed forward terms (tools/sweep.py: 12 slots, 2 CTAs per SM is the optimum for n = 3).  With several
// sub-steps (large ||tL||) only the one being reversed is stored; earlier ones are recomputed from the initial state.
// One forward loop and one reverse loop serve all of these cases.
template <class ST, class SC>
BP_HD bool bp_series_fwd_bwd(const real (&A)[BP_N * BP_N], const real (&H)[BP_N * BP_S], const BpPlan& p, const real (&imu)[BP_N],
                             const real (&iS)[BP_S], SC&& score, ST st, int soff, int stride, real (&gA)[BP_N * BP_N],
                             real (&gH)[BP_N * BP_S]) {
  const int m = p.m;
  const int nseg = (m + BP_MSLOT - 1) / BP_MSLOT;   // the m terms are reversed in segments of at most BP_MSLOT, top first
  real bmu[BP_N], bS[BP_S], zmu[BP_N], zS[BP_S];
  real cser = real(1.0);
  const real hinv = real(1.0) / p.h;
  bool scored = false;
  for (int target = p.s - 1; target >= 0; --target) {
    // items 0..target-1: plain sub-steps up to the one being reversed; item target: that sub-step (its pass leaves the top
    // segment in the ring); items target+1..: recompute of terms 0..hi of the next lower segment (hi applications)
    real ymu[BP_N], yS[BP_S], y0mu[BP_N], y0S[BP_S];
#pragma unroll
    for (int i = 0; i < BP_N; ++i) { ymu[i] = imu[i]; y0mu[i] = imu[i]; }
#pragma unroll
    for (int q = 0; q < BP_S; ++q) { yS[q] = iS[q]; y0S[q] = iS[q]; }
    const int nitems = target + nseg;
    for (int it = 0; it < nitems; ++it) {
      const int seg = it - target;
      const bool is_main = seg == 0, is_rec = seg > 0;
      const int hi = m - 1 - (is_rec ? seg : 0) * BP_MSLOT;
      const int lo = hi - BP_MSLOT + 1 > 0 ? hi - BP_MSLOT + 1 : 0;
      if (is_main) {
#pragma unroll
        for (int i = 0; i < BP_N; ++i) y0mu[i] = ymu[i];
#pragma unroll
        for (int q = 0; q < BP_S; ++q) y0S[q] = yS[q];
      }
      real smu[BP_N], sS[BP_S];
#pragma unroll
      for (int i = 0; i < BP_N; ++i) smu[i] = is_rec ? y0mu[i] : ymu[i];
#pragma unroll
      for (int q = 0; q < BP_S; ++q) sS[q] = is_rec ? y0S[q] : yS[q];
      real cfw;
      bp_series_fwd<true>(A, H, p.h, is_rec ? hi : m, smu, sS, !is_rec, ymu, yS, is_rec, st, soff, stride, cfw);
      if (is_main) {
        if (!scored) {
          scored = true;
          if (!score(ymu, yS, bmu, bS)) return false;
        }
        cser = cfw;   // c_m: the reverse sweep starts from z_m = c_m b
#pragma unroll
        for (int i = 0; i < BP_N; ++i) zmu[i] = cser * bmu[i];
#pragma unroll
        for (int q = 0; q < BP_S; ++q) zS[q] = cser * bS[q];
      }
      if (seg >= 0) bp_series_bwd(A, H, hinv, hi, lo, st, soff, stride, bmu, bS, zmu, zS, cser, gA, gH);
    }
    // cotangent of this sub-step's input = cotangent of the previous sub-step's result
#pragma unroll
    for (int i = 0; i < BP_N; ++i) bmu[i] = zmu[i];
#pragma unroll
    for (int q = 0; q < BP_S; ++q) bS[q] = zS[q];
  }
  return true;
}

// score of one observation given its moment state: noise term, multi_normal, cotangents
BP_HD bool bp_score_obs(const real (&xo)[BP_N], real sigma, real (&ymu)[BP_N], real (&yS)[BP_S], real& lp, real& sb, real (&bmu)[BP_N],
                        real (&bS)[BP_S]) {
#if BP_NOISE
#pragma unroll
  for (int i = 0; i < BP_N; ++i) yS[bp_sidx(i, i)] += sigma * ymu[i];
#endif
  real lpk;
  if (!bp_mvn_score(xo, ymu, yS, lpk, bmu, bS)) return false;
  lp += lpk;
#if BP_NOISE
#pragma unroll
  for (int i = 0; i < BP_N; ++i) { sb += bS[bp_sidx(i, i)] * ymu[i]; bmu[i] += sigma * bS[bp_sidx(i, i)]; }
#endif
  return true;
}

#if BP_N <= 3 && !defined(BP_HOSTSIM)
// the same through the adjugate (bp_mvn_score_small), for kernels that score many observations per thread: the quadratic forms are
// summed in qsum and the determinants multiplied up in dp; the thread's log density is -1/2 (qsum + dp.logdet()) at the end.
template <bool PACKED = false>
__device__ __forceinline__ bool bp_score_obs_small(const double (&xo)[BP_N], double sigma, double (&ymu)[BP_N], double (&yS)[BP_S], double& qsum,
                                                   BpDetProd& dp, double& sb, double (&bmu)[BP_N], double (&bS)[BP_S]) {
#if BP_NOISE
#pragma unroll
  for (int i = 0; i < BP_N; ++i) yS[bp_sidx(i, i)] += sigma * ymu[i];
#endif
  double det, q;
  if (!bp_mvn_score_small<PACKED>(xo, ymu, yS, det, q, bmu, bS)) return false;
  qsum += q;
  dp.mul(det);
#if BP_NOISE
#pragma unroll
  for (int i = 0; i < BP_N; ++i) { sb += bS[bp_sidx(i, i)] * ymu[i]; bmu[i] += sigma * bS[bp_sidx(i, i)]; }
#endif
  return true;
}
#endif

// ---------------------------------------------------------------------------------------------
// one observation of the multitype templates: forward series, score, reverse series.
// Accumulates lp, gA, gH, sb (d/dsigma_obs), apps; ORs status.  soff/stride: this thread's column of the term ring.
BP_HD void bp_obs_lpgrad(const real (&A)[BP_N * BP_N], const real (&H)[BP_N * BP_S], real a1, const real (&z0)[BP_N],
                         const real (&xo)[BP_N], real t, real sigma, int soff, int stride, real& lp, real (&gA)[BP_N * BP_N],
                         real (&gH)[BP_N * BP_S], real& sb, int& status, int& apps, int& dhint) {
  BpPlan p;
  if (!bp_make_plan(a1, t, p, dhint)) { status |= BP_ST_RATE_NONFINITE; return; }
  apps += p.s * p.m;   // algorithmic applications of L (and as many of L^T); recomputed terms are not counted
  real iS[BP_S];
#pragma unroll
  for (int q = 0; q < BP_S; ++q) iS[q] = real(0.0);
  const bool ok = bp_series_fwd_bwd(A, H, p, z0, iS,
                                    [&](real (&ymu)[BP_N], real (&yS)[BP_S], real (&bmu)[BP_N], real (&bS)[BP_S]) {
                                      return bp_score_obs(xo, sigma, ymu, yS, lp, sb, bmu, bS);
                                    },
                                    BpRingShared(), soff, stride, gA, gH);
  if (!ok) status |= BP_ST_NOT_PD;
}

// forward only: per-observation log_lik with the normalising constant (stanfiles/multitype_loo_block.stan:15-32)
BP_HD real bp_obs_loglik(const real (&A)[BP_N * BP_N], const real (&H)[BP_N * BP_S], real a1, const real (&z0)[BP_N],
                         const real (&xo)[BP_N], real t, real sigma) {
  BpPlan p;
  int dhint = 1;
  const real nan_ = real(0.0) / real(0.0);
  if (!bp_make_plan(a1, t, p, dhint)) return nan_;
  real ymu[BP_N], yS[BP_S], bmu[BP_N], bS[BP_S];
#pragma unroll
  for (int i = 0; i < BP_N; ++i) ymu[i] = z0[i];
#pragma unroll
  for (int q = 0; q < BP_S; ++q) yS[q] = real(0.0);
  real cfw;
  for (int j = 0; j < p.s; ++j) bp_series_fwd<false>(A, H, p.h, p.m, ymu, yS, true, ymu, yS, false, BpRingShared(), 0, 0, cfw);
  // (no sigma_obs term, also for the noise template: generate() appends the same LOO block to both multitype templates,
  // R/stan_utils.R:189-191, and multitype_loo_block.stan:22-31 builds sigma_t without it)
  (void)sigma;
  real lpk;
  if (!bp_mvn_score(xo, ymu, yS, lpk, bmu, bS)) return nan_;
  return lpk - real(0.5 * BP_N * 1.8378770664093454835606594728112);   // n/2 log(2 pi)
}
#endif  // BP_KIND != 2

// ---------------------------------------------------------------------------------------------
// one observation of the one-type template (stanfiles/one_type_template.stan:38-45):
//   g = b - d, e = exp(t g), mu = e z0, var = (e (b (2e - 1) - d)/g - e^2) z0, x ~ normal(mu, sqrt(var))
// rb[0], rb[1] += dlp/db, dlp/dd.
BP_HD void bp_onetype_obs(real b, real d, real t, real z0, real x, real& lp, real (&rb)[2], int& status, bool add_const) {
  const real g = b - d;
  if (g == real(0.0) || !bp_finite(g)) { status |= BP_ST_ONETYPE_SINGULAR; return; }
  const real e = exp(t * g);
  const real mu = e * z0;
  const real Nn = b * (real(2.0) * e - real(1.0)) - d;
  const real ig = real(1.0) / g;
  const real q = e * Nn * ig - e * e;
  const real var = q * z0;
  if (!(var > real(0.0)) || !bp_finite(var)) { status |= BP_ST_ONETYPE_SINGULAR; return; }
  const real iv = real(1.0) / var;
  const real res = x - mu;
  lp += real(-0.5) * log(var) - real(0.5) * res * res * iv - (add_const ? real(0.91893853320467274178) : real(0.0));
  const real dmu = res * iv;
  const real dvar = real(0.5) * (res * res * iv - real(1.0)) * iv;
  const real te = t * e;
  const real common = e * Nn * ig * ig;
  const real dq_db = (te * Nn + e * (real(2.0) * e - real(1.0) + real(2.0) * b * te)) * ig - common - real(2.0) * te * e;
  const real dq_dd = (-te * Nn - e * (real(2.0) * b * te + real(1.0))) * ig + common + real(2.0) * te * e;
  rb[0] += dmu * te * z0 + dvar * z0 * dq_db;
  rb[1] += -dmu * te * z0 + dvar * z0 * dq_dd;
}

// ---------------------------------------------------------------------------------------------
// priors: log-density kernel of Theta_k up to additive constants (Stan `~`), its derivative, support.
// ids follow ALLOWED_DISTRIBUTIONS (R/sysdata.rda, consumed at R/stan_utils.R:373-390).
#define BP_PR_NORMAL 0
#define BP_PR_EXP_MOD_NORMAL 1
#define BP_PR_SKEW_NORMAL 2
#define BP_PR_STUDENT_T 3
#define BP_PR_CAUCHY 4
#define BP_PR_DOUBLE_EXPONENTIAL 5
#define BP_PR_LOGISTIC 6
#define BP_PR_GUMBEL 7
#define BP_PR_LOGNORMAL 8
#define BP_PR_CHI_SQUARE 9
#define BP_PR_INV_CHI_SQUARE 10
#define BP_PR_SCALED_INV_CHI_SQUARE 11
#define BP_PR_EXPONENTIAL 12
#define BP_PR_GAMMA 13
#define BP_PR_INV_GAMMA 14
#define BP_PR_WEIBULL 15
#define BP_PR_FRECHET 16
#define BP_PR_RAYLEIGH 17
#define BP_PR_PARETO 18
#define BP_PR_PARETO_TYPE_2 19
#define BP_PR_BETA 20
#define BP_PR_UNIFORM 21

#ifndef BP_HOSTSIM
// log(erfc(z)) and d/dz, stable for large z through erfcx
__device__ inline void bp_log_erfc(double z, double& le, double& dle) {
  const double two_over_sqrtpi = 1.1283791670955125739;
  if (z > 0.0) {
    const double ex = erfcx(z);
    le = log(ex) - z * z;
    dle = -two_over_sqrtpi / ex;
  } else {
    const double ec = erfc(z);
    le = log(ec);
    dle = -two_over_sqrtpi * exp(-z * z) / ec;
  }
}

__device__ inline bool bp_prior(int id, double p0, double p1, double p2, double y, double& lp, double& g) {
  const double SQ2 = 1.4142135623730950488;
  bool ok = true;
  switch (id) {
    case BP_PR_NORMAL: { const double z = (y - p0) / p1; lp = -0.5 * z * z; g = -z / p1; break; }
    case BP_PR_EXP_MOD_NORMAL: {
      const double z = (p0 + p2 * p1 * p1 - y) / (SQ2 * p1); double le, dle; bp_log_erfc(z, le, dle);
      lp = -p2 * y + le; g = -p2 - dle / (SQ2 * p1); break; }
    case BP_PR_SKEW_NORMAL: {
      const double z = (y - p0) / p1; double le, dle; bp_log_erfc(-p2 * z / SQ2, le, dle);
      lp = -0.5 * z * z + le; g = -z / p1 - dle * p2 / (SQ2 * p1); break; }
    case BP_PR_STUDENT_T: { const double z = (y - p1) / p2; lp = -0.5 * (p0 + 1.0) * log1p(z * z / p0); g = -(p0 + 1.0) * z / (p2 * p0 * (1.0 + z * z / p0)); break; }
    case BP_PR_CAUCHY: { const double z = (y - p0) / p1; lp = -log1p(z * z); g = -2.0 * z / (p1 * (1.0 + z * z)); break; }
    case BP_PR_DOUBLE_EXPONENTIAL: { const double r = y - p0; lp = -fabs(r) / p1; g = (r > 0.0 ? -1.0 : (r < 0.0 ? 1.0 : 0.0)) / p1; break; }
    case BP_PR_LOGISTIC: { const double z = (y - p0) / p1; const double em = exp(-z); lp = -z - 2.0 * log1p(em); g = (-1.0 + 2.0 * em / (1.0 + em)) / p1; break; }
    case BP_PR_GUMBEL: { const double z = (y - p0) / p1; const double em = exp(-z); lp = -z - em; g = (-1.0 + em) / p1; break; }
    case BP_PR_LOGNORMAL: { ok = y > 0.0; const double ly = log(y); const double z = (ly - p0) / p1; lp = -ly - 0.5 * z * z; g = (-1.0 - z / p1) / y; break; }
    case BP_PR_CHI_SQUARE: { ok = y >= 0.0; lp = (0.5 * p0 - 1.0) * log(y) - 0.5 * y; g = (0.5 * p0 - 1.0) / y - 0.5; break; }
    case BP_PR_INV_CHI_SQUARE: { ok = y > 0.0; lp = -(0.5 * p0 + 1.0) * log(y) - 0.5 / y; g = -(0.5 * p0 + 1.0) / y + 0.5 / (y * y); break; }
    case BP_PR_SCALED_INV_CHI_SQUARE: { ok = y > 0.0; const double c = 0.5 * p0 * p1 * p1; lp = -(0.5 * p0 + 1.0) * log(y) - c / y; g = -(0.5 * p0 + 1.0) / y + c / (y * y); break; }
    case BP_PR_EXPONENTIAL: { ok = y >= 0.0; lp = -p0 * y; g = -p0; break; }
    case BP_PR_GAMMA: { ok = y >= 0.0; lp = (p0 - 1.0) * log(y) - p1 * y; g = (p0 - 1.0) / y - p1; break; }
    case BP_PR_INV_GAMMA: { ok = y > 0.0; lp = -(p0 + 1.0) * log(y) - p1 / y; g = -(p0 + 1.0) / y + p1 / (y * y); break; }
    case BP_PR_WEIBULL: { ok = y >= 0.0; const double w = pow(y / p1, p0); lp = (p0 - 1.0) * log(y) - w; g = (p0 - 1.0) / y - p0 * w / y; break; }
    case BP_PR_FRECHET: { ok = y > 0.0; const double w = pow(y / p1, -p0); lp = -(p0 + 1.0) * log(y) - w; g = -(p0 + 1.0) / y + p0 * w / y; break; }
    case BP_PR_RAYLEIGH: { ok = y >= 0.0; lp = log(y) - y * y / (2.0 * p0 * p0); g = 1.0 / y - y / (p0 * p0); break; }
    case BP_PR_PARETO: { ok = y >= p0; lp = -(p1 + 1.0) * log(y); g = -(p1 + 1.0) / y; break; }
    case BP_PR_PARETO_TYPE_2: { ok = y >= p0; lp = -(p2 + 1.0) * log1p((y - p0) / p1); g = -(p2 + 1.0) / (p1 + y - p0); break; }
    case BP_PR_BETA: { ok = (y >= 0.0) && (y <= 1.0); lp = (p0 - 1.0) * log(y) + (p1 - 1.0) * log1p(-y); g = (p0 - 1.0) / y - (p1 - 1.0) / (1.0 - y); break; }
    case BP_PR_UNIFORM: { ok = (y >= p0) && (y <= p1); lp = 0.0; g = 0.0; break; }
    default: ok = false; lp = 0.0; g = 0.0;
  }
  return ok && isfinite(lp);
}

// =============================================================================================
// kernels
// =============================================================================================

// ---- prologue: unconstrained -> constrained, one thread per chain -------------------------------
// theta[c][k] = lo + (hi - lo) inv_logit(u) for bounded parameters, u otherwise; sigma_obs = exp(u).
// (also called by bp_grouped when it does its own input / output: one launch per evaluation instead of three)
__device__ __forceinline__ void bp_transform(const double* __restrict__ u, double (&th)[BP_DIM]) {
#pragma unroll
  for (int k = 0; k < BP_P; ++k) {
    const double uk = u[k];
    double t = uk;
    if (bp_has_bounds[k]) {
      const double s = (uk >= 0.0) ? 1.0 / (1.0 + exp(-uk)) : exp(uk) / (1.0 + exp(uk));
      t = bp_lo[k] + (bp_hi[k] - bp_lo[k]) * s;
    }
    th[k] = t;
  }
#if BP_NOISE
  th[BP_P] = exp(u[BP_P]);
#endif
}

// Programmatic dependent launch (griddepcontrol): every kernel of an evaluation starts with bp_pdl_enter().  Launched with the
// programmatic-stream-serialization attribute (bpc::pdl_launch) its CTAs may become resident while the previous kernel of the stream is
// still running; `wait` returns once that kernel has completed and its writes are visible, `launch_dependents` then lets the NEXT kernel
// of the stream do the same.  Both are no-ops in a kernel launched the ordinary way.  Every CTA executes the wait before anything else
// (also the ones that return early): a grid that completes has therefore waited for its predecessor, and completion stays transitive.
__device__ __forceinline__ void bp_pdl_enter() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

extern "C" __global__ void bp_prologue(const double* __restrict__ u, int nchains, double* __restrict__ theta,
                                       int* __restrict__ status) {
  bp_pdl_enter();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nchains) return;
  double th[BP_DIM];
  bp_transform(u + (size_t)c * BP_DIM, th);
#pragma unroll
  for (int k = 0; k < BP_DIM; ++k) theta[(size_t)c * BP_DIM + k] = th[k];
  status[c] = 0;
}

// ---- end of an evaluation: priors, log-Jacobian, chain rule of the bound transform, rejection (bp_epilogue; bp_grouped) ----
// acc: the chain's sums (BP_NPART); st: its status bits so far.  One thread.
__device__ __forceinline__ void bp_finish(int c, const double (&acc)[BP_NPART], int st, const double* __restrict__ u,
                                          const double* __restrict__ theta, int jacobian, double* __restrict__ lp_out,
                                          double* __restrict__ grad_out, int* __restrict__ status, double* __restrict__ apps_out) {
  double lp = acc[0];
  double grad[BP_DIM];
#pragma unroll
  for (int k = 0; k < BP_P; ++k) {
    const double th = theta[(size_t)c * BP_DIM + k];
    const double uk = u[(size_t)c * BP_DIM + k];
    double pl, pg;
    if (!bp_prior(bp_prior_id[k], bp_prior_par[k][0], bp_prior_par[k][1], bp_prior_par[k][2], th, pl, pg)) { st |= BP_ST_PRIOR_SUPPORT; pl = 0.0; pg = 0.0; }
    lp += pl;
    double gth = acc[1 + k] + pg;
    if (bp_has_bounds[k]) {
      const double s = (uk >= 0.0) ? 1.0 / (1.0 + exp(-uk)) : exp(uk) / (1.0 + exp(uk));
      gth *= (bp_hi[k] - bp_lo[k]) * s * (1.0 - s);
      if (jacobian) {
        const double au = fabs(uk);
        lp += log(bp_hi[k] - bp_lo[k]) - au - 2.0 * log1p(exp(-au));
        gth += 1.0 - 2.0 * s;
      }
    }
    grad[k] = gth;
  }
#if BP_NOISE
  {
    // sigma_obs ~ normal(0, .1) on sigma_obs > 0 (multitype_noise_template.stan:131); sigma_obs = exp(u)
    const double sig = theta[(size_t)c * BP_DIM + BP_P];
    lp += -0.5 * (sig / 0.1) * (sig / 0.1);
    double g = (acc[BP_P + 1] - sig / 0.01) * sig;
    if (jacobian) { lp += u[(size_t)c * BP_DIM + BP_P]; g += 1.0; }
    grad[BP_P] = g;
  }
#endif
  bool bad = st != 0 || !isfinite(lp);
#pragma unroll
  for (int k = 0; k < BP_DIM; ++k) bad = bad || !isfinite(grad[k]);
  if (bad && st == 0) st = BP_ST_RATE_NONFINITE;
  status[c] = st;
  lp_out[c] = bad ? -(1.0 / 0.0) : lp;
  if (grad_out) {
#pragma unroll
    for (int k = 0; k < BP_DIM; ++k) grad_out[(size_t)c * BP_DIM + k] = bad ? 0.0 : grad[k];
  }
  if (apps_out) apps_out[c] = acc[BP_P + 2];
}

// ---- block reduction of BP_NPART partial sums, fixed order (deterministic) ----------------------
__device__ __forceinline__ double bp_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// vals: this thread's BP_NPART sums.  scratch: >= BP_NPART * (BP_THREADS/32) doubles of shared memory.
__device__ __forceinline__ void bp_block_reduce_store(double (&vals)[BP_NPART], double* scratch, double* __restrict__ out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NW = BP_THREADS / 32;
#pragma unroll
  for (int i = 0; i < BP_NPART; ++i) {
    const double s = bp_warp_sum(vals[i]);
    if (lane == 0) scratch[i * NW + warp] = s;
  }
  __syncthreads();
  if (threadIdx.x < BP_NPART) {
    double s = 0.0;
    for (int w = 0; w < NW; ++w) s += scratch[threadIdx.x * NW + w];
    out[threadIdx.x] = s;
  }
}

// sampler phases 4 (waiting at a pooled window), 5 (finished), 6 (failed): nothing reads the chain's result, so it is not evaluated --
// the passes of a run's tail, when most chains are done, cost what the chains still running cost
__device__ __forceinline__ bool bp_skip(const int* skip, int c) { return skip != nullptr && (unsigned)(skip[c] - 4) < 3u; }

struct BpObsArgs {
  double* theta;         // [C][BP_DIM]
  const double* z0;      // [BP_N][npad]   init_pop, SoA, sorted
  const double* xo;      // [BP_N][npad]   pop_vec
  const double* t;       // [npad]         duration of each observation
  const double* xv;      // [BP_KDEP][npad] dependent variables of each observation (only when BP_XDEP)
  double* part;          // [C][ntiles][BP_NPART]
  int* status;           // [C]
  double* loglik;        // [C][N] (bp_*_loglik only)
  const int* perm;       // [npad] original index of sorted observation (bp_*_loglik only)
  int nobs, npad, tile, ntiles, nchains;
  double* wpart;         // [C][ntiles][BP_WFRAG | BP_CFRAG] per-CTA partial sums of the W_j matrices (bp_fusedw / bp_fusedc -> bp_wpost)
  int wmode;             // 1: chains whose series need a single sub-step are handled by bp_fusedw + bp_wpost; 2: by bp_ctable + bp_fusedc + bp_wpost
  // centred W path (wmode 2): groups of consecutive 32-observation chunks of the sorted table that share a series centre
  int cgroups, cgpc;     // number of groups; groups per CTA
  const int* cg_chunk0;  // [cgroups + 1] first chunk of each group
  const int* cg_idx;     // [cgroups] centre of the group on the fine time grid: tau_g = cg_idx * cdelta
  const double* cg_tau;  // [cgroups] tau_g
  const double* cg_hmax; // [cgroups] max |t_k - tau_g| over the group's observations
  const double* cg_coef; // [cgroups][BP_CCTAB] c_i(tau_g) = tau_g^i / i! at index i + 8 (i = 0 .. BP_WCAP + 1), zero elsewhere
  double cdelta;         // fine grid spacing = longest duration / (BP_CGA * BP_CGB)
  double* ctab_m1;       // [C][BP_CGA + 1][D * n]  single-ancestor moments exp(a * BP_CGB * cdelta * L) [e_1 .. e_n]
  double* ctab_e2;       // [C][BP_CGB][D * D]      exp(b * cdelta * L)
  double* csg;           // [C][cgroups][BP_CSG]    scratch: the groups' S^g between a warp's main loop and its Toeplitz products
  // octet path (wmode 3, bp_otable + bp_fusedo + bp_toep + bp_wpost): eight chains share the tensor tiles of a pass
  int noct;              // octets = ceil(C / 8)
  int* cg_jm;            // [noct * 8][cgroups]             J | m << 8 of (chain, group); -1: the chain is not on this path
  double* cg_flops;      // [noct * 8][cgroups]             algorithmic flops of (chain, group)
  double* osg;           // [noct][cgroups][BP_XR][8 BP_D]  S'^g of the octet: row (j, l), column cc * BP_D + d
  double* wplain;        // [C][BP_CROWS][BP_WCOLS]         W_n of each chain, row n - 1, column d * BP_N + l
  int* fb;               // [1 + C]  wmode >= 2: number and ids of the chains left to bp_fused (written by bp_ctable)
  double* ctab_l;        // [noct * 8][BP_D * BP_D + 1]  octet path: L of the chain, then ||A||_1 (-1: the chain is not on this path)
  // own input / output (bp_onetype with one tile per chain: one launch per evaluation): u != nullptr
  const double* u;       // [C][BP_DIM] unconstrained parameters; theta is written, not read
  int jacobian;
  double* lp_out; double* grad_out; double* apps_out;
  double* n0tab;         // [C][cgroups][BP_D * BP_N]  octet path: N_0(tau_g) of every (chain, group), entry d * BP_N + l (bp_otable)
  const int* skip;       // [C] or nullptr: the sampler's phase of each chain; chains that wait or have finished (bp_skip) are not evaluated
  int operm;             // octet path: 1 = the columns of S' are in by-chain order (written by bp_fusedo9), 0 = (cc, d) order (bp_fusedo)
};

// ---- W path: the reverse sweep of a whole chain as one small matrix polynomial ---------------------
// When the rates do not depend on per-observation variables, every observation of a chain shares the operator L and
// (for a single sub-step) its state is a polynomial in L:  y_k = sum_j c_j(t_k) L^j y0_k,  c_j(t) = t^j / j!.  Its
// gradient w.r.t. L,  sum_k sum_j c_j(t_k) sum_i (L^T)^(j-1-i) b_k (L^i y0_k)^T,  is LINEAR in the D x n matrices
//     W_j = sum_k c_j(t_k) b_k y0_k^T          (b_k = d lp_k / d y_k in packed coordinates, j = 1..m_k)
// so the per-observation reverse sweep (its Horner recurrence with L^T, the term ring in shared memory and the
// recomputation that ring costs) is replaced by accumulating W_j over observations and ONE per-chain evaluation of
//     Lbar = sum_i Y_i P^i,   Y_i = W_(i+1) + P Y_(i+1),   P = L^T                       (bp_wpost).
// The accumulation  W[j][(d,l)] += T[j][k] * B[k][(d,l)]  over the 32 observations of a warp is a 24 x 32 x 32 matrix
// product in fp64: it runs on the tensor cores' fp64 path (mma.sync m8n8k4, DMMA in SASS), whose accumulators are
// distributed over the warp's registers, exactly what a per-thread accumulation could not afford (648 doubles).
// Operands are staged through shared memory ([row][observation], row stride BP_WLD, conflict-free fragment loads).
// Chains that need sub-steps (2 ||A||_1 t_max > BP_THETA_CAP) are left to bp_fused; the test is a function of the
// chain's parameters and the longest duration only, so both kernels take the same decision without communicating.
#define BP_WROWS (BP_WCAP + 1)              // j = 1 .. BP_WCAP + 1: without a term ring the series may be longer than bp_fused's
#define BP_WCOLS (BP_D * BP_N)             // (d, l) pairs
#define BP_WRT ((BP_WROWS + 7) / 8)        // 8-row tiles
#define BP_WCT ((BP_WCOLS + 7) / 8)        // 8-column tiles
#define BP_WFRAG (BP_WRT * BP_WCT * 64)    // accumulator elements of one warp (fragment layout)
#define BP_WLD 36                          // leading dimension of the staged operands
#define BP_WSTAGE ((BP_WRT * 8 + BP_WCT * 8) * BP_WLD)   // doubles per warp
// centred W path (bp_fusedc, below): short series of BP_CJ terms around a group centre
#define BP_CJ 8                                 // rows of the short series: j = 0 .. 7
#define BP_CROWS (BP_WCAP + 1 + BP_CJ)          // powers n = 1 .. BP_WCAP + 1 + (BP_CJ - 1), rounded to the row tile
#define BP_CRT ((BP_CROWS + 7) / 8)
#define BP_CFRAG (BP_CRT * BP_WCT * 64)
#define BP_CNLD (BP_WCOLS + (BP_WCOLS & 1))     // row stride of the staged N_j
#define BP_CCTAB (BP_CRT * 8 + 16)              // coefficient table c_i(tau_g), index i + 8, zero outside 0 .. m_g
#define BP_CGA 64                               // coarse grid intervals
#define BP_CGB 64                               // fine points per coarse interval
#define BP_CPMAX 14                             // longest series of the fine table E2 (durations <= tmax / BP_CGA)
#ifndef BP_CSPLIT
#define BP_CSPLIT 1                              // the B operand is staged in BP_CSPLIT column blocks (less shared memory per warp)
#endif
#define BP_CHCT (BP_WCT / BP_CSPLIT)            // column tiles per staged block
#define BP_CWARP (BP_CJ * BP_CNLD + (8 + BP_CHCT * 8) * BP_WLD)   // doubles of shared memory per warp
#define BP_CSG (8 * BP_WCT * 8)                 // doubles of one stored S^g: [j][8 BP_WCT columns]
// X formulation of bp_fusedc (BP_CX, default): BOTH products of an observation pass run on the fp64 tensor path.  With
// X[(j, l)][k] = c_j(h_k) z0_kl (staged once per pass):  forward  Y[d][k] = sum_(j,l) N_j[d][l] X[(j,l)][k]  (components d < 8; a
// ninth component, n = 3, is a short DFMA dot product), and  S'[(j,l)][d] += sum_k X[(j,l)][k] b_k[d]  (the same X as the A operand;
// the ninth component as T[j][k] x (b_k[8] z0_kl)).  S^g is parked with columns (l, d) instead of (d, l); bp_wpost undoes that.
#ifndef BP_CX
#define BP_CX 1
#endif
#define BP_XR (BP_CJ * BP_N)                    // rows of X: (j, l), j-major
#define BP_XKS (BP_XR / 4)                      // k-steps of the forward product
#define BP_XRT (BP_XR / 8)                      // row tiles of the S' update
#define BP_XDM (BP_D < 8 ? BP_D : 8)            // components on the tensor path
#define BP_XDR (BP_D - BP_XDM)                  // remaining components (1 for n = 3)
#define BP_XNLD 12                              // row stride of the staged N: [(j, l)][d < 8]  (conflict-free A fragments)
#define BP_XYLD 40                              // row stride of Y[d][obs] (conflict-free 16-byte stores of the accumulator fragments)
#define BP_XWARP (BP_XR * BP_XNLD + BP_XR * (BP_XDR ? BP_XDR : 1) + BP_XR * BP_WLD + 8 * BP_XYLD)   // doubles per warp

BP_HD bool bp_w_path(bool rates_finite, real a1, real tmax) {
  const real xx = (a1 + a1) * tmax;
  return rates_finite && xx >= real(0.0) && xx <= real(BP_WTHETA_CAP);
}

#if !defined(BP_HOSTSIM) && BP_KIND != 2 && !BP_XDEP
__device__ __forceinline__ void bp_dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

extern "C" __global__ void __launch_bounds__(BP_THREADS, BP_WMINBLOCKS) bp_fusedw(BpObsArgs a) {
  bp_pdl_enter();
  const int c = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (bp_skip(a.skip, c)) return;
  double th[BP_DIM];
#pragma unroll
  for (int k = 0; k < BP_DIM; ++k) th[k] = a.theta[(size_t)c * BP_DIM + k];
#if BP_NOISE
  const double sigma = th[BP_P];
#else
  const double sigma = 0.0;
#endif
  double A[BP_N * BP_N], H[BP_N * BP_S], r[BP_E], xk[BP_KDEP];
#pragma unroll
  for (int k = 0; k < BP_KDEP; ++k) xk[k] = 0.0;
  bp_rates(th, xk, r);
  bool fin = true;
#pragma unroll
  for (int e = 0; e < BP_E; ++e) fin = fin && isfinite(r[e]);
  bp_assemble(r, A, H);
  const double a1 = bp_norm1(A);
  if (!bp_w_path(fin, a1, a.nobs > 0 ? a.t[0] : 0.0)) return;   // bp_fused handles this chain

  double* sT = bp_smem + (size_t)warp * BP_WSTAGE;   // [BP_WRT*8][BP_WLD]  T[j-1][obs] = c_j(t_obs)
  double* sB = sT + BP_WRT * 8 * BP_WLD;             // [BP_WCT*8][BP_WLD]  B[(d,l)][obs] = b_d z0_l
  for (int i = lane; i < BP_WSTAGE; i += 32) sT[i] = 0.0;   // padding rows / columns stay zero
  __syncwarp();
  double acc[BP_WRT * BP_WCT][2];
#pragma unroll
  for (int i = 0; i < BP_WRT * BP_WCT; ++i) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
  double lp = 0.0, sb = 0.0;
  int status = 0, apps = 0, dhint = 1;
  const int k0 = blockIdx.x * a.tile;
  const int k1 = min(a.nobs, k0 + a.tile);
  for (int kb = k0 + warp * 32; kb < k1; kb += BP_THREADS) {   // warp-uniform: 32 consecutive observations per pass
    const int k = kb + lane;
    bool valid = k < k1;
    int m = 0;
    double t = 0.0, z0[BP_N], bmu[BP_N], bS[BP_S];
#pragma unroll
    for (int i = 0; i < BP_N; ++i) { z0[i] = 0.0; bmu[i] = 0.0; }
#pragma unroll
    for (int q = 0; q < BP_S; ++q) bS[q] = 0.0;
    if (valid) {
      double xo[BP_N], ymu[BP_N], yS[BP_S];
#pragma unroll
      for (int i = 0; i < BP_N; ++i) { z0[i] = a.z0[(size_t)i * a.npad + k]; xo[i] = a.xo[(size_t)i * a.npad + k]; ymu[i] = z0[i]; }
#pragma unroll
      for (int q = 0; q < BP_S; ++q) yS[q] = 0.0;
      t = a.t[k];
      BpPlan p;
      if (!bp_make_plan(a1, t, p, dhint, BP_WCAP, BP_WTHETA_CAP)) { status |= BP_ST_RATE_NONFINITE; valid = false; }
      else {
        m = p.m;
        apps += p.m;
        double cfw;
        bp_series_fwd<false>(A, H, p.h, p.m, ymu, yS, true, ymu, yS, false, BpRingShared(), 0, 0, cfw);
        if (!bp_score_obs(xo, sigma, ymu, yS, lp, sb, bmu, bS)) { status |= BP_ST_NOT_PD; valid = false; }
      }
    }
    // stage this observation's operands: T column (c_j(t), j = 1..m) and B column (packed-coordinate cotangent x z0)
    {
      double cj = 1.0;
#pragma unroll
      for (int j = 0; j < BP_WROWS; ++j) {
        cj = cj * (t * bp_rk[j]);
        sT[j * BP_WLD + lane] = (valid && j < m) ? cj : 0.0;
      }
#pragma unroll
      for (int d = 0; d < BP_D; ++d) {
        double bd;
        if (d < BP_N) bd = bmu[d];
        else {
          // packed coordinates: an off-diagonal entry stands for both symmetric positions
          bool diag = false;
#pragma unroll
          for (int i = 0; i < BP_N; ++i) diag = diag || (bp_sidx(i, i) == d - BP_N);
          bd = diag ? bS[d - BP_N] : bS[d - BP_N] + bS[d - BP_N];
        }
        if (!valid) bd = 0.0;
#pragma unroll
        for (int l = 0; l < BP_N; ++l) sB[(d * BP_N + l) * BP_WLD + lane] = bd * z0[l];
      }
    }
    __syncwarp();
    const int mmax = __reduce_max_sync(0xffffffffu, m);   // row tiles above the warp's longest series hold zeros: skip them
    const int g = lane >> 2, tg = lane & 3;
#pragma unroll
    for (int ob = 0; ob < 8; ++ob) {   // 4 observations per mma
      double bf[BP_WCT];
#pragma unroll
      for (int ct = 0; ct < BP_WCT; ++ct) bf[ct] = sB[(ct * 8 + g) * BP_WLD + ob * 4 + tg];
#pragma unroll
      for (int rt = 0; rt < BP_WRT; ++rt) {
        if (rt * 8 < mmax) {
          const double af = sT[(rt * 8 + g) * BP_WLD + ob * 4 + tg];
#pragma unroll
          for (int ct = 0; ct < BP_WCT; ++ct) bp_dmma(acc[rt * BP_WCT + ct][0], acc[rt * BP_WCT + ct][1], af, bf[ct]);
        }
      }
    }
    __syncwarp();
  }
  // CTA partial of W: sum the warps' accumulators in warp order (fragment layout kept; bp_wpost knows it)
  __syncthreads();
  double* red = bp_smem;   // [warps][BP_WFRAG]
#pragma unroll
  for (int i = 0; i < BP_WRT * BP_WCT; ++i) { red[(size_t)warp * BP_WFRAG + i * 64 + lane * 2] = acc[i][0]; red[(size_t)warp * BP_WFRAG + i * 64 + lane * 2 + 1] = acc[i][1]; }
  __syncthreads();
  double* wout = a.wpart + ((size_t)c * a.ntiles + blockIdx.x) * BP_WFRAG;
  for (int i = tid; i < BP_WFRAG; i += BP_THREADS) {
    double sacc = 0.0;
#pragma unroll
    for (int w = 0; w < BP_THREADS / 32; ++w) sacc += red[(size_t)w * BP_WFRAG + i];
    wout[i] = sacc;
  }
  __syncthreads();
  double vals[BP_NPART];
  vals[0] = lp;
#pragma unroll
  for (int k = 0; k < BP_P; ++k) vals[1 + k] = 0.0;
  vals[BP_P + 1] = sb;
  vals[BP_P + 2] = (double)apps;
  bp_block_reduce_store(vals, bp_smem, a.part + ((size_t)c * a.ntiles + blockIdx.x) * BP_NPART);
  if (status) atomicOr(a.status + c, status);
}

#endif  // W path

#if !defined(BP_HOSTSIM) && BP_KIND != 2
// ---- cooperative (whole CTA) pieces of the per-chain matrix polynomial, shared by bp_wpost and bp_grouped ----------
// column cc of L in packed coordinates = L applied to the packed unit vector e_cc; Lm[r * BP_D + cc]
__device__ __forceinline__ void bp_coop_build_L(const double (&A)[BP_N * BP_N], const double (&H)[BP_N * BP_S], double* Lm) {
  for (int cc = threadIdx.x; cc < BP_D; cc += blockDim.x) {
    double umu[BP_N], uS[BP_S], omu[BP_N], oS[BP_S];
#pragma unroll
    for (int i = 0; i < BP_N; ++i) umu[i] = (cc == i) ? 1.0 : 0.0;
#pragma unroll
    for (int q = 0; q < BP_S; ++q) uS[q] = (cc == BP_N + q) ? 1.0 : 0.0;
    bp_L_apply(A, H, umu, uS, omu, oS);
#pragma unroll
    for (int i = 0; i < BP_N; ++i) Lm[i * BP_D + cc] = omu[i];
#pragma unroll
    for (int q = 0; q < BP_S; ++q) Lm[(BP_N + q) * BP_D + cc] = oS[q];
  }
}

// Lbar = sum_i Y_i P^i with Y_i = W_(i+1) + P Y_(i+1), P = L^T.  W: M rows of D x n ([d * n + l]) in shared memory
// (overwritten by Y), Lm: [D][D], Z: 2 * D * D scratch.  All threads of the CTA must call; returns the index (0/1) of the
// Z buffer that holds Lbar[r * D + c] = d lp / d L[r][c].
__device__ __forceinline__ int bp_coop_lbar(double* W, int M, const double* Lm, double* Z) {
  constexpr int DN = BP_D * BP_N, DD = BP_D * BP_D, MAXO = (DD + 31) / 32;
  const int tid = threadIdx.x, T = blockDim.x;
  for (int i = M - 2; i >= 0; --i) {
    double tmp[MAXO];
    int cnt = 0;
    for (int o = tid; o < DN; o += T, ++cnt) {
      const int d = o / BP_N, l = o - d * BP_N;
      double y = W[i * DN + o];
#pragma unroll
      for (int rr = 0; rr < BP_D; ++rr) y += Lm[rr * BP_D + d] * W[(i + 1) * DN + rr * BP_N + l];
      tmp[cnt] = y;
    }
    __syncthreads();
    cnt = 0;
    for (int o = tid; o < DN; o += T, ++cnt) W[i * DN + o] = tmp[cnt];
    __syncthreads();
  }
  int cur = 0;
  for (int o = tid; o < DD; o += T) {
    const int d = o / BP_D, e = o - d * BP_D;
    Z[o] = (e < BP_N && M > 0) ? W[(M - 1) * DN + d * BP_N + e] : 0.0;
  }
  __syncthreads();
  for (int i = M - 2; i >= 0; --i) {
    for (int o = tid; o < DD; o += T) {
      const int d = o / BP_D, e = o - d * BP_D;
      double z = (e < BP_N) ? W[i * DN + d * BP_N + e] : 0.0;
#pragma unroll
      for (int rr = 0; rr < BP_D; ++rr) z += Z[cur * DD + d * BP_D + rr] * Lm[e * BP_D + rr];
      Z[(cur ^ 1) * DD + o] = z;
    }
    __syncthreads();
    cur ^= 1;
  }
  return cur;
}

// The same polynomial for W_j with D columns (rows of D x D in shared memory, overwritten): the reverse sweep of a series whose
// inputs are full state vectors, as the sub-steps of bp_grouped's cooperative multi-step path are.
__device__ __forceinline__ int bp_coop_lbar_full(double* W, int M, const double* Lm, double* Z) {
  constexpr int DD = BP_D * BP_D;
  const int tid = threadIdx.x, T = blockDim.x;
  for (int i = M - 2; i >= 0; --i) {
    for (int o = tid; o < DD; o += T) {
      const int d = o / BP_D, e = o - d * BP_D;
      double y = W[i * DD + o];
#pragma unroll
      for (int rr = 0; rr < BP_D; ++rr) y += Lm[rr * BP_D + d] * W[(i + 1) * DD + rr * BP_D + e];
      W[i * DD + o] = y;
    }
    __syncthreads();
  }
  int cur = 0;
  for (int o = tid; o < DD; o += T) Z[o] = M > 0 ? W[(M - 1) * DD + o] : 0.0;
  __syncthreads();
  for (int i = M - 2; i >= 0; --i) {
    for (int o = tid; o < DD; o += T) {
      const int d = o / BP_D, e = o - d * BP_D;
      double z = W[i * DD + o];
#pragma unroll
      for (int rr = 0; rr < BP_D; ++rr) z += Z[cur * DD + d * BP_D + rr] * Lm[e * BP_D + rr];
      Z[(cur ^ 1) * DD + o] = z;
    }
    __syncthreads();
    cur ^= 1;
  }
  return cur;
}

// Lbar (packed coordinates) -> (gA, gH) with the accumulation rule of the reverse sweep: input u = e_cc, output
// cotangent = column cc of Lbar converted to the full symmetric convention (off-diagonals halved).  One thread.
__device__ __forceinline__ void bp_lbar_to_grad(const double* Zc, double (&gA)[BP_N * BP_N], double (&gH)[BP_N * BP_S]) {
  for (int cc = 0; cc < BP_D; ++cc) {
    double umu[BP_N], uS[BP_S], zmu[BP_N], zS[BP_S];
#pragma unroll
    for (int i = 0; i < BP_N; ++i) { umu[i] = (cc == i) ? 1.0 : 0.0; zmu[i] = Zc[i * BP_D + cc]; }
#pragma unroll
    for (int i = 0; i < BP_N; ++i)
#pragma unroll
      for (int j = i; j < BP_N; ++j) {
        const int q = bp_sidx(i, j);
        uS[q] = (cc == BP_N + q) ? 1.0 : 0.0;
        zS[q] = (i == j) ? Zc[(BP_N + q) * BP_D + cc] : 0.5 * Zc[(BP_N + q) * BP_D + cc];
      }
#pragma unroll
    for (int i = 0; i < BP_N; ++i)
#pragma unroll
      for (int j = 0; j < BP_N; ++j)
        if (BP_ANZ(i, j)) {
          double s2 = 0.0;
#pragma unroll
          for (int q = 0; q < BP_N; ++q) s2 += uS[bp_sidx(i, q)] * zS[bp_sidx(q, j)];
          gA[i * BP_N + j] += umu[i] * zmu[j] + 2.0 * s2;
        }
#pragma unroll
    for (int l = 0; l < BP_N; ++l)
#pragma unroll
      for (int q = 0; q < BP_S; ++q)
        if (BP_HNZ(l, q)) gH[l * BP_S + q] += umu[l] * zS[q];
  }
}

// per chain: W_j summed over tiles -> Lbar -> (gA, gH) -> Theta-bar, added to the chain's first tile partial
#if !BP_XDEP
extern "C" __global__ void __launch_bounds__(BP_WPOST_THREADS) bp_wpost(BpObsArgs a) {
  bp_pdl_enter();
  const int c = blockIdx.x, tid = threadIdx.x;
  if (bp_skip(a.skip, c)) return;
  __shared__ double W[BP_CROWS * BP_WCOLS];   // W_(j+1) as [j][d * n + l]; then Y_j in place (BP_WROWS rows from bp_fusedw, BP_CROWS from bp_fusedc)
  __shared__ double Lm[BP_D * BP_D];
  __shared__ double Z[2 * BP_D * BP_D];
  double th[BP_DIM];
#pragma unroll
  for (int k = 0; k < BP_DIM; ++k) th[k] = a.theta[(size_t)c * BP_DIM + k];
  double A[BP_N * BP_N], H[BP_N * BP_S], r[BP_E], xk[BP_KDEP];
#pragma unroll
  for (int k = 0; k < BP_KDEP; ++k) xk[k] = 0.0;
  bp_rates(th, xk, r);
  bool fin = true;
#pragma unroll
  for (int e = 0; e < BP_E; ++e) fin = fin && isfinite(r[e]);
  bp_assemble(r, A, H);
  const double a1 = bp_norm1(A);
  const bool wpath = bp_w_path(fin, a1, a.nobs > 0 ? a.t[0] : 0.0);   // (uniform over the CTA)
  if (!wpath && !(a.wmode == 3 && a.u)) return;                         // bp_fused handled this chain; bp_epilogue finishes it
  if (wpath) {
  // sum the tile partials in tile order; fragment element (rt, ct, lane, e) is W[8 rt + lane/4][8 ct + 2 (lane%4) + e]
  const int frag = a.wmode == 1 ? BP_WFRAG : BP_CFRAG, rows = a.wmode == 1 ? BP_WROWS : BP_CROWS;
#if BP_N <= 3
  if (a.wmode == 3) {   // octet path: bp_toep wrote W of this chain in place
    for (int i = tid; i < BP_CROWS * BP_WCOLS; i += blockDim.x) W[i] = a.wplain[(size_t)c * BP_CROWS * BP_WCOLS + i];
  } else
#endif
  for (int i = tid; i < frag; i += blockDim.x) {
    double sacc = 0.0;
    for (int tI = 0; tI < a.ntiles; ++tI) sacc += a.wpart[((size_t)c * a.ntiles + tI) * frag + i];
    const int tile = i >> 6, ln = (i & 63) >> 1, e = i & 1;
    const int row = (tile / BP_WCT) * 8 + (ln >> 2), col = (tile % BP_WCT) * 8 + (ln & 3) * 2 + e;
#if BP_CX && BP_N <= 3
    if (a.wmode == 2) {   // bp_fusedc (X formulation) parks columns as (l, d) for d < BP_XDM, then (q, l) for the remaining components
      int dd = -1, ll = 0;
      if (col < BP_N * 8) { ll = col >> 3; dd = (col & 7) < BP_XDM ? (col & 7) : -1; }
      else { const int rem = col - BP_N * 8, q = rem / BP_N; ll = rem - q * BP_N; dd = q < BP_XDR ? BP_XDM + q : -1; }
      if (row < rows && dd >= 0) W[row * BP_WCOLS + dd * BP_N + ll] = sacc;
      continue;
    }
#endif
    if (row < rows && col < BP_WCOLS) W[row * BP_WCOLS + col] = sacc;
  }
  bp_coop_build_L(A, H, Lm);
  __syncthreads();
  const int cur = bp_coop_lbar(W, rows, Lm, Z);
  if (tid == 0) {
    double gA[BP_N * BP_N], gH[BP_N * BP_S], rb[BP_E], cb[BP_P];
#pragma unroll
    for (int i = 0; i < BP_N * BP_N; ++i) gA[i] = 0.0;
#pragma unroll
    for (int i = 0; i < BP_N * BP_S; ++i) gH[i] = 0.0;
#pragma unroll
    for (int k = 0; k < BP_P; ++k) cb[k] = 0.0;
    bp_lbar_to_grad(Z + cur * BP_D * BP_D, gA, gH);
    bp_assemble_vjp(gA, gH, rb);
    bp_rates_vjp(th, xk, rb, cb);
    double* p0 = a.part + (size_t)c * a.ntiles * BP_NPART;
#pragma unroll
    for (int k = 0; k < BP_P; ++k) p0[1 + k] += cb[k];
  }
  }
  // octet path: the work of bp_epilogue for this chain by the CTA's first warp (one launch less per evaluation): tile sums in the same
  // fixed order (lane i takes tiles i, i + 32, ..., then the shuffle tree), priors, Jacobian, chain rule, rejection.  Chains on the
  // sub-step fallback arrive here with the partials bp_fused wrote (it is launched before this kernel).
  if (a.wmode == 3 && a.u) {
    __syncthreads();     // thread 0's update of the first tile partial is visible to the first warp
    if (tid < 32) {
      double acc[BP_NPART];
#pragma unroll
      for (int i = 0; i < BP_NPART; ++i) acc[i] = 0.0;
      for (int tI = tid; tI < a.ntiles; tI += 32) {
        const double* p = a.part + ((size_t)c * a.ntiles + tI) * BP_NPART;
#pragma unroll
        for (int i = 0; i < BP_NPART; ++i) acc[i] += __ldcg(p + i);
      }
#pragma unroll
      for (int i = 0; i < BP_NPART; ++i) acc[i] = bp_warp_sum(acc[i]);
      if (tid == 0) bp_finish(c, acc, a.status[c], a.u, a.theta, a.jacobian, a.lp_out, a.grad_out, a.status, a.apps_out);
    }
  }
}

// ---- centred W path: bp_ctable + bp_fusedc -----------------------------------------------------------
// The observation table is sorted by duration, so the 32 observations of a warp pass -- and a few consecutive
// passes -- have almost the same duration.  For a GROUP of such passes with centre tau_g (a point of a fine time
// grid) write t_k = tau_g + h_k.  Then, by the binomial theorem c_n(tau + h) = sum_j c_(n-j)(tau) c_j(h),
//     y_k = exp(t_k L) y0_k = sum_(j <= J) c_j(h_k) N_j z0_k,      N_j = L^j M(tau_g),   M(tau) = exp(tau L) [e_1 .. e_n]
// where J is the SHORT series length |h| needs (<= BP_CJ - 1 because |h_k| is a small fraction of the longest
// duration) instead of the long one t_k needs, and the W matrices of bp_fusedw factor the same way:
//     W_n = sum_g sum_j c_(n-j)(tau_g) S^g_j,        S^g_j = sum_(k in g) c_j(h_k) b_k z0_k^T      (j = 0 .. J).
// So per observation: J + 1 terms of a D x n matrix-vector product (N_j from shared memory), the score, and ONE 8-row
// tile of the S update on the fp64 tensor path; per group: M(tau_g) = E2[b] M1[a] from two small per-chain tables
// (bp_ctable; tau_g = (a BP_CGB + b) cdelta), the chain N_(j+1) = L N_j, and the Toeplitz product W += C(tau_g) S^g,
// again as DMMAs.  bp_wpost turns W into the gradient exactly as for bp_fusedw (with BP_CROWS rows).
#if BP_N <= 3

// per chain: the two tables M(tau) is assembled from.  One CTA per chain; Krylov blocks U_j = L^j [e_1 .. e_n] and powers
// P_j = L^j in shared memory, then every table entry is a short dot product with the coefficients c_j(time).
extern "C" __global__ void __launch_bounds__(512) bp_ctable(BpObsArgs a) {
  bp_pdl_enter();
  const int c = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
  if (bp_skip(a.skip, c)) return;
  constexpr int DN = BP_D * BP_N, DD = BP_D * BP_D;
  __shared__ double Lm[DD];
  __shared__ double U[(BP_WROWS + 1) * DN];
  __shared__ double Pw[(BP_CPMAX + 1) * DD];
  double th[BP_DIM];
#pragma unroll
  for (int k = 0; k < BP_DIM; ++k) th[k] = a.theta[(size_t)c * BP_DIM + k];
  double A[BP_N * BP_N], H[BP_N * BP_S], r[BP_E], xk[BP_KDEP];
#pragma unroll
  for (int k = 0; k < BP_KDEP; ++k) xk[k] = 0.0;
  bp_rates(th, xk, r);
  bool fin = true;
#pragma unroll
  for (int e = 0; e < BP_E; ++e) fin = fin && isfinite(r[e]);
  bp_assemble(r, A, H);
  const double a1 = bp_norm1(A);
  const double tmax = a.nobs > 0 ? a.t[0] : 0.0;
  if (!bp_w_path(fin, a1, tmax)) {   // bp_fused takes this chain
    if (tid == 0) {
      a.fb[1 + atomicAdd(a.fb, 1)] = c;
      if (a.wmode == 3) a.ctab_l[(size_t)c * (DD + 1) + DD] = -1.0;
    }
    return;
  }
  BpPlan p;
  int dh = 1;
  bp_make_plan(a1, tmax, p, dh, BP_WCAP, BP_WTHETA_CAP);
  const int mtop = p.m;
  const double d1 = a.cdelta * BP_CGB;
  dh = 1;
  bp_make_plan(a1, d1, p, dh, BP_CPMAX - 1, BP_WTHETA_CAP);
  const int mfine = p.m;
  bp_coop_build_L(A, H, Lm);
  if (a.wmode == 3) {
    __syncthreads();
    for (int o = tid; o < DD; o += T) a.ctab_l[(size_t)c * (DD + 1) + o] = Lm[o];
    if (tid == 0) a.ctab_l[(size_t)c * (DD + 1) + DD] = a1;
  }
  for (int o = tid; o < DN; o += T) { const int d = o / BP_N, l = o - d * BP_N; U[o] = (d == l) ? 1.0 : 0.0; }
  for (int o = tid; o < DD; o += T) { const int d = o / BP_D, e = o - d * BP_D; Pw[o] = (d == e) ? 1.0 : 0.0; }
  // series lengths of the table points, once per point (not per entry)
  __shared__ int mA[BP_CGA + 1], mB[BP_CGB];
  for (int i = tid; i < BP_CGA + 1 + BP_CGB; i += T) {
    BpPlan q;
    int dq = 1;
    if (i <= BP_CGA) { bp_make_plan(a1, i * d1, q, dq, BP_WCAP, BP_WTHETA_CAP); mA[i] = q.m; }
    else { bp_make_plan(a1, (i - BP_CGA - 1) * a.cdelta, q, dq, BP_CPMAX - 1, BP_WTHETA_CAP); mB[i - BP_CGA - 1] = q.m; }
  }
  __syncthreads();
  for (int j = 1; j <= mtop; ++j) {
    for (int o = tid; o < DN; o += T) {
      const int d = o / BP_N, l = o - d * BP_N;
      double u = 0.0;
#pragma unroll
      for (int rr = 0; rr < BP_D; ++rr) u += Lm[d * BP_D + rr] * U[(j - 1) * DN + rr * BP_N + l];
      U[j * DN + o] = u;
    }
    if (j <= mfine)
      for (int o = tid; o < DD; o += T) {
        const int d = o / BP_D, e = o - d * BP_D;
        double u = 0.0;
#pragma unroll
        for (int rr = 0; rr < BP_D; ++rr) u += Lm[d * BP_D + rr] * Pw[(j - 1) * DD + rr * BP_D + e];
        Pw[j * DD + o] = u;
      }
    __syncthreads();
  }
  double* m1 = a.ctab_m1 + (size_t)c * (BP_CGA + 1) * DN;
  for (int i = tid; i < (BP_CGA + 1) * DN; i += T) {
    const int ai = i / DN, o = i - ai * DN;
    const double ta = ai * d1;
    const int qm = mA[ai];
    double y = U[o], cj = 1.0;
    for (int j = 1; j <= qm; ++j) { cj = cj * (ta * bp_rk[j - 1]); y += cj * U[j * DN + o]; }
    m1[i] = y;
  }
  double* e2 = a.ctab_e2 + (size_t)c * BP_CGB * DD;
  for (int i = tid; i < BP_CGB * DD; i += T) {
    const int bi = i / DD, o = i - bi * DD;
    const double tb = bi * a.cdelta;
    const int qm = mB[bi];
    double y = Pw[o], cj = 1.0;
    for (int j = 1; j <= qm; ++j) { cj = cj * (tb * bp_rk[j - 1]); y += cj * Pw[j * DD + o]; }
    e2[i] = y;
  }
}

// y = sum_(j < NT) N_j (w_j z0): NT terms of a D x n matrix-vector product, N_j broadcast from shared memory
template <int NT>
__device__ __forceinline__ void bp_c_forward(const double* __restrict__ sN, const double (&w)[BP_CJ], const double (&z0)[BP_N],
                                             double (&ymu)[BP_N], double (&yS)[BP_S]) {
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    double wz[BP_N];
#pragma unroll
    for (int l = 0; l < BP_N; ++l) wz[l] = w[j] * z0[l];
    const double* nrow = sN + j * BP_CNLD;
#pragma unroll
    for (int d = 0; d < BP_D; ++d) {
      double s = (d < BP_N) ? ymu[d] : yS[d - BP_N];
#pragma unroll
      for (int l = 0; l < BP_N; ++l) s += nrow[d * BP_N + l] * wz[l];
      if (d < BP_N) ymu[d] = s; else yS[d - BP_N] = s;
    }
  }
}

extern "C" __global__ void __launch_bounds__(BP_THREADS, BP_CMINBLOCKS) bp_fusedc(BpObsArgs a) {
  bp_pdl_enter();
  const int c = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (bp_skip(a.skip, c)) return;
  constexpr int DN = BP_D * BP_N, DD = BP_D * BP_D, NW = BP_THREADS / 32;
  double a1;
#if BP_NOISE
  const double sigma = a.theta[(size_t)c * BP_DIM + BP_P];
#else
  const double sigma = 0.0;
#endif
#if BP_CX
  constexpr int WARPD = BP_XWARP;
#else
  constexpr int WARPD = BP_CWARP;
#endif
  double* Lm = bp_smem + (size_t)NW * WARPD;   // [D][D], after the warps' areas
  {
    double th[BP_DIM];
#pragma unroll
    for (int k = 0; k < BP_DIM; ++k) th[k] = a.theta[(size_t)c * BP_DIM + k];
    double A[BP_N * BP_N], H[BP_N * BP_S], r[BP_E], xk[BP_KDEP];
#pragma unroll
    for (int k = 0; k < BP_KDEP; ++k) xk[k] = 0.0;
    bp_rates(th, xk, r);
    bool fin = true;
#pragma unroll
    for (int e = 0; e < BP_E; ++e) fin = fin && isfinite(r[e]);
    bp_assemble(r, A, H);
    a1 = bp_norm1(A);
    if (!bp_w_path(fin, a1, a.nobs > 0 ? a.t[0] : 0.0)) return;   // bp_fused handles this chain
    bp_coop_build_L(A, H, Lm);
  }
#if BP_CX
  double* sN = bp_smem + (size_t)warp * WARPD;      // [BP_XR][BP_XNLD]   N_j[d][l] at [(j, l)][d], d < BP_XDM
  double* sN9 = sN + BP_XR * BP_XNLD;               // [BP_XR][BP_XDR]    the remaining components of N_j
  double* sX = sN9 + BP_XR * (BP_XDR ? BP_XDR : 1); // [BP_XR][BP_WLD]    X[(j, l)][obs] = c_j(h_obs) z0_obs,l
  double* sBm = sX + BP_XR * BP_WLD;                // [8][BP_XYLD]       Y[d][obs] out of the forward product, then b[d][obs] (row stride BP_WLD)
#else
  double* sN = bp_smem + (size_t)warp * WARPD;      // [BP_CJ][BP_CNLD]   N_j[(d, l)]
  double* sT = sN + BP_CJ * BP_CNLD;                // [8][BP_WLD]        T[j][obs] = c_j(h_obs)
  double* sB = sT + 8 * BP_WLD;                     // [BP_CHCT*8][BP_WLD] B[(d,l)][obs] = b_d z0_l, one column block at a time
#endif
  for (int i = lane; i < WARPD; i += 32) sN[i] = 0.0;   // padding rows / columns stay zero
  __syncthreads();
  const double* m1 = a.ctab_m1 + (size_t)c * (BP_CGA + 1) * DN;
  const double* e2 = a.ctab_e2 + (size_t)c * BP_CGB * DD;
  double accS[BP_WCT][2];
  double lp = 0.0, sb = 0.0;
  int status = 0, dhint = 1, dhh = 1;
  double flops = 0.0;   // algorithmic flops of the series terms and of the per-group work (bpc_count_flops adds the per-observation constant)
  const int g = lane >> 2, tg = lane & 3;
  const int od = lane / BP_N, ol = lane - od * BP_N;   // this lane's (d, l) entry of the D x n blocks (lanes < DN)
  const int g0 = blockIdx.x * a.cgpc, g1 = min(a.cgroups, g0 + a.cgpc);
  for (int grp = g0 + warp; grp < g1; grp += NW) {   // warp-uniform
    const double tau = a.cg_tau[grp];
    const int gi = a.cg_idx[grp];
    BpPlan pc, ph;
    const bool okc = bp_make_plan(a1, tau, pc, dhint, BP_WCAP, BP_WTHETA_CAP);
    const bool okh = bp_make_plan(a1, a.cg_hmax[grp], ph, dhh, BP_CJ - 1, BP_WTHETA_CAP);
    double* sg = a.csg + ((size_t)c * a.cgroups + grp) * BP_CSG;   // this group's S^g, row-major [j][8 BP_WCT]
    if (!okc || !okh || pc.s != 1 || ph.s != 1 || !((a1 + a1) * a.cg_hmax[grp] <= bp_th[ph.m - 1])) {   // (never for data bpc_data_create accepted)
      status |= BP_ST_RATE_NONFINITE;
      for (int i = lane; i < BP_CSG; i += 32) sg[i] = 0.0;
      continue;
    }
    // short series: j = 0 .. J with J = the degree d of the plan (theta_h^d / d! <= 2^-55), not d + 1 as for the long
    // series: theta_h <= bp_th[BP_CJ - 1] << 1, so each further term is smaller by theta_h / j and the coupling factor
    // (j ||H|| / 2||A||) of the block-triangular operator cannot bring term d + 1 back above the threshold
    const int J = ph.m - 1;
    // N_0 and the chain N_1 .. N_J (2 D^2 n each), and the Toeplitz product ((J + 1)(m_g + 1) - 1 coefficient x row updates of 2 D n)
    if (lane == 0) flops += 2.0 * BP_D * BP_D * BP_N * (J + 1) + 2.0 * BP_D * BP_N * ((J + 1) * (pc.m + 1) - 1);
#if BP_CX
    // ---- group prologue: N_0 = E2[b] M1[a], N_(j+1) = L N_j; stored transposed, [(j, l)][d] ----
    {
      const int ai = gi / BP_CGB, bi = gi - ai * BP_CGB;
      double nj = 0.0;
      if (lane < DN) {
#pragma unroll
        for (int rr = 0; rr < BP_D; ++rr) nj += e2[(size_t)bi * DD + od * BP_D + rr] * m1[(size_t)ai * DN + rr * BP_N + ol];
        if (od < BP_XDM) sN[ol * BP_XNLD + od] = nj; else sN9[ol * BP_XDR + (od - BP_XDM)] = nj;
      }
      __syncwarp();
      for (int j = 1; j <= J; ++j) {
        if (lane < DN) {
          const double* pn = sN + ((j - 1) * BP_N + ol) * BP_XNLD;
          nj = 0.0;
#pragma unroll
          for (int rr = 0; rr < BP_XDM; ++rr) nj += Lm[od * BP_D + rr] * pn[rr];
#pragma unroll
          for (int rr = BP_XDM; rr < BP_D; ++rr) nj += Lm[od * BP_D + rr] * sN9[((j - 1) * BP_N + ol) * BP_XDR + (rr - BP_XDM)];
          if (od < BP_XDM) sN[(j * BP_N + ol) * BP_XNLD + od] = nj; else sN9[(j * BP_N + ol) * BP_XDR + (od - BP_XDM)] = nj;
        }
        __syncwarp();
      }
    }
#pragma unroll
    for (int ct = 0; ct < BP_WCT; ++ct) { accS[ct][0] = 0.0; accS[ct][1] = 0.0; }
    // the A fragments of the forward product (N, shared by the group's passes) stay in registers
    double nf[BP_XKS];
#pragma unroll
    for (int kk = 0; kk < BP_XKS; ++kk) nf[kk] = sN[(kk * 4 + tg) * BP_XNLD + g];
#if BP_XDR
    // S update of the components beyond the tensor tile: per-lane accumulators over the group's passes, summed over lanes once per group
    double acc9[BP_XDR][BP_XR];
#pragma unroll
    for (int q = 0; q < BP_XDR; ++q)
#pragma unroll
      for (int r = 0; r < BP_XR; ++r) acc9[q][r] = 0.0;
#endif
    // ---- the group's observations, 32 per pass ----
    const int ch1 = a.cg_chunk0[grp + 1];
    for (int ch = a.cg_chunk0[grp]; ch < ch1; ++ch) {
      const int k = ch * 32 + lane;
      bool valid = k < a.nobs;
      double z0[BP_N], xo[BP_N], ymu[BP_N], yS[BP_S], bmu[BP_N], bS[BP_S], w[BP_CJ];
#pragma unroll
      for (int i = 0; i < BP_N; ++i) { z0[i] = a.z0[(size_t)i * a.npad + k]; xo[i] = a.xo[(size_t)i * a.npad + k]; bmu[i] = 0.0; }
#pragma unroll
      for (int q = 0; q < BP_S; ++q) bS[q] = 0.0;
      const double h = valid ? a.t[k] - tau : 0.0;
      // X[(j, l)] = c_j(h) z0_l (zero beyond J); the components beyond the tensor tile as a dot product on the way
      double yr[BP_XDR ? BP_XDR : 1][BP_N];   // (one partial sum per l: independent chains)
#pragma unroll
      for (int q = 0; q < (BP_XDR ? BP_XDR : 1); ++q)
#pragma unroll
        for (int l = 0; l < BP_N; ++l) yr[q][l] = 0.0;
      w[0] = 1.0;
#pragma unroll
      for (int j = 0; j < BP_CJ; ++j) {
        if (j > 0) w[j] = w[j - 1] * (h * (j <= J ? bp_rk[j - 1] : 0.0));
#pragma unroll
        for (int l = 0; l < BP_N; ++l) {
          const double x = w[j] * z0[l];
          sX[(j * BP_N + l) * BP_WLD + lane] = x;
#pragma unroll
          for (int q = 0; q < BP_XDR; ++q) yr[q][l] += sN9[(j * BP_N + l) * BP_XDR + q] * x;
        }
      }
#pragma unroll
      for (int q = 0; q < BP_XDR; ++q)
#pragma unroll
        for (int l = 1; l < BP_N; ++l) yr[q][0] += yr[q][l];
      __syncwarp();
      // forward product on the tensor path: Y[d][obs] = sum_r N[r][d] X[r][obs], d < BP_XDM, four 8-observation tiles
      {
        double accY[4][2];
#pragma unroll
        for (int ot = 0; ot < 4; ++ot) { accY[ot][0] = 0.0; accY[ot][1] = 0.0; }
#pragma unroll
        for (int kk = 0; kk < BP_XKS; ++kk)
#pragma unroll
          for (int ot = 0; ot < 4; ++ot) bp_dmma(accY[ot][0], accY[ot][1], nf[kk], sX[(kk * 4 + tg) * BP_WLD + ot * 8 + g]);
#pragma unroll
        for (int ot = 0; ot < 4; ++ot) *reinterpret_cast<double2*>(sBm + g * BP_XYLD + ot * 8 + 2 * tg) = make_double2(accY[ot][0], accY[ot][1]);
      }
      __syncwarp();
#pragma unroll
      for (int d = 0; d < BP_D; ++d) {
        const double yd = d < BP_XDM ? sBm[(d < BP_XDM ? d : 0) * BP_XYLD + lane] : yr[d < BP_XDM ? 0 : d - BP_XDM][0];
        if (d < BP_N) ymu[d] = yd; else yS[d - BP_N] = yd;
      }
      __syncwarp();
      if (valid) {
        flops += (J + 1) * (2.0 + BP_N + 4.0 * BP_D * BP_N);   // per term: c_j(h), w_j z0, N_j (w_j z0), one row of the S update
        if (!bp_score_obs(xo, sigma, ymu, yS, lp, sb, bmu, bS)) { status |= BP_ST_NOT_PD; valid = false; }
      }
      // stage the cotangent in packed coordinates (zero for an observation that does not count)
#pragma unroll
      for (int d = 0; d < BP_D; ++d) {
        double bd;
        if (d < BP_N) bd = bmu[d];
        else {
          bool diag = false;
#pragma unroll
          for (int i = 0; i < BP_N; ++i) diag = diag || (bp_sidx(i, i) == d - BP_N);
          bd = diag ? bS[d - BP_N] : bS[d - BP_N] + bS[d - BP_N];
        }
        bd = valid ? bd : 0.0;
        if (d < BP_XDM) sBm[d * BP_WLD + lane] = bd;
#if BP_XDR
        else {
#pragma unroll
          for (int j = 0; j < BP_CJ; ++j) {
            const double wb = w[j] * bd;
#pragma unroll
            for (int l = 0; l < BP_N; ++l) acc9[d < BP_XDM ? 0 : d - BP_XDM][j * BP_N + l] += wb * z0[l];
          }
        }
#endif
      }
      __syncwarp();
#pragma unroll
      for (int ob = 0; ob < 8; ++ob) {   // 4 observations per mma
        const double bf = sBm[g * BP_WLD + ob * 4 + tg];
#pragma unroll
        for (int rt = 0; rt < BP_XRT; ++rt) bp_dmma(accS[rt][0], accS[rt][1], sX[(rt * 8 + g) * BP_WLD + ob * 4 + tg], bf);
      }
      __syncwarp();
    }
    // ---- park S^g: row j, columns (l, d) for d < BP_XDM, then the remaining components (q, l) ----
#pragma unroll
    for (int rt = 0; rt < BP_XRT; ++rt) {
      const int r = rt * 8 + g, jj = r / BP_N, ll = r - jj * BP_N;
      *reinterpret_cast<double2*>(sg + jj * (BP_WCT * 8) + ll * 8 + 2 * tg) = make_double2(accS[rt][0], accS[rt][1]);
    }
#if BP_XDR
    // sum the per-lane accumulators over the warp's lanes: one product with a matrix of ones (rows (j, l) x lanes)
#pragma unroll
    for (int q = 0; q < BP_XDR; ++q) {
#pragma unroll
      for (int r = 0; r < BP_XR; ++r) sX[r * BP_WLD + lane] = acc9[q][r];
      __syncwarp();
      double red9[BP_XRT][2];
#pragma unroll
      for (int rt = 0; rt < BP_XRT; ++rt) { red9[rt][0] = 0.0; red9[rt][1] = 0.0; }
#pragma unroll
      for (int ob = 0; ob < 8; ++ob)
#pragma unroll
        for (int rt = 0; rt < BP_XRT; ++rt) bp_dmma(red9[rt][0], red9[rt][1], sX[(rt * 8 + g) * BP_WLD + ob * 4 + tg], 1.0);
      if (tg == 0) {
#pragma unroll
        for (int rt = 0; rt < BP_XRT; ++rt) {
          const int r = rt * 8 + g, jj = r / BP_N, ll = r - jj * BP_N;
          sg[jj * (BP_WCT * 8) + BP_N * 8 + q * BP_N + ll] = red9[rt][0];
        }
      }
      __syncwarp();
    }
#endif
#else
    // ---- group prologue: N_0 = E2[b] M1[a], N_(j+1) = L N_j, coefficient table c_i(tau_g) ----
    {
      const int ai = gi / BP_CGB, bi = gi - ai * BP_CGB;
      double nj = 0.0;
      if (lane < DN) {
#pragma unroll
        for (int rr = 0; rr < BP_D; ++rr) nj += e2[(size_t)bi * DD + od * BP_D + rr] * m1[(size_t)ai * DN + rr * BP_N + ol];
        sN[lane] = nj;
      }
      __syncwarp();
      for (int j = 1; j <= J; ++j) {
        if (lane < DN) {
          nj = 0.0;
#pragma unroll
          for (int rr = 0; rr < BP_D; ++rr) nj += Lm[od * BP_D + rr] * sN[(j - 1) * BP_CNLD + rr * BP_N + ol];
          sN[j * BP_CNLD + lane] = nj;
        }
        __syncwarp();
      }
    }
#pragma unroll
    for (int ct = 0; ct < BP_WCT; ++ct) { accS[ct][0] = 0.0; accS[ct][1] = 0.0; }
    __syncwarp();
    // ---- the group's observations, 32 per pass ----
    const int ch1 = a.cg_chunk0[grp + 1];
    for (int ch = a.cg_chunk0[grp]; ch < ch1; ++ch) {
      const int k = ch * 32 + lane;
      bool valid = k < a.nobs;
      double z0[BP_N], xo[BP_N], ymu[BP_N], yS[BP_S], bmu[BP_N], bS[BP_S], w[BP_CJ];
#pragma unroll
      for (int i = 0; i < BP_N; ++i) { z0[i] = a.z0[(size_t)i * a.npad + k]; xo[i] = a.xo[(size_t)i * a.npad + k]; ymu[i] = 0.0; bmu[i] = 0.0; }
#pragma unroll
      for (int q = 0; q < BP_S; ++q) { yS[q] = 0.0; bS[q] = 0.0; }
      const double h = valid ? a.t[k] - tau : 0.0;
      w[0] = 1.0;
#pragma unroll
      for (int j = 1; j < BP_CJ; ++j) w[j] = w[j - 1] * (h * (j <= J ? bp_rk[j - 1] : 0.0));   // zero beyond J
      // branch-free variants by number of terms (terms beyond J have w_j = 0 and read stale but finite rows of sN)
      if (J <= 5) bp_c_forward<6>(sN, w, z0, ymu, yS);
      else if (J == 6) bp_c_forward<7>(sN, w, z0, ymu, yS);
      else bp_c_forward<8>(sN, w, z0, ymu, yS);
      if (valid) {
        flops += (J + 1) * (2.0 + BP_N + 4.0 * BP_D * BP_N);   // per term: c_j(h), w_j z0, N_j (w_j z0), one row of the S update
        if (!bp_score_obs(xo, sigma, ymu, yS, lp, sb, bmu, bS)) { status |= BP_ST_NOT_PD; valid = false; }
      }
      // stage this observation's operands: T column (c_j(h), j = 0..J) and B column (packed-coordinate cotangent x z0)
#pragma unroll
      for (int j = 0; j < BP_CJ; ++j) sT[j * BP_WLD + lane] = valid ? w[j] : 0.0;   // a zero T column removes the observation from S
#pragma unroll
      for (int hb = 0; hb < BP_CSPLIT; ++hb) {   // column blocks of the B operand (stale padding columns of a later block are
                                                 // harmless: columns >= D n of S and W are never read)
#pragma unroll
        for (int d = 0; d < BP_D; ++d) {
          double bd;
          if (d < BP_N) bd = bmu[d];
          else {
            bool diag = false;
#pragma unroll
            for (int i = 0; i < BP_N; ++i) diag = diag || (bp_sidx(i, i) == d - BP_N);
            bd = diag ? bS[d - BP_N] : bS[d - BP_N] + bS[d - BP_N];
          }
#pragma unroll
          for (int l = 0; l < BP_N; ++l)
            if ((d * BP_N + l) / (BP_CHCT * 8) == hb) sB[((d * BP_N + l) - hb * BP_CHCT * 8) * BP_WLD + lane] = bd * z0[l];
        }
        __syncwarp();
#pragma unroll
        for (int ob = 0; ob < 8; ++ob) {   // 4 observations per mma
          const double af = sT[g * BP_WLD + ob * 4 + tg];
#pragma unroll
          for (int ct = 0; ct < BP_CHCT; ++ct)
            bp_dmma(accS[hb * BP_CHCT + ct][0], accS[hb * BP_CHCT + ct][1], af, sB[(ct * 8 + g) * BP_WLD + ob * 4 + tg]);
        }
        __syncwarp();
      }
    }
    // ---- park S^g (accumulator fragment -> row-major, 16-byte stores; L2-resident until the Toeplitz products below) ----
#pragma unroll
    for (int ct = 0; ct < BP_WCT; ++ct) *reinterpret_cast<double2*>(sg + g * (BP_WCT * 8) + ct * 8 + 2 * tg) = make_double2(accS[ct][0], accS[ct][1]);
#endif
  }
  // ---- W += C(tau_g) S^g for this warp's groups: rows n - 1 (n = 1 .. m_g + J), k = j (8), columns (d, l).  Deferred to here so
  // that the 2 BP_CRT BP_WCT accumulator registers do not live through the observation loop. ----
  __syncwarp();
  double accW[BP_CRT * BP_WCT][2];
#pragma unroll
  for (int i = 0; i < BP_CRT * BP_WCT; ++i) { accW[i][0] = 0.0; accW[i][1] = 0.0; }
  dhint = 1; dhh = 1;
  for (int grp = g0 + warp; grp < g1; grp += NW) {
    BpPlan pc, ph;
    const bool okc = bp_make_plan(a1, a.cg_tau[grp], pc, dhint, BP_WCAP, BP_WTHETA_CAP);
    const bool okh = bp_make_plan(a1, a.cg_hmax[grp], ph, dhh, BP_CJ - 1, BP_WTHETA_CAP);
    if (!okc || !okh || pc.s != 1 || ph.s != 1 || !((a1 + a1) * a.cg_hmax[grp] <= bp_th[ph.m - 1])) continue;   // flagged above
    const int mg = pc.m, J = ph.m - 1;
    const double* sg = a.csg + ((size_t)c * a.cgroups + grp) * BP_CSG;
    const double* cf = a.cg_coef + (size_t)grp * BP_CCTAB + 8;
    double bf[2][BP_WCT];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int ct = 0; ct < BP_WCT; ++ct) bf[ks][ct] = __ldcg(sg + (ks * 4 + tg) * (BP_WCT * 8) + ct * 8 + g);
#pragma unroll
    for (int rt = 0; rt < BP_CRT; ++rt) {
      if (rt * 8 < mg + J) {
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const int ix = (rt * 8 + g + 1) - (ks * 4 + tg);   // c_(n-j)(tau_g), n = 8 rt + g + 1, j = 4 ks + tg
          const double af = (ix <= mg) ? cf[ix] : 0.0;
#pragma unroll
          for (int ct = 0; ct < BP_WCT; ++ct) bp_dmma(accW[rt * BP_WCT + ct][0], accW[rt * BP_WCT + ct][1], af, bf[ks][ct]);
        }
      }
    }
  }
  // CTA partial of W: sum the warps' accumulators in warp order (fragment layout kept; bp_wpost knows it)
  __syncthreads();
  double* red = bp_smem;   // [warps][BP_CFRAG]
#pragma unroll
  for (int i = 0; i < BP_CRT * BP_WCT; ++i) { red[(size_t)warp * BP_CFRAG + i * 64 + lane * 2] = accW[i][0]; red[(size_t)warp * BP_CFRAG + i * 64 + lane * 2 + 1] = accW[i][1]; }
  __syncthreads();
  double* wout = a.wpart + ((size_t)c * a.ntiles + blockIdx.x) * BP_CFRAG;
  for (int i = tid; i < BP_CFRAG; i += BP_THREADS) {
    double sacc = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) sacc += red[(size_t)w * BP_CFRAG + i];
    wout[i] = sacc;
  }
  __syncthreads();
  double vals[BP_NPART];
  vals[0] = lp;
#pragma unroll
  for (int k = 0; k < BP_P; ++k) vals[1 + k] = 0.0;
  vals[BP_P + 1] = sb;
  vals[BP_P + 2] = flops;
  bp_block_reduce_store(vals, bp_smem, a.part + ((size_t)c * a.ntiles + blockIdx.x) * BP_NPART);
  if (status) atomicOr(a.status + c, status);
}

// ---- octet path: bp_otable + bp_fusedo + bp_toep ---------------------------------------------------------------
// X[(j, l)][k] = c_j(h_k) z0_kl does not depend on the chain, and 8 chains x D components fill the 8-wide tensor tiles exactly
// (for every n).  So a CTA takes one pass of 32 observations for EIGHT chains at once:
//     forward   Y[(cc, d)][k]  = sum_(j,l) N^cc_j[d][l] X[(j,l)][k]          D row tiles x 4 observation tiles x BP_XKS k-steps
//     score     per (chain, observation): 256 per pass, two per thread
//     S update  S'[(j,l)][(cc, d)] += sum_k X[(j,l)][k] b^cc_k[d]              BP_XRT x D tiles x 8 k-steps
// with X staged once per pass, its fragments shared by all column tiles, no padded tile rows and no ninth-component side path.
// The short series is never masked in the hot loop: N rows beyond a chain's own degree J are zero, and the rows of
// S' beyond it are dropped by bp_toep.  bp_otable finds the series lengths (J, m) and N_0(tau_g) of every (chain, group); bp_fusedo
// builds N_j = L N_(j-1) of its octet in the group prologue, bp_toep is the Toeplitz product W_n = sum_g sum_j c_(n-j)(tau_g) S^g_j for all groups at once, bp_wpost finishes.
#define BP_OU (8 * BP_D)                         // columns of an octet: (cc, d)
#define BP_ONLD (BP_OU + 4)                      // row stride of the staged N (conflict-free A fragments: stride = 12 mod 16)
// S' of one (octet, group) in global memory: element (row (j, l), column 8 ct + c8) at ((ct n + l) 8 + j) 8 + c8, i.e. the 8 x 8 block
// (all j, the 8 columns of tile ct) that one CTA of bp_toep reads is 512 contiguous bytes (row-major [(j, l)][column] made it eight
// 64-byte pieces: bp_toep ran at 1.7 TB/s on 103 MB, bound by that access pattern).  The stores of the passes' kernels are 64-byte
// pieces inside the group's 8 n x 8 D block either way.
__device__ __forceinline__ int bp_osg_off(int row, int col) {
  const int j = row / BP_N, l = row - j * BP_N;
  return ((((col >> 3) * BP_N + l) * BP_CJ + j) << 3) + (col & 7);
}
#define BP_OTHREADS 128                          // one CTA = four warps on the same pass
#ifndef BP_OMINBLOCKS
#define BP_OMINBLOCKS 4
#endif
#define BP_OYLD 40                               // row stride of Y / b; element (r, k) lives at column k ^ ((r & 2) << 1): the 16-byte
                                                 // accumulator stores, the per-lane column reads and the B fragments are all conflict-free
#define BP_OY(r, k) ((r) * BP_OYLD + ((k) ^ (((r) & 2) << 1)))
#define BP_OOBS ((2 * BP_N + 1) * 32)             // one chunk of the observation table: z0[n], xo[n], t, 32 doubles each
#define BP_OSMEM ((BP_XR * BP_ONLD + 2 * BP_XR * BP_WLD + BP_OU * BP_OYLD + 2 * BP_OOBS) * 8)

// 16-byte asynchronous copy global -> shared (LDGSTS), used to stage the next chunk of the observation table
__device__ __forceinline__ void bp_cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void bp_cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bp_cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// per chain, everything the octet path needs before its passes: the bound transform (when the call gives u: no bp_prologue launch),
// the path decision (chains that need sub-steps go to bp_fused's list), L, the series lengths (J, m) and algorithmic flops of every
// (chain, group), and N_0(tau_g) = exp(tau_g L) [e_1 .. e_n] = sum_(j <= m_g) c_j(tau_g) U_j from the Krylov blocks U_j = L^j [e_1 .. e_n]
// (c_j(tau_g): the coefficient table bp_toep uses).  One CTA per chain.
#define BP_OT_GROUPS 128                         // groups per batch of the N_0 loop (one thread each)
extern "C" __global__ void __launch_bounds__(256) bp_otable(BpObsArgs a) {
  bp_pdl_enter();
  const int c = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
  constexpr int DN = BP_D * BP_N, DD = BP_D * BP_D;
  __shared__ double Lm[DD];
  __shared__ double U[(BP_WROWS + 1) * DN];
  __shared__ double n0s[BP_OT_GROUPS * DN];    // N_0 rows of a batch of groups on their way to global memory
  double th[BP_DIM];
  if (a.u) {
    bp_transform(a.u + (size_t)c * BP_DIM, th);
    if (tid == 0) {
#pragma unroll
      for (int k = 0; k < BP_DIM; ++k) a.theta[(size_t)c * BP_DIM + k] = th[k];
      a.status[c] = 0;
    }
    __syncthreads();   // the status is reset before anybody ORs into it
  } else {
#pragma unroll
    for (int k = 0; k < BP_DIM; ++k) th[k] = a.theta[(size_t)c * BP_DIM + k];
  }
  double A[BP_N * BP_N], H[BP_N * BP_S], r[BP_E], xk[BP_KDEP];
#pragma unroll
  for (int k = 0; k < BP_KDEP; ++k) xk[k] = 0.0;
  bp_rates(th, xk, r);
  bool fin = true;
#pragma unroll
  for (int e = 0; e < BP_E; ++e) fin = fin && isfinite(r[e]);
  bp_assemble(r, A, H);
  const double a1 = bp_norm1(A);
  const double tmax = a.nobs > 0 ? a.t[0] : 0.0;
  int* jm = a.cg_jm + (size_t)c * a.cgroups;
  double* fl = a.cg_flops + (size_t)c * a.cgroups;
  if (bp_skip(a.skip, c)) {          // nobody takes this chain
    if (tid == 0) a.ctab_l[(size_t)c * (DD + 1) + DD] = -1.0;
    for (int g = tid; g < a.cgroups; g += T) { jm[g] = -1; fl[g] = 0.0; }
    return;
  }
  if (!bp_w_path(fin, a1, tmax)) {   // bp_fused takes this chain
    if (tid == 0) {
      a.fb[1 + atomicAdd(a.fb, 1)] = c;
      a.ctab_l[(size_t)c * (DD + 1) + DD] = -1.0;
    }
    for (int g = tid; g < a.cgroups; g += T) { jm[g] = -1; fl[g] = 0.0; }
    return;
  }
  BpPlan p;
  int dh = 1;
  bp_make_plan(a1, tmax, p, dh, BP_WCAP, BP_WTHETA_CAP);
  const int mtop = p.m;
  bp_coop_build_L(A, H, Lm);
  for (int o = tid; o < DN; o += T) { const int d = o / BP_N, l = o - d * BP_N; U[o] = (d == l) ? 1.0 : 0.0; }
  // series lengths and flops of the groups
  for (int g = tid; g < a.cgroups; g += T) {
    BpPlan pc, ph;
    int d1 = 1, d2 = 1, J = -1, mg = 0;
    const bool okc = bp_make_plan(a1, a.cg_tau[g], pc, d1, BP_WCAP, BP_WTHETA_CAP);
    const bool okh = bp_make_plan(a1, a.cg_hmax[g], ph, d2, BP_CJ - 1, BP_WTHETA_CAP);
    if (!okc || !okh || pc.s != 1 || ph.s != 1 || !((a1 + a1) * a.cg_hmax[g] <= bp_th[ph.m - 1])) {   // (never for data bpc_data_create accepted)
      atomicOr(a.status + c, BP_ST_RATE_NONFINITE);
    } else {
      J = ph.m - 1;
      mg = pc.m;
    }
    jm[g] = J < 0 ? -1 : (J | (mg << 8));
    double f = 0.0;
    if (J >= 0) {
      const int k0 = a.cg_chunk0[g] * 32, k1 = min(a.nobs, a.cg_chunk0[g + 1] * 32);
      f = 2.0 * BP_D * BP_D * BP_N * (J + 1) + 2.0 * BP_D * BP_N * ((J + 1) * (mg + 1) - 1)      // N_0 .. N_J and the Toeplitz product (as bp_fusedc counts them)
          + (double)max(0, k1 - k0) * (J + 1) * (2.0 + BP_N + 4.0 * BP_D * BP_N);               // per observation and term: c_j(h), w_j z0, N_j (w_j z0), one row of the S update
    }
    fl[g] = f;
  }
  __syncthreads();   // L, U_0 and the groups' series lengths (global, this CTA's own) are written
  for (int o = tid; o < DD; o += T) a.ctab_l[(size_t)c * (DD + 1) + o] = Lm[o];
  if (tid == 0) a.ctab_l[(size_t)c * (DD + 1) + DD] = a1;
  for (int j = 1; j <= mtop; ++j) {
    for (int o = tid; o < DN; o += T) {
      const int d = o / BP_N, l = o - d * BP_N;
      double u = 0.0;
#pragma unroll
      for (int rr = 0; rr < BP_D; ++rr) u += Lm[d * BP_D + rr] * U[(j - 1) * DN + rr * BP_N + l];
      U[j * DN + o] = u;
    }
    __syncthreads();
  }
  // N_0 of the groups: thread = one group with its D n sums in registers, so that a term costs ONE coefficient load and D n broadcast
  // reads of U (one (group, entry) per thread made this loop LSU-bound: two memory instructions per FMA, 22 of the kernel's 38 us);
  // the rows go through shared memory to be stored coalesced.  Same operations in the same order for every entry.
  double* n0 = a.n0tab + (size_t)c * a.cgroups * DN;
  for (int gb = 0; gb < a.cgroups; gb += BP_OT_GROUPS) {
    const int g = gb + tid;
    if (tid < BP_OT_GROUPS && g < a.cgroups) {
      const int q = jm[g];
      double y[DN];
#pragma unroll
      for (int o = 0; o < DN; ++o) y[o] = 0.0;
      if (q >= 0) {
        const int mg = q >> 8;
        const double* cf = a.cg_coef + (size_t)g * BP_CCTAB + 8;
#pragma unroll
        for (int o = 0; o < DN; ++o) y[o] = U[o];
        for (int j = 1; j <= mg; ++j) {
          const double cj = cf[j];
#pragma unroll
          for (int o = 0; o < DN; ++o) y[o] += cj * U[j * DN + o];
        }
      }
#pragma unroll
      for (int o = 0; o < DN; ++o) n0s[tid * DN + o] = y[o];
    }
    __syncthreads();
    const int cnt = min(BP_OT_GROUPS, a.cgroups - gb) * DN;
    for (int i = tid; i < cnt; i += T) n0[(size_t)gb * DN + i] = n0s[i];
    __syncthreads();
  }
}

// One CTA (four warps) per pass of 32 observations x 8 chains; every phase is split four ways, three barriers per pass:
//   stage    warp w writes the rows of X of terms j = 2w, 2w + 1 (X is double-buffered: staging pass p + 1 overlaps the S update of p);
//            the chunk of the observation table of pass p + 1 arrives by cp.async while pass p runs
//   forward  warp w = (wlo, whi): observation tiles 2 wlo, 2 wlo + 1 x the (cc, d) tiles BP_OHT whi .. BP_OHT whi + BP_OHT - 1, plus
//            (BP_D odd) the last (cc, d) tile x observation tile 2 wlo + whi
//   score    thread (w, lane): observation `lane` under chains 2w and 2w + 1; the cotangent replaces Y in place
//   S update warp w: all 32 observations x the (cc, d) tiles BP_OTW w .. BP_OTW w + BP_OTW - 1 x all row tiles, plus one
//            (row tile, remaining (cc, d) tile) pair: every tile of S' belongs to one warp (no reduction between warps)
// so that every warp carries the same number of tensor products without a predicate in the loops; all shared-memory addresses are a
// per-lane base plus a compile-time offset.  The FP64 datapath (DMMA and DFMA share it) is kept busy by the other CTAs of the SM
// while one waits at a barrier.
#define BP_OHT (BP_D / 2)                         // (cc, d) tiles of a warp pair
#define BP_OEX (BP_D & 1)                         // one more tile shared by the four warps
#define BP_OTW (BP_D / 4)                         // S update: (cc, d) tiles owned by each warp with all their row tiles
#define BP_OREM (BP_D - 4 * BP_OTW)               // remaining (cc, d) tiles: their BP_OREM * BP_XRT (<= 4) tiles go to warps 0, 1, ..
extern "C" __global__ void __launch_bounds__(BP_OTHREADS, BP_OMINBLOCKS) bp_fusedo(BpObsArgs a) {
  bp_pdl_enter();
  const int o = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, tg = lane & 3;
  double* sN = bp_smem;                         // [BP_XR][BP_ONLD]     N_j[d][l] of chain cc at [(j, l)][cc * BP_D + d]
  double* sXb = sN + BP_XR * BP_ONLD;           // [2][BP_XR][BP_WLD]   X[(j, l)][obs]
  double* sY = sXb + 2 * BP_XR * BP_WLD;        // [BP_OU][BP_OYLD]     Y, then b: [(cc, d)][obs] (swizzled, BP_OY)
  double* sOb = sY + BP_OU * BP_OYLD;           // [2][2 BP_N + 1][32]  the pass's chunk of the observation table (z0, xo, t), double-buffered
  __shared__ int sjm[8];
  const int wlo = warp & 1, whi = warp >> 1;
  const int swz = (g >> 1) & 1;                 // rows r of Y with bit 1 set keep column k at k ^ 4 (r & 2 == g & 2 for tile rows 8 t + g)
  // per-lane bases (doubles)
  const int fN = tg * BP_ONLD + whi * (BP_OHT * 8) + g;            // A fragments of the forward product: + kk * 4 * BP_ONLD + t * 8
  const int fNx = tg * BP_ONLD + 2 * BP_OHT * 8 + g;               // ... of the last tile
  const int fX = tg * BP_WLD + wlo * 16 + g;                       // B fragments: + kk * 4 * BP_WLD + q * 8
  const int fXx = tg * BP_WLD + (2 * wlo + whi) * 8 + g;           // ... for the last tile's observation tile
  const int fY = (whi * (BP_OHT * 8) + g) * BP_OYLD + wlo * 16 + ((2 * tg) ^ (swz << 2));   // accumulator stores: + t * 8 * BP_OYLD + q * 8
  const int fYx = (2 * BP_OHT * 8 + g) * BP_OYLD + (2 * wlo + whi) * 8 + ((2 * tg) ^ (swz << 2));
  const int uX = g * BP_WLD + tg;                                  // A fragments of the S update: + rt * 8 * BP_WLD + ob * 4
  const int uY = (warp * (BP_OTW * 8) + g) * BP_OYLD + tg;         // B fragments: + ct * 8 * BP_OYLD + (ob ^ swz) * 4
  const int wx = min(warp, BP_OREM * BP_XRT - 1);                  // this warp's remaining tile (warps beyond the last one repeat it and drop the result)
  const int xct = 4 * BP_OTW + wx / BP_XRT, xrt = wx % BP_XRT;
  const int uXx = (xrt * 8 + g) * BP_WLD + tg;
  const int uYx = (xct * 8 + g) * BP_OYLD + tg;
  // score: row r = (2 warp + i2) BP_D + d; r & 2 = ((9 i2 + d) & 2) ^ (2 wlo) for BP_D = 9 (18 warp = 2 wlo mod 4), computed generally below
  double qs[2] = {0.0, 0.0}, sb[2] = {0.0, 0.0};   // sums of the quadratic forms; the determinants are multiplied up (one log per thread at the end)
  BpDetProd dp[2];
  dp[0].init(); dp[1].init();
  int notpd = 0;                                // bit i: an observation of chain 2 warp + i had a covariance that is not positive definite
  int pass = 0;
  const int g0 = blockIdx.x * a.cgpc, g1 = min(a.cgroups, g0 + a.cgpc);
  {   // an octet none of whose chains is on this path (all skipped by the sampler, or all on the sub-step fallback): nothing to do.
      // (jm < 0 holds for every group of such a chain, so the first group decides; bp_toep masks S' by jm and reads nothing from it)
    bool any = false;
    if (g0 < g1)
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) any = any || (o * 8 + cc < a.nchains && a.cg_jm[(size_t)(o * 8 + cc) * a.cgroups + g0] >= 0);
    if (!any) return;
  }
  // the chunks of this CTA's groups are consecutive: chunk `ch` of pass p is copied into sOb[p & 1] while pass p - 1 runs
  // (thread = one 16-byte piece of one array; complete before the barrier that follows the score of pass p - 1)
  const int chend = g0 < g1 ? a.cg_chunk0[g1] : 0;
  const double* osrc = nullptr;
  {
    const int ai = tid >> 4, q = tid & 15;
    if (ai < 2 * BP_N + 1) osrc = (ai < BP_N ? a.z0 + (size_t)ai * a.npad : (ai < 2 * BP_N ? a.xo + (size_t)(ai - BP_N) * a.npad : a.t)) + 2 * q;
    if (osrc && g0 < g1) bp_cp_async16(sOb + tid * 2, osrc + (size_t)a.cg_chunk0[g0] * 32);
    bp_cp_async_commit();
    bp_cp_async_wait_all();
    __syncthreads();
  }
  for (int grp = g0; grp < g1; ++grp) {
    // ---- N_j(tau_g) of the octet's chains: N_0 from bp_otable, N_(j+1) = L N_j; thread = entries (cc, d, l).
    // Rows beyond a chain's own degree J are zero.  (The previous group's last forward product is behind a barrier.) ----
    if (tid < 8) sjm[tid] = (o * 8 + tid < a.nchains) ? a.cg_jm[(size_t)(o * 8 + tid) * a.cgroups + grp] : -1;
    __syncthreads();
    {
      constexpr int DN = BP_D * BP_N, DD = BP_D * BP_D, NQ = (8 * DN + BP_OTHREADS - 1) / BP_OTHREADS;
      double Lr[NQ][BP_D];
      int Jq[NQ], eo[NQ], dq[NQ];   // degree of the entry's chain (-1: not on this path, -2: no entry), offset of (row (0, l), column (cc, 0)) in sN, d
      int Jmax = -1;
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) Jmax = max(Jmax, sjm[cc] < 0 ? -1 : (sjm[cc] & 255));
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int e = tid + q * BP_OTHREADS;
        const int cc = e / DN, rem = e - cc * DN, d = rem / BP_N, l = rem - d * BP_N;
        Jq[q] = e < 8 * DN ? (sjm[cc & 7] < 0 ? -1 : (sjm[cc & 7] & 255)) : -2;
        eo[q] = l * BP_ONLD + cc * BP_D;
        dq[q] = d;
        double n0 = 0.0;
        if (Jq[q] >= 0) {
          const size_t c = (size_t)o * 8 + cc;
          const double* lrow = a.ctab_l + c * (DD + 1) + d * BP_D;
          n0 = a.n0tab[(c * a.cgroups + grp) * DN + rem];
#pragma unroll
          for (int rr = 0; rr < BP_D; ++rr) Lr[q][rr] = lrow[rr];
        } else {
#pragma unroll
          for (int rr = 0; rr < BP_D; ++rr) Lr[q][rr] = 0.0;
        }
        if (Jq[q] >= -1) sN[eo[q] + d] = n0;
      }
      for (int j = 1; j < BP_CJ; ++j) {
        if (j <= Jmax) __syncthreads();   // row j - 1 complete (uniform: Jmax comes from shared memory)
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          if (Jq[q] >= -1) {
            double v = 0.0;
            if (j <= Jq[q]) {
              const double* prev = sN + (j - 1) * BP_N * BP_ONLD + eo[q];
#pragma unroll
              for (int rr = 0; rr < BP_D; ++rr) v += Lr[q][rr] * prev[rr];
            }
            sN[j * BP_N * BP_ONLD + eo[q] + dq[q]] = v;
          }
        }
      }
    }
    const double tau = a.cg_tau[grp];
    double accS[BP_XRT][BP_OTW ? BP_OTW : 1][2], accX[2] = {0.0, 0.0};
#pragma unroll
    for (int rt = 0; rt < BP_XRT; ++rt)
#pragma unroll
      for (int ct = 0; ct < (BP_OTW ? BP_OTW : 1); ++ct) { accS[rt][ct][0] = 0.0; accS[rt][ct][1] = 0.0; }
    const int ch1 = a.cg_chunk0[grp + 1];
    for (int ch = a.cg_chunk0[grp]; ch < ch1; ++ch, ++pass) {
      double* sX = sXb + (pass & 1) * (BP_XR * BP_WLD);
      const double* sO = sOb + (pass & 1) * BP_OOBS;
      const bool valid = ch * 32 + lane < a.nobs;
      if (osrc && ch + 1 < chend) bp_cp_async16(sOb + ((pass + 1) & 1) * BP_OOBS + tid * 2, osrc + (size_t)(ch + 1) * 32);
      bp_cp_async_commit();
      {   // stage: terms j = 2 warp, 2 warp + 1 of this lane's observation
        double z0[BP_N];
#pragma unroll
        for (int i = 0; i < BP_N; ++i) z0[i] = sO[i * 32 + lane];
        const double h = valid ? sO[2 * BP_N * 32 + lane] - tau : 0.0;
        double we = 1.0;
        for (int j = 1; j <= 2 * warp; ++j) we = we * (h * bp_rk[j - 1]);   // warp-uniform trip count
        const double wo = we * (h * bp_rk[2 * warp]);
        double* xs = sX + (2 * warp * BP_N) * BP_WLD + lane;
#pragma unroll
        for (int l = 0; l < BP_N; ++l) {
          xs[l * BP_WLD] = we * z0[l];
          xs[(BP_N + l) * BP_WLD] = wo * z0[l];
        }
      }
      __syncthreads();   // X (and, for the first pass, N) staged; every warp is done with the previous pass
      {   // forward: Y[(cc, d)][obs] = sum_(j,l) N[(j,l)][(cc,d)] X[(j,l)][obs]
        double accY[BP_OHT][2][2];
        double accE[2] = {0.0, 0.0};
#pragma unroll
        for (int t = 0; t < BP_OHT; ++t)
#pragma unroll
          for (int q = 0; q < 2; ++q) { accY[t][q][0] = 0.0; accY[t][q][1] = 0.0; }
        const double* pn = sN + fN;
        const double* px = sX + fX;
#pragma unroll
        for (int kk = 0; kk < BP_XKS; ++kk) {
          const double xb0 = px[kk * 4 * BP_WLD], xb1 = px[kk * 4 * BP_WLD + 8];
#pragma unroll
          for (int t = 0; t < BP_OHT; ++t) {
            const double af = pn[kk * 4 * BP_ONLD + t * 8];
            bp_dmma(accY[t][0][0], accY[t][0][1], af, xb0);
            bp_dmma(accY[t][1][0], accY[t][1][1], af, xb1);
          }
#if BP_OEX
          bp_dmma(accE[0], accE[1], sN[fNx + kk * 4 * BP_ONLD], sX[fXx + kk * 4 * BP_WLD]);
#endif
        }
        double* py = sY + fY;
#pragma unroll
        for (int t = 0; t < BP_OHT; ++t)
#pragma unroll
          for (int q = 0; q < 2; ++q) *reinterpret_cast<double2*>(py + t * 8 * BP_OYLD + q * 8) = make_double2(accY[t][q][0], accY[t][q][1]);
#if BP_OEX
        *reinterpret_cast<double2*>(sY + fYx) = make_double2(accE[0], accE[1]);
#endif
      }
      __syncthreads();
      // score: this lane's observation under chains 2 warp and 2 warp + 1 (independent work); the cotangent in packed coordinates replaces Y
      {
        double xo[BP_N];
#pragma unroll
        for (int i = 0; i < BP_N; ++i) xo[i] = sO[(BP_N + i) * 32 + lane];
        const int rbase = 2 * warp * BP_D;
        const int kA = lane ^ ((rbase & 2) << 1), kB = kA ^ 4;   // column of this lane in rows r with ((r - rbase) & 2) == 0 / != 0
        double* yrow = sY + rbase * BP_OYLD;
#pragma unroll
        for (int i2 = 0; i2 < 2; ++i2) {
          double ymu[BP_N], yS[BP_S], bmu[BP_N], bS[BP_S];
          bool ok = valid && sjm[2 * warp + i2] >= 0;
#pragma unroll
          for (int i = 0; i < BP_N; ++i) { ymu[i] = yrow[(i2 * BP_D + i) * BP_OYLD + (((i2 * BP_D + i) & 2) ? kB : kA)]; bmu[i] = 0.0; }
#pragma unroll
          for (int q = 0; q < BP_S; ++q) { yS[q] = yrow[(i2 * BP_D + BP_N + q) * BP_OYLD + (((i2 * BP_D + BP_N + q) & 2) ? kB : kA)]; bS[q] = 0.0; }
          if (ok) {
#if BP_NOISE
            const double sigma = a.theta[(size_t)(o * 8 + 2 * warp + i2) * BP_DIM + BP_P];
#else
            const double sigma = 0.0;
#endif
            if (!bp_score_obs_small<true>(xo, sigma, ymu, yS, qs[i2], dp[i2], sb[i2], bmu, bS)) { notpd |= 1 << i2; ok = false; }
          }
#pragma unroll
          for (int i = 0; i < BP_N; ++i) yrow[(i2 * BP_D + i) * BP_OYLD + (((i2 * BP_D + i) & 2) ? kB : kA)] = ok ? bmu[i] : 0.0;
#pragma unroll
          for (int q = 0; q < BP_S; ++q)     // (bS is already in packed coordinates: bp_mvn_score_small<true>)
            yrow[(i2 * BP_D + BP_N + q) * BP_OYLD + (((i2 * BP_D + BP_N + q) & 2) ? kB : kA)] = ok ? bS[q] : 0.0;
        }
      }
      bp_cp_async_wait_all();   // the next pass's chunk has landed (this thread's piece; the barrier publishes all of them)
      __syncthreads();
      {   // S' update over the 32 observations: X fragments (A) shared by the column tiles, b fragments (B) by the row tiles
        const double* px = sX + uX;
        const double* pbE = sY + uY + (swz << 2);   // k-steps ob even: column block ob ^ swz = ob + swz
        const double* pbO = sY + uY - (swz << 2);   // k-steps ob odd:  ob - swz
#pragma unroll
        for (int ob = 0; ob < 8; ++ob) {
#if BP_OTW
          double xa[BP_XRT];
#pragma unroll
          for (int rt = 0; rt < BP_XRT; ++rt) xa[rt] = px[rt * 8 * BP_WLD + ob * 4];
#pragma unroll
          for (int ct = 0; ct < BP_OTW; ++ct) {
            const double bf = ((ob & 1) ? pbO : pbE)[ct * 8 * BP_OYLD + ob * 4];
#pragma unroll
            for (int rt = 0; rt < BP_XRT; ++rt) bp_dmma(accS[rt][ct][0], accS[rt][ct][1], xa[rt], bf);
          }
#endif
          bp_dmma(accX[0], accX[1], sX[uXx + ob * 4], sY[uYx + ((ob & 1) ? -(swz << 2) : (swz << 2)) + ob * 4]);
        }
      }
    }
    // ---- the group's S' in the block layout of bp_osg_off: every tile from the warp that owns it ----
    // (a layout with the GROUP inside, [tile][l][group][j][8], scatters these stores over the octet's whole region and costs more here
    // than it saves in bp_toep; re-ordering inside the group's own block is free here)
    double* sg = a.osg + ((size_t)o * a.cgroups + grp) * BP_XR * BP_OU;
#if BP_OTW
#pragma unroll
    for (int rt = 0; rt < BP_XRT; ++rt)
#pragma unroll
      for (int ct = 0; ct < BP_OTW; ++ct)
        __stcg(reinterpret_cast<double2*>(sg + bp_osg_off(rt * 8 + g, (warp * BP_OTW + ct) * 8 + 2 * tg)), make_double2(accS[rt][ct][0], accS[rt][ct][1]));
#endif
    if (warp < BP_OREM * BP_XRT) __stcg(reinterpret_cast<double2*>(sg + bp_osg_off(xrt * 8 + g, xct * 8 + 2 * tg)), make_double2(accX[0], accX[1]));
    // (the next group's prologue barrier separates this group's last S update from the next writes of N and Y)
  }
  // ---- per chain: lp (+ d/dsigma_obs, flops of this CTA's groups) into the tile partial; chain 2 warp + i belongs to this warp alone ----
#pragma unroll
  for (int i2 = 0; i2 < 2; ++i2) {
    const double l = bp_warp_sum(-0.5 * (qs[i2] + dp[i2].logdet())), s2 = bp_warp_sum(sb[i2]);
    const int c = o * 8 + 2 * warp + i2;
    if (lane == 0 && c < a.nchains) {
      double fl = 0.0;
      bool mine = false;
      for (int grp = g0; grp < g1; ++grp) { fl += a.cg_flops[(size_t)c * a.cgroups + grp]; mine = mine || a.cg_jm[(size_t)c * a.cgroups + grp] >= 0; }
      if (mine || g0 >= g1 || a.cg_jm[(size_t)c * a.cgroups] >= 0) {   // chains on the fallback path keep the partials bp_fused writes
        double* p = a.part + ((size_t)c * a.ntiles + blockIdx.x) * BP_NPART;
        p[0] = l;
#pragma unroll
        for (int k = 0; k < BP_P; ++k) p[1 + k] = 0.0;
        p[BP_P + 1] = s2;
        p[BP_P + 2] = fl;
      }
    }
  }
  notpd = __reduce_or_sync(0xffffffffu, notpd);
  if (lane < 2 && ((notpd >> lane) & 1) && o * 8 + 2 * warp + lane < a.nchains) atomicOr(a.status + o * 8 + 2 * warp + lane, BP_ST_NOT_PD);
}

#if BP_ALT_KERNELS   // the measured alternatives of bp_fusedo (BPC_O2 / BPC_O9): compiled only when asked for (codegen.cpp)
// ---- bp_fusedo2: the octet passes with WARP-SPECIALISED roles -------------------------------------------------------------
// ncu on bp_fusedo (profiles/r02_ncu_bp_fusedo_cfg5U.txt): the shared FP64 datapath is 78 % busy; the warps spend 36 % of their time in the
// score phase, which holds 20 % of the datapath work -- its dependent chains queue behind the 16-cycle tensor products of the other
// CTAs' warps, and the three barriers of a pass tie the four schedulers of the SM together.  Here a CTA has EIGHT warps with two roles:
//   tensor warps 0 .. 3   stage X of pass p, forward product of pass p (into Y buffer p mod 3), then the S' update of pass p - 2
//   score warps 4 .. 7    score of pass p (Y -> cotangent b in place)
// Y is triple-buffered and the S' update lags TWO passes behind the forward product, so a score may take up to two tensor pass-times
// before a tensor warp waits for it; X has four buffers (X of pass p is read again by the S' update of pass p during pass p + 2, and
// written for pass p + 4 ... p + 1 before the tensor warps' barrier of that pass).  The hand-offs are named barriers (bar.arrive by the
// producer, bar.sync by the consumer; ids below).  Observation data come straight from global memory into registers, one pass ahead.
// Same arithmetic, same tile split and shared-memory layouts as bp_fusedo, same results bit for bit.
// MEASURED (cfg5U, 512 chains x 100 000 observations; tools/exp/dmma_issue_probe.cu shows that ONE warp per scheduler with two
// accumulator chains already saturates DMMA issue, so four tensor warps per CTA are enough): see DESIGN.md §4.
#define BP_O2THREADS 256
#define BP_O2SMEM ((BP_XR * BP_ONLD + 4 * BP_XR * BP_WLD + 3 * BP_OU * BP_OYLD) * 8)
#define BP_B_FULLY 1      // + Y buffer (3): Y of a pass is complete         tensor -> score
#define BP_B_FULLB 4      // + Y buffer (3): the cotangent replaced Y        score -> tensor
#define BP_B_TENSOR 7     // the four tensor warps among themselves
// (barrier ids as immediates: with a run-time id ptxas reserves all 16 barriers for the CTA)
template <int ID, int N> __device__ __forceinline__ void bp_bar_sync_i() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(N) : "memory"); }
template <int ID, int N> __device__ __forceinline__ void bp_bar_arrive_i() { asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(N) : "memory"); }
#define bp_bar_sync(id, n) bp_bar_sync_i<id, n>()
#define bp_bar_sync3(id, k) do { if ((k) == 0) bp_bar_sync_i<(id), 256>(); else if ((k) == 1) bp_bar_sync_i<(id) + 1, 256>(); else bp_bar_sync_i<(id) + 2, 256>(); } while (0)
#define bp_bar_arrive3(id, k) do { if ((k) == 0) bp_bar_arrive_i<(id), 256>(); else if ((k) == 1) bp_bar_arrive_i<(id) + 1, 256>(); else bp_bar_arrive_i<(id) + 2, 256>(); } while (0)

extern "C" __global__ void __launch_bounds__(BP_O2THREADS, 2) bp_fusedo2(BpObsArgs a) {
  bp_pdl_enter();
  const int o = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp8 = tid >> 5;
  const int warp = warp8 & 3, rtid = tid & 127;       // index within the role
  const bool scorer = warp8 >= 4;
  const int g = lane >> 2, tg = lane & 3;
  double* sN = bp_smem;                               // [BP_XR][BP_ONLD]
  double* sXb = sN + BP_XR * BP_ONLD;                 // [4][BP_XR][BP_WLD]    X of pass p in buffer p mod 4
  double* sYb = sXb + 4 * BP_XR * BP_WLD;             // [3][BP_OU][BP_OYLD]   Y, then b, of pass p in buffer p mod 3
  __shared__ int sjm[8];
  const int g0 = blockIdx.x * a.cgpc, g1 = min(a.cgroups, g0 + a.cgpc);
  {
    bool any = false;
    if (g0 < g1)
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) any = any || (o * 8 + cc < a.nchains && a.cg_jm[(size_t)(o * 8 + cc) * a.cgroups + g0] >= 0);
    if (!any) return;
  }
  const int chbeg = a.cg_chunk0[g0], chend = a.cg_chunk0[g1], npass = chend - chbeg;
  const int swz = (g >> 1) & 1;

  if (!scorer) {
    // =========================================== tensor warps ===========================================
    const int wlo = warp & 1, whi = warp >> 1;
    const int fN = tg * BP_ONLD + whi * (BP_OHT * 8) + g;
    const int fNx = tg * BP_ONLD + 2 * BP_OHT * 8 + g;
    const int fX = tg * BP_WLD + wlo * 16 + g;
    const int fXx = tg * BP_WLD + (2 * wlo + whi) * 8 + g;
    const int fY = (whi * (BP_OHT * 8) + g) * BP_OYLD + wlo * 16 + ((2 * tg) ^ (swz << 2));
    const int fYx = (2 * BP_OHT * 8 + g) * BP_OYLD + (2 * wlo + whi) * 8 + ((2 * tg) ^ (swz << 2));
    const int uX = g * BP_WLD + tg;
    const int uY = (warp * (BP_OTW * 8) + g) * BP_OYLD + tg;
    const int wx = min(warp, BP_OREM * BP_XRT - 1);
    const int xct = 4 * BP_OTW + wx / BP_XRT, xrt = wx % BP_XRT;
    const int uXx = (xrt * 8 + g) * BP_WLD + tg;
    const int uYx = (xct * 8 + g) * BP_OYLD + tg;
    // this lane's observation of the next pass (z0, t), loaded one pass ahead
    double nz0[BP_N], nt;
#pragma unroll
    for (int i = 0; i < BP_N; ++i) nz0[i] = a.z0[(size_t)i * a.npad + (size_t)chbeg * 32 + lane];
    nt = a.t[(size_t)chbeg * 32 + lane];
    int pass = 0, y3 = 0, x4 = 0;                     // pass mod 3, pass mod 4
    for (int grp = g0; grp < g1; ++grp) {
      // ---- N_j(tau_g) of the octet's chains, by the 128 tensor threads (as in bp_fusedo) ----
      if (rtid < 8) sjm[rtid] = (o * 8 + rtid < a.nchains) ? a.cg_jm[(size_t)(o * 8 + rtid) * a.cgroups + grp] : -1;
      bp_bar_sync(BP_B_TENSOR, 128);                  // (every tensor warp has finished the previous group's products)
      {
        constexpr int DN = BP_D * BP_N, DD = BP_D * BP_D, NQ = (8 * DN + 127) / 128;
        double Lr[NQ][BP_D];
        int Jq[NQ], eo[NQ], dq[NQ];
        int Jmax = -1;
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) Jmax = max(Jmax, sjm[cc] < 0 ? -1 : (sjm[cc] & 255));
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const int e = rtid + q * 128;
          const int cc = e / DN, rem = e - cc * DN, d = rem / BP_N, l = rem - d * BP_N;
          Jq[q] = e < 8 * DN ? (sjm[cc & 7] < 0 ? -1 : (sjm[cc & 7] & 255)) : -2;
          eo[q] = l * BP_ONLD + cc * BP_D;
          dq[q] = d;
          double n0 = 0.0;
          if (Jq[q] >= 0) {
            const size_t c = (size_t)o * 8 + cc;
            const double* lrow = a.ctab_l + c * (DD + 1) + d * BP_D;
            n0 = a.n0tab[(c * a.cgroups + grp) * DN + rem];
#pragma unroll
            for (int rr = 0; rr < BP_D; ++rr) Lr[q][rr] = lrow[rr];
          } else {
#pragma unroll
            for (int rr = 0; rr < BP_D; ++rr) Lr[q][rr] = 0.0;
          }
          if (Jq[q] >= -1) sN[eo[q] + d] = n0;
        }
        for (int j = 1; j < BP_CJ; ++j) {
          if (j <= Jmax) bp_bar_sync(BP_B_TENSOR, 128);
#pragma unroll
          for (int q = 0; q < NQ; ++q) {
            if (Jq[q] >= -1) {
              double v = 0.0;
              if (j <= Jq[q]) {
                const double* prev = sN + (j - 1) * BP_N * BP_ONLD + eo[q];
#pragma unroll
                for (int rr = 0; rr < BP_D; ++rr) v += Lr[q][rr] * prev[rr];
              }
              sN[j * BP_N * BP_ONLD + eo[q] + dq[q]] = v;
            }
          }
        }
      }
      const double tau = a.cg_tau[grp];
      double accS[BP_XRT][BP_OTW ? BP_OTW : 1][2], accX[2] = {0.0, 0.0};
#pragma unroll
      for (int rt = 0; rt < BP_XRT; ++rt)
#pragma unroll
        for (int ct = 0; ct < (BP_OTW ? BP_OTW : 1); ++ct) { accS[rt][ct][0] = 0.0; accS[rt][ct][1] = 0.0; }
      auto supdate = [&](int xbuf, int ybuf) {
        const double* sX = sXb + xbuf * (BP_XR * BP_WLD);
        const double* px = sX + uX;
        const double* sY = sYb + ybuf * (BP_OU * BP_OYLD);
        const double* pbE = sY + uY + (swz << 2);
        const double* pbO = sY + uY - (swz << 2);
#pragma unroll
        for (int ob = 0; ob < 8; ++ob) {
#if BP_OTW
          double xa[BP_XRT];
#pragma unroll
          for (int rt = 0; rt < BP_XRT; ++rt) xa[rt] = px[rt * 8 * BP_WLD + ob * 4];
#pragma unroll
          for (int ct = 0; ct < BP_OTW; ++ct) {
            const double bf = ((ob & 1) ? pbO : pbE)[ct * 8 * BP_OYLD + ob * 4];
#pragma unroll
            for (int rt = 0; rt < BP_XRT; ++rt) bp_dmma(accS[rt][ct][0], accS[rt][ct][1], xa[rt], bf);
          }
#endif
          bp_dmma(accX[0], accX[1], sX[uXx + ob * 4], sY[uYx + ((ob & 1) ? -(swz << 2) : (swz << 2)) + ob * 4]);
        }
      };
      const int ch0 = a.cg_chunk0[grp], ch1 = a.cg_chunk0[grp + 1];
      for (int ch = ch0; ch < ch1; ++ch, ++pass, y3 = (y3 == 2 ? 0 : y3 + 1), x4 = (x4 + 1) & 3) {
        {   // stage: terms j = 2 warp, 2 warp + 1 of this lane's observation
          const bool valid = ch * 32 + lane < a.nobs;
          double z0[BP_N];
#pragma unroll
          for (int i = 0; i < BP_N; ++i) z0[i] = nz0[i];
          const double h = valid ? nt - tau : 0.0;
          if (ch + 1 < chend) {
#pragma unroll
            for (int i = 0; i < BP_N; ++i) nz0[i] = a.z0[(size_t)i * a.npad + (size_t)(ch + 1) * 32 + lane];
            nt = a.t[(size_t)(ch + 1) * 32 + lane];
          }
          double we = 1.0;
          for (int j = 1; j <= 2 * warp; ++j) we = we * (h * bp_rk[j - 1]);
          const double wo = we * (h * bp_rk[2 * warp]);
          double* xs = sXb + x4 * (BP_XR * BP_WLD) + (2 * warp * BP_N) * BP_WLD + lane;
#pragma unroll
          for (int l = 0; l < BP_N; ++l) {
            xs[l * BP_WLD] = we * z0[l];
            xs[(BP_N + l) * BP_WLD] = wo * z0[l];
          }
        }
        bp_bar_sync(BP_B_TENSOR, 128);                // X (and, for the group's first pass, N) staged; every tensor warp is past the S' update of pass - 3
        {
          double accY[BP_OHT][2][2], accE[2] = {0.0, 0.0};
#pragma unroll
          for (int t = 0; t < BP_OHT; ++t)
#pragma unroll
            for (int q = 0; q < 2; ++q) { accY[t][q][0] = 0.0; accY[t][q][1] = 0.0; }
          const double* sX = sXb + x4 * (BP_XR * BP_WLD);
          double* sY = sYb + y3 * (BP_OU * BP_OYLD);
          const double* pn = sN + fN;
          const double* px = sX + fX;
#pragma unroll
          for (int kk = 0; kk < BP_XKS; ++kk) {
            const double xb0 = px[kk * 4 * BP_WLD], xb1 = px[kk * 4 * BP_WLD + 8];
#pragma unroll
            for (int t = 0; t < BP_OHT; ++t) {
              const double af = pn[kk * 4 * BP_ONLD + t * 8];
              bp_dmma(accY[t][0][0], accY[t][0][1], af, xb0);
              bp_dmma(accY[t][1][0], accY[t][1][1], af, xb1);
            }
#if BP_OEX
            bp_dmma(accE[0], accE[1], sN[fNx + kk * 4 * BP_ONLD], sX[fXx + kk * 4 * BP_WLD]);
#endif
          }
          double* py = sY + fY;
#pragma unroll
          for (int t = 0; t < BP_OHT; ++t)
#pragma unroll
            for (int q = 0; q < 2; ++q) *reinterpret_cast<double2*>(py + t * 8 * BP_OYLD + q * 8) = make_double2(accY[t][q][0], accY[t][q][1]);
#if BP_OEX
          *reinterpret_cast<double2*>(sY + fYx) = make_double2(accE[0], accE[1]);
#endif
        }
        bp_bar_arrive3(BP_B_FULLY, y3);               // Y of this pass is complete
        if (ch - ch0 >= 2) {                          // the S' update of pass - 2
          const int yb = (y3 == 2 ? 0 : y3 + 1);      // (pass - 2) mod 3
          bp_bar_sync3(BP_B_FULLB, yb);               // its cotangent is in its Y buffer
          supdate((x4 + 2) & 3, yb);
        }
      }
      // ---- the group's last two S' updates, then S' to global memory ----
      for (int back = min(2, ch1 - ch0); back >= 1; --back) {
        const int yb = (y3 + 3 - back) % 3;           // (pass - back) mod 3
        bp_bar_sync3(BP_B_FULLB, yb);
        supdate((x4 + 4 - back) & 3, yb);
      }
      double* sg = a.osg + ((size_t)o * a.cgroups + grp) * BP_XR * BP_OU;
#if BP_OTW
#pragma unroll
      for (int rt = 0; rt < BP_XRT; ++rt)
#pragma unroll
        for (int ct = 0; ct < BP_OTW; ++ct)
          __stcg(reinterpret_cast<double2*>(sg + bp_osg_off(rt * 8 + g, (warp * BP_OTW + ct) * 8 + 2 * tg)), make_double2(accS[rt][ct][0], accS[rt][ct][1]));
#endif
      if (warp < BP_OREM * BP_XRT) __stcg(reinterpret_cast<double2*>(sg + bp_osg_off(xrt * 8 + g, xct * 8 + 2 * tg)), make_double2(accX[0], accX[1]));
    }
    return;
  }

  // ============================================= score warps =============================================
  double qs[2] = {0.0, 0.0}, sb[2] = {0.0, 0.0};
  BpDetProd dp[2];
  dp[0].init(); dp[1].init();
  int notpd = 0;
  int pgrp = g0 - 1;                                  // group of the pass being scored
  bool onp[2] = {false, false};                       // chains 2 warp, 2 warp + 1 are on this path in group pgrp
  double nxo[BP_N];                                   // this lane's observed counts of the next pass
#pragma unroll
  for (int i = 0; i < BP_N; ++i) nxo[i] = a.xo[(size_t)i * a.npad + (size_t)chbeg * 32 + lane];
  int y3 = 0;
  for (int p = 0; p < npass; ++p, y3 = (y3 == 2 ? 0 : y3 + 1)) {
    const int ch = chbeg + p;
    const bool valid = ch * 32 + lane < a.nobs;
    double* sY = sYb + y3 * (BP_OU * BP_OYLD);
    while (pgrp < g0 || ch >= a.cg_chunk0[pgrp + 1]) {
      ++pgrp;
#pragma unroll
      for (int i2 = 0; i2 < 2; ++i2) onp[i2] = o * 8 + 2 * warp + i2 < a.nchains && a.cg_jm[(size_t)(o * 8 + 2 * warp + i2) * a.cgroups + pgrp] >= 0;
    }
    double xo[BP_N];
#pragma unroll
    for (int i = 0; i < BP_N; ++i) xo[i] = nxo[i];
    if (p + 1 < npass) {
#pragma unroll
      for (int i = 0; i < BP_N; ++i) nxo[i] = a.xo[(size_t)i * a.npad + (size_t)(ch + 1) * 32 + lane];
    }
    bp_bar_sync3(BP_B_FULLY, y3);                     // Y of this pass is complete
    {
      const int rbase = 2 * warp * BP_D;
      const int kA = lane ^ ((rbase & 2) << 1), kB = kA ^ 4;
      double* yrow = sY + rbase * BP_OYLD;
#pragma unroll
      for (int i2 = 0; i2 < 2; ++i2) {
        double ymu[BP_N], yS[BP_S], bmu[BP_N], bS[BP_S];
        bool ok = valid && onp[i2];
#pragma unroll
        for (int i = 0; i < BP_N; ++i) { ymu[i] = yrow[(i2 * BP_D + i) * BP_OYLD + (((i2 * BP_D + i) & 2) ? kB : kA)]; bmu[i] = 0.0; }
#pragma unroll
        for (int q = 0; q < BP_S; ++q) { yS[q] = yrow[(i2 * BP_D + BP_N + q) * BP_OYLD + (((i2 * BP_D + BP_N + q) & 2) ? kB : kA)]; bS[q] = 0.0; }
        if (ok) {
#if BP_NOISE
          const double sigma = a.theta[(size_t)(o * 8 + 2 * warp + i2) * BP_DIM + BP_P];
#else
          const double sigma = 0.0;
#endif
          if (!bp_score_obs_small<true>(xo, sigma, ymu, yS, qs[i2], dp[i2], sb[i2], bmu, bS)) { notpd |= 1 << i2; ok = false; }
        }
#pragma unroll
        for (int i = 0; i < BP_N; ++i) yrow[(i2 * BP_D + i) * BP_OYLD + (((i2 * BP_D + i) & 2) ? kB : kA)] = ok ? bmu[i] : 0.0;
#pragma unroll
        for (int q = 0; q < BP_S; ++q)
          yrow[(i2 * BP_D + BP_N + q) * BP_OYLD + (((i2 * BP_D + BP_N + q) & 2) ? kB : kA)] = ok ? bS[q] : 0.0;
      }
    }
    bp_bar_arrive3(BP_B_FULLB, y3);                   // the cotangent replaced Y
  }
  // ---- per chain: lp (+ d/dsigma_obs, flops of this CTA's groups) into the tile partial ----
#pragma unroll
  for (int i2 = 0; i2 < 2; ++i2) {
    const double l = bp_warp_sum(-0.5 * (qs[i2] + dp[i2].logdet())), s2 = bp_warp_sum(sb[i2]);
    const int c = o * 8 + 2 * warp + i2;
    if (lane == 0 && c < a.nchains) {
      double fl = 0.0;
      bool mine = false;
      for (int grp = g0; grp < g1; ++grp) { fl += a.cg_flops[(size_t)c * a.cgroups + grp]; mine = mine || a.cg_jm[(size_t)c * a.cgroups + grp] >= 0; }
      if (mine || g0 >= g1 || a.cg_jm[(size_t)c * a.cgroups] >= 0) {
        double* p = a.part + ((size_t)c * a.ntiles + blockIdx.x) * BP_NPART;
        p[0] = l;
#pragma unroll
        for (int k = 0; k < BP_P; ++k) p[1 + k] = 0.0;
        p[BP_P + 1] = s2;
        p[BP_P + 2] = fl;
      }
    }
  }
  notpd = __reduce_or_sync(0xffffffffu, notpd);
  if (lane < 2 && ((notpd >> lane) & 1) && o * 8 + 2 * warp + lane < a.nchains) atomicOr(a.status + o * 8 + 2 * warp + lane, BP_ST_NOT_PD);
}

// ---- bp_fusedo9: the octet passes for n = 3 (D = 9) with TWO barriers per pass ------------------------------------------
// The same arithmetic as bp_fusedo, re-laid so that the warps of a CTA depend on each other as little as possible:
//   * rows of Y / b (and columns of N) are ordered BY CHAIN: tile cc (cc = 0 .. 7) holds components 0 .. 7 of chain cc, tile 8 holds
//     component 8 of the eight chains (BP_O9R).  Warp w owns chains 2w and 2w + 1: it computes their tiles of the forward product for
//     all 32 observations, scores them, and owns the same tiles of the S' update -- so Y and b of tiles 0 .. 7 never cross a warp
//     (__syncwarp instead of a CTA barrier).  Only tile 8 (written by observation tile, read by chain) and X (staged by term, read by
//     all) cross warps; tile 8 is double-buffered.
//   * the S' update of pass p - 1 and the forward product of pass p run back to back (110 tensor products per warp without a barrier
//     in between); X of pass p + 1 is staged, and the x_obs of pass p + 1 fetched into registers, during the score of pass p.
//   pass p:  [S'(p-1); forward(p); store Y(p)]  barrier  [score(p); stage X(p+1)]  barrier  (cp.async of chunk p + 2 issued here)
// S' of every (octet, group) is written with its columns in the same by-chain order (a.operm tells bp_toep).
// MEASURED (cfg5U, 512 chains x 100 000 observations): 1.991 ms per step against 1.965 for bp_fusedo -- no gain: with four CTAs per SM
// the barriers of one CTA are covered by the other three, and what limits the passes is the shared FP64 datapath with four warps per
// scheduler (78 % busy; registers cap the occupancy at 16 warps per SM).  Not the default; BPC_O9=1 selects it (same results, same tests).
#if BP_D == 9
#define BP_O9R(cc, d) ((d) < 8 ? (cc) * 8 + (d) : 64 + (cc))
#define BP_O9SMEM ((BP_XR * BP_ONLD + 2 * BP_XR * BP_WLD + 64 * BP_OYLD + 2 * 8 * BP_OYLD + BP_OOBS) * 8)
extern "C" __global__ void __launch_bounds__(BP_OTHREADS, BP_OMINBLOCKS) bp_fusedo9(BpObsArgs a) {
  bp_pdl_enter();
  const int o = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, tg = lane & 3;
  double* sN = bp_smem;                         // [24][BP_ONLD]        N_j[d][l] of chain cc at [(j, l)][BP_O9R(cc, d)]
  double* sXb = sN + BP_XR * BP_ONLD;           // [2][24][BP_WLD]      X[(j, l)][obs]
  double* sY = sXb + 2 * BP_XR * BP_WLD;        // [64][BP_OYLD]        Y, then b, of tiles 0 .. 7 (swizzled, BP_OY)
  double* sY8 = sY + 64 * BP_OYLD;              // [2][8][BP_OYLD]      tile 8 (component 8 of chain cc in row cc), double-buffered by pass
  double* sOb = sY8 + 2 * 8 * BP_OYLD;          // [7][32]              one chunk of the observation table (z0, xo, t)
  __shared__ int sjm[8];
  const int swz = (g >> 1) & 1;                 // rows with bit 1 set keep column k at k ^ 4
  const int g0 = blockIdx.x * a.cgpc, g1 = min(a.cgroups, g0 + a.cgpc);
  {
    bool any = false;
    if (g0 < g1)
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) any = any || (o * 8 + cc < a.nchains && a.cg_jm[(size_t)(o * 8 + cc) * a.cgroups + g0] >= 0);
    if (!any) return;
  }
  // per-lane bases (doubles)
  const int fN = tg * BP_ONLD + warp * 16 + g;                    // A fragments of the forward product: + kk * 4 * BP_ONLD + t * 8   (t = 0, 1: this warp's tiles)
  const int fN8 = tg * BP_ONLD + 64 + g;                          // ... of tile 8
  const int fX = tg * BP_WLD + g;                                 // B fragments: + kk * 4 * BP_WLD + q * 8   (q = observation tile)
  const int fY = (warp * 16 + g) * BP_OYLD + ((2 * tg) ^ (swz << 2));   // accumulator stores: + t * 8 * BP_OYLD + q * 8
  const int fY8 = g * BP_OYLD + warp * 8 + ((2 * tg) ^ (swz << 2));     // tile 8, observation tile `warp`
  const int uX = g * BP_WLD + tg;                                 // A fragments of the S' update: + rt * 8 * BP_WLD + ob * 4
  const int uY = (warp * 16 + g) * BP_OYLD + tg;                  // B fragments: + ct * 8 * BP_OYLD + (ob ^ swz) * 4
  const int uY8 = g * BP_OYLD + tg;                               // ... of tile 8
  double qs[2] = {0.0, 0.0}, sb[2] = {0.0, 0.0};
  BpDetProd dp[2];
  dp[0].init(); dp[1].init();
  int notpd = 0;
  const int chbeg = a.cg_chunk0[g0], chend = a.cg_chunk0[g1];
  const double* osrc = nullptr;
  {
    const int ai = tid >> 4, q = tid & 15;
    if (ai < 2 * BP_N + 1) osrc = (ai < BP_N ? a.z0 + (size_t)ai * a.npad : (ai < 2 * BP_N ? a.xo + (size_t)(ai - BP_N) * a.npad : a.t)) + 2 * q;
    if (osrc) bp_cp_async16(sOb + tid * 2, osrc + (size_t)chbeg * 32);
    bp_cp_async_commit();
    bp_cp_async_wait_all();
    __syncthreads();
  }
  // X rows of terms j = 2 warp, 2 warp + 1 of this lane's observation, from the chunk in sOb; also fetches this lane's x_obs
  double xo[BP_N];
  auto stage = [&](double* sX, double tau, bool valid) {
    double z0[BP_N];
#pragma unroll
    for (int i = 0; i < BP_N; ++i) { z0[i] = sOb[i * 32 + lane]; xo[i] = sOb[(BP_N + i) * 32 + lane]; }
    const double h = valid ? sOb[2 * BP_N * 32 + lane] - tau : 0.0;
    double we = 1.0;
    for (int j = 1; j <= 2 * warp; ++j) we = we * (h * bp_rk[j - 1]);   // warp-uniform trip count
    const double wo = we * (h * bp_rk[2 * warp]);
    double* xs = sX + (2 * warp * BP_N) * BP_WLD + lane;
#pragma unroll
    for (int l = 0; l < BP_N; ++l) {
      xs[l * BP_WLD] = we * z0[l];
      xs[(BP_N + l) * BP_WLD] = wo * z0[l];
    }
  };
  stage(sXb, a.cg_tau[g0], chbeg * 32 + lane < a.nobs);
  __syncthreads();                               // every thread is done with chunk `chbeg` in sOb
  if (osrc && chbeg + 1 < chend) bp_cp_async16(sOb + tid * 2, osrc + (size_t)(chbeg + 1) * 32);
  bp_cp_async_commit();
  int pass = 0;
  for (int grp = g0; grp < g1; ++grp) {
    // ---- N_j(tau_g) of the octet's chains (as in bp_fusedo; columns in by-chain order) ----
    if (tid < 8) sjm[tid] = (o * 8 + tid < a.nchains) ? a.cg_jm[(size_t)(o * 8 + tid) * a.cgroups + grp] : -1;
    __syncthreads();                             // (also: every warp is done with the previous group's N)
    {
      constexpr int DN = BP_D * BP_N, DD = BP_D * BP_D, NQ = (8 * DN + BP_OTHREADS - 1) / BP_OTHREADS;
      double Lr[NQ][BP_D];
      int Jq[NQ], eo[NQ], cq[NQ], dq[NQ];
      int Jmax = -1;
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) Jmax = max(Jmax, sjm[cc] < 0 ? -1 : (sjm[cc] & 255));
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int e = tid + q * BP_OTHREADS;
        const int cc = e / DN, rem = e - cc * DN, d = rem / BP_N, l = rem - d * BP_N;
        Jq[q] = e < 8 * DN ? (sjm[cc & 7] < 0 ? -1 : (sjm[cc & 7] & 255)) : -2;
        eo[q] = l * BP_ONLD;
        cq[q] = cc & 7;
        dq[q] = d;
        double n0 = 0.0;
        if (Jq[q] >= 0) {
          const size_t c = (size_t)o * 8 + cc;
          const double* lrow = a.ctab_l + c * (DD + 1) + d * BP_D;
          n0 = a.n0tab[(c * a.cgroups + grp) * DN + rem];
#pragma unroll
          for (int rr = 0; rr < BP_D; ++rr) Lr[q][rr] = lrow[rr];
        } else {
#pragma unroll
          for (int rr = 0; rr < BP_D; ++rr) Lr[q][rr] = 0.0;
        }
        if (Jq[q] >= -1) sN[eo[q] + BP_O9R(cq[q], d)] = n0;
      }
      for (int j = 1; j < BP_CJ; ++j) {
        if (j <= Jmax) __syncthreads();
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          if (Jq[q] >= -1) {
            double v = 0.0;
            if (j <= Jq[q]) {
              const double* prev = sN + (j - 1) * BP_N * BP_ONLD + eo[q];
#pragma unroll
              for (int rr = 0; rr < 8; ++rr) v += Lr[q][rr] * prev[cq[q] * 8 + rr];
              v += Lr[q][8] * prev[64 + cq[q]];
            }
            sN[j * BP_N * BP_ONLD + eo[q] + BP_O9R(cq[q], dq[q])] = v;
          }
        }
      }
    }
    __syncthreads();                             // N complete (and X of the group's first pass: staged before a barrier already)
    double accS[BP_XRT][2][2], accX[2] = {0.0, 0.0};
#pragma unroll
    for (int rt = 0; rt < BP_XRT; ++rt)
#pragma unroll
      for (int ct = 0; ct < 2; ++ct) { accS[rt][ct][0] = 0.0; accS[rt][ct][1] = 0.0; }
    // S' update of the pass whose X is in sX and whose b is in sY / y8
    auto supdate = [&](const double* sX, const double* y8) {
      const double* px = sX + uX;
      const double* pbE = sY + uY + (swz << 2);
      const double* pbO = sY + uY - (swz << 2);
      const double* p8E = y8 + uY8 + (swz << 2);
      const double* p8O = y8 + uY8 - (swz << 2);
      const double* px8 = px + (warp < BP_XRT ? warp : BP_XRT - 1) * 8 * BP_WLD;   // tile 8: row tile `warp`
#pragma unroll
      for (int ob = 0; ob < 8; ++ob) {
        double xa[BP_XRT];
#pragma unroll
        for (int rt = 0; rt < BP_XRT; ++rt) xa[rt] = px[rt * 8 * BP_WLD + ob * 4];
#pragma unroll
        for (int ct = 0; ct < 2; ++ct) {
          const double bf = ((ob & 1) ? pbO : pbE)[ct * 8 * BP_OYLD + ob * 4];
#pragma unroll
          for (int rt = 0; rt < BP_XRT; ++rt) bp_dmma(accS[rt][ct][0], accS[rt][ct][1], xa[rt], bf);
        }
        // tile 8: row tile `warp` (warps 0 .. 2; warp 3 repeats row tile 2 and drops the result)
        bp_dmma(accX[0], accX[1], px8[ob * 4], ((ob & 1) ? p8O : p8E)[ob * 4]);
      }
    };
    const double tau = a.cg_tau[grp];
    const int ch1 = a.cg_chunk0[grp + 1];
    bool pend = false;
    for (int ch = a.cg_chunk0[grp]; ch < ch1; ++ch, ++pass) {
      double* sX = sXb + (pass & 1) * (BP_XR * BP_WLD);
      double* y8 = sY8 + (pass & 1) * (8 * BP_OYLD);
      // ---- tensor phase: S' of the previous pass, forward product of this one ----
      if (pend) supdate(sXb + ((pass - 1) & 1) * (BP_XR * BP_WLD), sY8 + ((pass - 1) & 1) * (8 * BP_OYLD));
      __syncwarp();                              // this warp's reads of b (tiles 2w, 2w + 1) are done before its Y stores below
      {
        double accY[2][4][2], accE[2] = {0.0, 0.0};
#pragma unroll
        for (int t = 0; t < 2; ++t)
#pragma unroll
          for (int q = 0; q < 4; ++q) { accY[t][q][0] = 0.0; accY[t][q][1] = 0.0; }
        const double* pn = sN + fN;
        const double* px = sX + fX;
#pragma unroll
        for (int kk = 0; kk < BP_XKS; ++kk) {
          double xb[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) xb[q] = px[kk * 4 * BP_WLD + q * 8];
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const double af = pn[kk * 4 * BP_ONLD + t * 8];
#pragma unroll
            for (int q = 0; q < 4; ++q) bp_dmma(accY[t][q][0], accY[t][q][1], af, xb[q]);
          }
          bp_dmma(accE[0], accE[1], sN[fN8 + kk * 4 * BP_ONLD], px[kk * 4 * BP_WLD + warp * 8]);
        }
        double* py = sY + fY;
#pragma unroll
        for (int t = 0; t < 2; ++t)
#pragma unroll
          for (int q = 0; q < 4; ++q) *reinterpret_cast<double2*>(py + t * 8 * BP_OYLD + q * 8) = make_double2(accY[t][q][0], accY[t][q][1]);
        *reinterpret_cast<double2*>(y8 + fY8) = make_double2(accE[0], accE[1]);
      }
      bp_cp_async_wait_all();                    // the chunk of pass + 1 has landed (this thread's piece; the barrier publishes all of them)
      __syncthreads();                           // tile 8 of Y complete; every warp is done with X / y8 of pass - 1
      // ---- score of this lane's observation under chains 2 warp and 2 warp + 1; the cotangent (packed coordinates) replaces Y ----
      {
        const bool valid = ch * 32 + lane < a.nobs;
        const int k8 = lane ^ ((warp & 1) << 2);   // tile 8: rows 2 warp + i2, r & 2 = 2 (warp & 1)
#pragma unroll
        for (int i2 = 0; i2 < 2; ++i2) {
          double ymu[BP_N], yS[BP_S], bmu[BP_N], bS[BP_S];
          bool ok = valid && sjm[2 * warp + i2] >= 0;
          double* yrow = sY + (2 * warp + i2) * 8 * BP_OYLD;
#pragma unroll
          for (int d = 0; d < 8; ++d) {
            const double v = yrow[d * BP_OYLD + (lane ^ ((d & 2) << 1))];
            if (d < BP_N) ymu[d] = v; else yS[d - BP_N] = v;
          }
          yS[8 - BP_N] = y8[(2 * warp + i2) * BP_OYLD + k8];
#pragma unroll
          for (int i = 0; i < BP_N; ++i) bmu[i] = 0.0;
#pragma unroll
          for (int q = 0; q < BP_S; ++q) bS[q] = 0.0;
          if (ok) {
#if BP_NOISE
            const double sigma = a.theta[(size_t)(o * 8 + 2 * warp + i2) * BP_DIM + BP_P];
#else
            const double sigma = 0.0;
#endif
            if (!bp_score_obs_small<true>(xo, sigma, ymu, yS, qs[i2], dp[i2], sb[i2], bmu, bS)) { notpd |= 1 << i2; ok = false; }
          }
#pragma unroll
          for (int d = 0; d < 8; ++d) {
            const double v = d < BP_N ? bmu[d < BP_N ? d : 0] : bS[d < BP_N ? 0 : d - BP_N];
            yrow[d * BP_OYLD + (lane ^ ((d & 2) << 1))] = ok ? v : 0.0;
          }
          y8[(2 * warp + i2) * BP_OYLD + k8] = ok ? bS[8 - BP_N] : 0.0;
        }
      }
      // ---- X of the next pass (and its x_obs), from the chunk that landed during the tensor phase ----
      if (ch + 1 < chend) stage(sXb + ((pass + 1) & 1) * (BP_XR * BP_WLD), ch + 1 < ch1 ? tau : a.cg_tau[grp + 1], (ch + 1) * 32 + lane < a.nobs);
      __syncthreads();                           // b (tile 8) and X of pass + 1 complete; every thread is done with the chunk in sOb
      if (osrc && ch + 2 < chend) bp_cp_async16(sOb + tid * 2, osrc + (size_t)(ch + 2) * 32);
      bp_cp_async_commit();
      pend = true;
    }
    // ---- the group's last S' update, then S' to global memory: rows (j, l), columns in by-chain order ----
    if (pend) supdate(sXb + ((pass - 1) & 1) * (BP_XR * BP_WLD), sY8 + ((pass - 1) & 1) * (8 * BP_OYLD));
    double* sg = a.osg + ((size_t)o * a.cgroups + grp) * BP_XR * BP_OU;
#pragma unroll
    for (int rt = 0; rt < BP_XRT; ++rt)
#pragma unroll
      for (int ct = 0; ct < 2; ++ct)
        __stcg(reinterpret_cast<double2*>(sg + bp_osg_off(rt * 8 + g, (warp * 2 + ct) * 8 + 2 * tg)), make_double2(accS[rt][ct][0], accS[rt][ct][1]));
    if (warp < BP_XRT) __stcg(reinterpret_cast<double2*>(sg + bp_osg_off(warp * 8 + g, 64 + 2 * tg)), make_double2(accX[0], accX[1]));
  }
  // ---- per chain: lp (+ d/dsigma_obs, flops of this CTA's groups) into the tile partial; chain 2 warp + i belongs to this warp alone ----
#pragma unroll
  for (int i2 = 0; i2 < 2; ++i2) {
    const double l = bp_warp_sum(-0.5 * (qs[i2] + dp[i2].logdet())), s2 = bp_warp_sum(sb[i2]);
    const int c = o * 8 + 2 * warp + i2;
    if (lane == 0 && c < a.nchains) {
      double fl = 0.0;
      bool mine = false;
      for (int grp = g0; grp < g1; ++grp) { fl += a.cg_flops[(size_t)c * a.cgroups + grp]; mine = mine || a.cg_jm[(size_t)c * a.cgroups + grp] >= 0; }
      if (mine || g0 >= g1 || a.cg_jm[(size_t)c * a.cgroups] >= 0) {
        double* p = a.part + ((size_t)c * a.ntiles + blockIdx.x) * BP_NPART;
        p[0] = l;
#pragma unroll
        for (int k = 0; k < BP_P; ++k) p[1 + k] = 0.0;
        p[BP_P + 1] = s2;
        p[BP_P + 2] = fl;
      }
    }
  }
  notpd = __reduce_or_sync(0xffffffffu, notpd);
  if (lane < 2 && ((notpd >> lane) & 1) && o * 8 + 2 * warp + lane < a.nchains) atomicOr(a.status + o * 8 + 2 * warp + lane, BP_ST_NOT_PD);
}
#endif  // BP_D == 9
#endif  // BP_ALT_KERNELS

// W_n = sum_g sum_j c_(n-j)(tau_g) S^g_j for every chain of an octet: one CTA per column tile (8 of the octet's 8 D columns) with all
// n ancestors l; its warps split the groups.  The A fragments (the coefficients c_(n-j)(tau_g)) are shared by the n products of a group,
// the B fragments of the n ancestors are n x 512 contiguous bytes (bp_osg_off), and the (D, octets) grid is ONE wave of 4 CTAs per SM.
// MEASURED on cfg5U (live, BPC_KTIME=1): one CTA per (column tile, l), 8 CTAs per SM, S' row-major: 62 us (29 us of them DMMA; ncu:
// tensor path 48 % busy, long_scoreboard 12 cycles per issue); block layout of S': 57.7; this kernel with 5 CTAs per SM (96 registers,
// 104 bytes spilled): 56, the same with the next group's loads issued ahead: 51.6; with 4 CTAs per SM (no spills): 45.4, with or
// without issuing the loads ahead; 3 CTAs per SM (two waves): 56 .. 64.
// (Not kept earlier: eight warps with four groups in flight, a group-major layout of S', more CTAs per SM through a register cap
// (74 .. 123 us), octets in reverse order (no change: it is not DRAM against L2).)
#ifndef BP_TOEP_MINBLOCKS
#define BP_TOEP_MINBLOCKS 4
#endif
extern "C" __global__ void __launch_bounds__(128, BP_TOEP_MINBLOCKS) bp_toep(BpObsArgs a) {
  bp_pdl_enter();
  const int o = blockIdx.y, ct = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tg = lane & 3;
  __shared__ double red[4][BP_N * BP_CRT * 64];
  const int u = ct * 8 + g;                       // this lane's column of the B fragments: chain cc of the octet
#if BP_D == 9
  const int cc = a.operm ? (u < 64 ? u >> 3 : u - 64) : u / BP_D;
#else
  const int cc = u / BP_D;
#endif
  const int c = o * 8 + cc;
  double accW[BP_N][BP_CRT][2];
#pragma unroll
  for (int l = 0; l < BP_N; ++l)
#pragma unroll
    for (int rt = 0; rt < BP_CRT; ++rt) { accW[l][rt][0] = 0.0; accW[l][rt][1] = 0.0; }
  {   // none of the octet's chains on this path: W stays unwritten (bp_wpost returns for those chains too)
    bool any = false;
#pragma unroll
    for (int cc2 = 0; cc2 < 8; ++cc2) any = any || (o * 8 + cc2 < a.nchains && a.cg_jm[(size_t)(o * 8 + cc2) * a.cgroups] >= 0);
    if (!any) return;
  }
  for (int grp = warp; grp < a.cgroups; grp += 4) {
    const int jm = c < a.nchains ? a.cg_jm[(size_t)c * a.cgroups + grp] : -1;
    const int J = jm < 0 ? -1 : (jm & 255);
    // rows n = 1 .. m_g + J of W get a contribution from this group: row tiles beyond that are skipped (warp-uniform bound)
    const int rows = __reduce_max_sync(0xffffffffu, jm < 0 ? 0 : (jm >> 8) + J);
    const double* sg = a.osg + ((size_t)o * a.cgroups + grp) * BP_XR * BP_OU;
    const double* cf = a.cg_coef + (size_t)grp * BP_CCTAB + 8;
    double bf[BP_N][2];
#pragma unroll
    for (int l = 0; l < BP_N; ++l)
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const int j = ks * 4 + tg;
        const double v = __ldcs(sg + bp_osg_off(j * BP_N + l, u));   // (a warp reads 256 contiguous bytes)
        bf[l][ks] = j <= J ? v : 0.0;   // rows beyond the chain's own degree hold X b products of terms the chain does not have
      }
#pragma unroll
    for (int rt = 0; rt < BP_CRT; ++rt)
      if (rt * 8 < rows) {
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const int ix = (rt * 8 + g + 1) - (ks * 4 + tg);   // c_(n-j)(tau_g), n = 8 rt + g + 1, j = 4 ks + tg (the table is zero outside 0 .. BP_WCAP + 1)
          const double af = cf[ix];
#pragma unroll
          for (int l = 0; l < BP_N; ++l) bp_dmma(accW[l][rt][0], accW[l][rt][1], af, bf[l][ks]);
        }
      }
  }
#pragma unroll
  for (int l = 0; l < BP_N; ++l)
#pragma unroll
    for (int rt = 0; rt < BP_CRT; ++rt) {
      red[warp][(l * BP_CRT + rt) * 64 + lane * 2] = accW[l][rt][0];
      red[warp][(l * BP_CRT + rt) * 64 + lane * 2 + 1] = accW[l][rt][1];
    }
  __syncthreads();
  for (int i = tid; i < BP_N * BP_CRT * 64; i += 128) {
    const double sacc = (red[0][i] + red[1][i]) + (red[2][i] + red[3][i]);
    const int l = i / (BP_CRT * 64), i1 = i - l * (BP_CRT * 64);
    const int rt = i1 >> 6, ln = (i1 & 63) >> 1, e = i1 & 1;
    const int row = rt * 8 + (ln >> 2), uu = ct * 8 + (ln & 3) * 2 + e;
#if BP_D == 9
    const int cc2 = a.operm ? (uu < 64 ? uu >> 3 : uu - 64) : uu / BP_D, d = a.operm ? (uu < 64 ? uu & 7 : 8) : uu % BP_D;
    const int c2 = o * 8 + cc2;
#else
    const int c2 = o * 8 + uu / BP_D, d = uu % BP_D;
#endif
    if (row < BP_CROWS && c2 < a.nchains) a.wplain[((size_t)c2 * BP_CROWS + row) * BP_WCOLS + d * BP_N + l] = sacc;
  }
}
#endif  // BP_N <= 3
#endif
#endif

#if BP_KIND != 2
// ---- fused per-(chain, observation) kernel of the multitype templates ---------------------------
// grid (ntiles, C); CTA = BP_THREADS threads = one tile of one chain; thread i takes observations
// tile*T + i, + BP_THREADS, ... (coalesced SoA loads).  Dynamic shared memory: the Taylor-term store,
// BP_MSLOT * BP_D doubles per thread (a ring of series terms), laid out [slot][component][thread] (conflict-free).
__device__ __forceinline__ void bp_fused_chain(const BpObsArgs& a, const int c) {
  const int tid = threadIdx.x;
  if (bp_skip(a.skip, c)) return;

  double th[BP_DIM];
#pragma unroll
  for (int k = 0; k < BP_DIM; ++k) th[k] = a.theta[(size_t)c * BP_DIM + k];
#if BP_NOISE
  const double sigma = th[BP_P];
#else
  const double sigma = 0.0;
#endif
  double lp = 0.0, sb = 0.0;
  double cb[BP_P];
#pragma unroll
  for (int k = 0; k < BP_P; ++k) cb[k] = 0.0;
  int status = 0, apps = 0, dhint = 1;

  double A[BP_N * BP_N], H[BP_N * BP_S], gA[BP_N * BP_N], gH[BP_N * BP_S];
  double r[BP_E], xk[BP_KDEP];
  double a1 = 0.0;
#pragma unroll
  for (int i = 0; i < BP_N * BP_N; ++i) gA[i] = 0.0;
#pragma unroll
  for (int i = 0; i < BP_N * BP_S; ++i) gH[i] = 0.0;
#if !BP_XDEP
  {
#pragma unroll
    for (int k = 0; k < BP_KDEP; ++k) xk[k] = 0.0;
    bp_rates(th, xk, r);
    bool fin = true;
#pragma unroll
    for (int e = 0; e < BP_E; ++e) fin = fin && isfinite(r[e]);
    if (!fin) status |= BP_ST_RATE_NONFINITE;
    bp_assemble(r, A, H);
    a1 = bp_norm1(A);
    if (a.wmode && bp_w_path(fin, a1, a.nobs > 0 ? a.t[0] : 0.0)) return;   // this chain belongs to bp_fusedw / bp_wpost
  }
#endif
  const int k0 = blockIdx.x * a.tile;
  const int k1 = min(a.nobs, k0 + a.tile);
  if (!(status & BP_ST_RATE_NONFINITE)) {
    for (int k = k0 + tid; k < k1; k += BP_THREADS) {
      double z0[BP_N], xo[BP_N];
#pragma unroll
      for (int i = 0; i < BP_N; ++i) { z0[i] = a.z0[(size_t)i * a.npad + k]; xo[i] = a.xo[(size_t)i * a.npad + k]; }
      const double t = a.t[k];
#if BP_XDEP
#pragma unroll
      for (int q = 0; q < BP_KDEP; ++q) xk[q] = a.xv[(size_t)q * a.npad + k];
      bp_rates(th, xk, r);
      bool fin = true;
#pragma unroll
      for (int e = 0; e < BP_E; ++e) fin = fin && isfinite(r[e]);
      if (!fin) { status |= BP_ST_RATE_NONFINITE; continue; }
      bp_assemble(r, A, H);
      a1 = bp_norm1(A);
#pragma unroll
      for (int i = 0; i < BP_N * BP_N; ++i) gA[i] = 0.0;
#pragma unroll
      for (int i = 0; i < BP_N * BP_S; ++i) gH[i] = 0.0;
#endif
      bp_obs_lpgrad(A, H, a1, z0, xo, t, sigma, tid, BP_THREADS, lp, gA, gH, sb, status, apps, dhint);
#if BP_XDEP
      double rb[BP_E];
      bp_assemble_vjp(gA, gH, rb);
      bp_rates_vjp(th, xk, rb, cb);
#endif
    }
  }
#if !BP_XDEP
  {
    double rb[BP_E];
    bp_assemble_vjp(gA, gH, rb);
    bp_rates_vjp(th, xk, rb, cb);
  }
#endif
  double vals[BP_NPART];
  vals[0] = lp;
#pragma unroll
  for (int k = 0; k < BP_P; ++k) vals[1 + k] = cb[k];
  vals[BP_P + 1] = sb;
  vals[BP_P + 2] = (double)apps;
  __syncthreads();   // the term store is dead: reuse its first words as reduction scratch
  bp_block_reduce_store(vals, bp_smem, a.part + ((size_t)c * a.ntiles + blockIdx.x) * BP_NPART);
  if (status) atomicOr(a.status + c, status);
}

// grid (tiles, chains).  On the centred / octet W paths (wmode >= 2) the chains that need sub-steps are rare: bp_ctable lists them
// (fb[0] = count, fb[1 ..] = chain ids) and a few rows of CTAs walk the list instead of one row per chain exiting at once.
extern "C" __global__ void __launch_bounds__(BP_THREADS, BP_MINBLOCKS) bp_fused(BpObsArgs a) {
  bp_pdl_enter();
  if (a.wmode >= 2) {
    const int nfb = a.fb[0];
    for (int i = blockIdx.y; i < nfb; i += gridDim.y) {
      bp_fused_chain(a, a.fb[1 + i]);
      __syncthreads();
    }
  } else {
    bp_fused_chain(a, blockIdx.y);
  }
}

// ---- grouped two-phase kernel of the multitype templates ------------------------------------------
// The reference computes single-ancestor moments once per (dependent-variable level, duration) and contracts them
// with init_pop[k] for every observation (multitype_template.stan:132-149).  When observations share few distinct
// (level, duration) pairs this kernel does the same: one CTA per chain,
//   phase A : thread (g,l) runs the forward series from (e_l, 0) for group g  -> moments in shared memory
//   phase B : all threads sweep the group's observations: y_k = sum_l z0_kl y^(g,l), score, and accumulate the
//             cotangents ybar^(g,l) += z0_kl ybar_k (block reduction per group, fixed order)
//   phase A': thread (g,l) reverses its series with that cotangent (terms in a global, L2-resident ring) -> Theta-bar
// When every group of the chain needs a single sub-step (the usual case) phases A and A' are done COOPERATIVELY by
// the whole CTA as small dense linear algebra per dependent-variable level (all groups of a level share L):
//   A : Krylov blocks U_j = L^j [e_1 .. e_n] once per level, moments of group g = sum_j c_j(t_g) U_j
//   A': W_j = sum_g c_j(t_g) Ybar_g, then Lbar = sum_i Y_i P^i exactly as in bp_wpost
// which replaces two long single-thread recurrences (the latency of a sampler pass on small data) by ~2 m steps of
// D x n and D x D updates spread over the threads.
// Observations are sorted so that groups are contiguous (gstart).  Dynamic shared memory: 2 * G*n*BP_D doubles (moments,
// cotangents), reduction scratch, and for the cooperative path L, U, W, Z and the per-group series lengths.
struct BpGroupArgs {
  double* theta;         // [C][BP_DIM]  (written when the kernel does its own input / output)
  const double* z0;      // [BP_N][npad]
  const double* xo;      // [BP_N][npad]
  const int* gstart;     // [G + 1]
  const double* gt;      // [G] duration of the group
  const double* gx;      // [BP_KDEP][G] dependent variables of the group's level
  const int* glev;       // [G] level id of the group, 0 .. nlevels-1
  double* ring;          // [C][BP_MSLOT * BP_D * G * BP_N] series terms of phase A'
  double* part;          // [C][1][BP_NPART]
  int* status;           // [C]
  int nobs, npad, ngroups, nchains, nlevels;
  // own input / output (one launch per evaluation: what a sampler pass on small data is made of): u != nullptr
  const double* u;       // [C][BP_DIM] unconstrained parameters; theta is written, not read
  int jacobian;
  double* lp_out; double* grad_out; double* apps_out;
  int yscap;             // cooperative multi-step path: D x n blocks of shared memory for the sub-step states (0: path not available)
  const int* skip;       // [C] or nullptr: see bp_skip
  // phase B on its own grid (data with many observations per group; bp_groupb): mode 0 = bp_grouped does everything (one launch),
  // 1 = phase A only (moments -> gmom, status reset), 2 = phases A and A' around bp_groupb's partial sums
  int mode, ntb;         // ntb: tiles of phase B
  const int* tgroup;     // [ntb] group of the tile
  const int* tstart;     // [ntb + 1] first observation of the tile (tiles of a group are consecutive; tstart[ntb] = nobs)
  const int* gtile0;     // [G + 1] first tile of each group
  double* gmom;          // [C][G * BP_N][BP_D] single-ancestor moments (phase A -> phase B)
  double* ypart;         // [C][ntb][BP_N * BP_D + 2] per-tile sums of the cotangents, lp and d/dsigma_obs (phase B -> phase A');
                         // ypt != 0 (bp_groupc): [ntb][BP_N * BP_D + 2][C]
  int ypt;
};

template <int NV>
__device__ __forceinline__ void bp_block_sum(double (&vals)[NV], double* red, double* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const double s = bp_warp_sum(vals[i]);
    if (lane == 0) red[i * nw + warp] = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NV; i += blockDim.x) {
    double s = 0.0;
    for (int w = 0; w < nw; ++w) s += red[i * nw + w];
    out[i] = s;
  }
  __syncthreads();
}

__device__ __forceinline__ bool bp_group_generator(const double (&th)[BP_DIM], const BpGroupArgs& a, int g, double (&xk)[BP_KDEP],
                                                   double (&A)[BP_N * BP_N], double (&H)[BP_N * BP_S], double& a1) {
  double r[BP_E];
#pragma unroll
  for (int q = 0; q < BP_KDEP; ++q) xk[q] = a.gx[(size_t)q * a.ngroups + g];
  bp_rates(th, xk, r);
  bool fin = true;
#pragma unroll
  for (int e = 0; e < BP_E; ++e) fin = fin && isfinite(r[e]);
  bp_assemble(r, A, H);
  a1 = bp_norm1(A);
  return fin;
}

// phase B for one observation: y_k = sum_l z0_kl y^(g,l), score, cotangents ybar^(g,l) += z0_kl ybar_k
struct BpGroupBAcc {
  double lp, sb;
#if BP_N <= 3
  double qsum;
  BpDetProd dprod;
#endif
  __device__ __forceinline__ void init() {
    lp = 0.0; sb = 0.0;
#if BP_N <= 3
    qsum = 0.0; dprod.init();
#endif
  }
  __device__ __forceinline__ double total() const {
#if BP_N <= 3
    return -0.5 * (qsum + dprod.logdet());
#else
    return lp;
#endif
  }
};

// (mom: [BP_N][BP_D] flattened -- registers in bp_grouped and bp_groupc, shared memory (broadcast reads) in bp_groupb)
__device__ __forceinline__ bool bp_groupb_core(const double (&z0)[BP_N], const double (&xo)[BP_N], double sigma, const double* __restrict__ mom,
                                               double (&acc)[BP_N * BP_D], BpGroupBAcc& s) {
  double ymu[BP_N], yS[BP_S], bmu[BP_N], bS[BP_S];
#pragma unroll
  for (int i = 0; i < BP_N; ++i) {
    double sacc = z0[0] * mom[i];
#pragma unroll
    for (int l = 1; l < BP_N; ++l) sacc += z0[l] * mom[l * BP_D + i];
    ymu[i] = sacc;
  }
#pragma unroll
  for (int q = 0; q < BP_S; ++q) {
    double sacc = z0[0] * mom[BP_N + q];
#pragma unroll
    for (int l = 1; l < BP_N; ++l) sacc += z0[l] * mom[l * BP_D + BP_N + q];
    yS[q] = sacc;
  }
#if BP_N <= 3
  if (!bp_score_obs_small(xo, sigma, ymu, yS, s.qsum, s.dprod, s.sb, bmu, bS)) return false;
#else
  if (!bp_score_obs(xo, sigma, ymu, yS, s.lp, s.sb, bmu, bS)) return false;
#endif
#pragma unroll
  for (int l = 0; l < BP_N; ++l) {
#pragma unroll
    for (int i = 0; i < BP_N; ++i) acc[l * BP_D + i] += z0[l] * bmu[i];
#pragma unroll
    for (int q = 0; q < BP_S; ++q) acc[l * BP_D + BP_N + q] += z0[l] * bS[q];
  }
  return true;
}

__device__ __forceinline__ bool bp_groupb_obs(const BpGroupArgs& a, int k, double sigma, const double* __restrict__ mom, double (&acc)[BP_N * BP_D],
                                              BpGroupBAcc& s) {
  double z0[BP_N], xo[BP_N];
#pragma unroll
  for (int i = 0; i < BP_N; ++i) { z0[i] = a.z0[(size_t)i * a.npad + k]; xo[i] = a.xo[(size_t)i * a.npad + k]; }
  return bp_groupb_core(z0, xo, sigma, mom, acc, s);
}

// phase B of bp_grouped on its own grid (tiles, chains): data with many observations per (level, duration) group.  One CTA per chain
// cannot hide the latency of 10^5 observations (profiles/r02_ncu_bp_grouped_cfg5G.txt: 12 % of the warps, long_scoreboard 6 cycles
// per issue); here every tile of <= 4096 observations of one group is a CTA.  Per-tile sums go to ypart in fixed order (no atomics).
extern "C" __global__ void __launch_bounds__(128, 4) bp_groupb(BpGroupArgs a) {
  bp_pdl_enter();
  const int c = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
  if (bp_skip(a.skip, c)) return;
  __shared__ double red[(BP_N * BP_D + 2) * 4];
  const int g = a.tgroup[tile], k0 = a.tstart[tile], k1 = a.tstart[tile + 1];
#if BP_NOISE
  const double sigma = a.theta[(size_t)c * BP_DIM + BP_P];
#else
  const double sigma = 0.0;
#endif
  __shared__ double mom[BP_N * BP_D];
  double vals[BP_N * BP_D + 2];      // the cotangent sums, then lp and d/dsigma_obs
  double (&acc)[BP_N * BP_D] = *reinterpret_cast<double (*)[BP_N * BP_D]>(vals);
  const double* gm = a.gmom + ((size_t)c * a.ngroups * BP_N + (size_t)g * BP_N) * BP_D;
  if (tid < BP_N * BP_D) mom[tid] = gm[tid];
#pragma unroll
  for (int i = 0; i < BP_N * BP_D; ++i) acc[i] = 0.0;
  __syncthreads();
  BpGroupBAcc s;
  s.init();
  int status = 0;
  for (int k = k0 + tid; k < k1; k += 128)
    if (!bp_groupb_obs(a, k, sigma, mom, acc, s)) status |= BP_ST_NOT_PD;
  vals[BP_N * BP_D] = s.total();
  vals[BP_N * BP_D + 1] = s.sb;
  bp_block_sum(vals, red, a.ypart + ((size_t)c * a.ntb + tile) * (BP_N * BP_D + 2));
  if (status) atomicOr(a.status + c, status);
}

// Phase B with ONE THREAD PER CHAIN (>= 128 chains): a CTA stages a tile of <= BP_GC_TILE observations of one group in shared memory
// and each of its 128 threads sweeps the whole tile for its own chain -- every observation is read from L2 once per 128 chains
// instead of once per chain, all threads read the same shared-memory word (broadcast), they run in lockstep (same observations,
// different parameters), and the cotangent sums of a (tile, chain) are the thread's own: no reduction at all.  The per-tile sums go
// to ypart in the transposed layout [tile][value][chain] (coalesced); bp_grouped mode 2 adds the tiles of every group in tile order.
#define BP_GC_TILE 512
#ifndef BP_GC_MINBLOCKS
#define BP_GC_MINBLOCKS 3
#endif
extern "C" __global__ void __launch_bounds__(128, BP_GC_MINBLOCKS) bp_groupc(BpGroupArgs a) {
  bp_pdl_enter();
  const int tile = blockIdx.x, tid = threadIdx.x, c = blockIdx.y * 128 + tid;
  __shared__ double sz[2 * BP_N][BP_GC_TILE];
  const int g = a.tgroup[tile], k0 = a.tstart[tile], nk = a.tstart[tile + 1] - k0;
  for (int i = tid; i < nk; i += 128) {
#pragma unroll
    for (int q = 0; q < BP_N; ++q) { sz[q][i] = a.z0[(size_t)q * a.npad + k0 + i]; sz[BP_N + q][i] = a.xo[(size_t)q * a.npad + k0 + i]; }
  }
  __syncthreads();
  if (c >= a.nchains || bp_skip(a.skip, c)) return;
#if BP_NOISE
  const double sigma = a.theta[(size_t)c * BP_DIM + BP_P];
#else
  const double sigma = 0.0;
#endif
  double mom[BP_N * BP_D], acc[BP_N * BP_D];
  const double* gm = a.gmom + ((size_t)c * a.ngroups * BP_N + (size_t)g * BP_N) * BP_D;
#pragma unroll
  for (int i = 0; i < BP_N * BP_D; ++i) { mom[i] = gm[i]; acc[i] = 0.0; }
  BpGroupBAcc s;
  s.init();
  int status = 0;
  for (int i = 0; i < nk; ++i) {
    double z0[BP_N], xo[BP_N];
#pragma unroll
    for (int q = 0; q < BP_N; ++q) { z0[q] = sz[q][i]; xo[q] = sz[BP_N + q][i]; }
    if (!bp_groupb_core(z0, xo, sigma, mom, acc, s)) status |= BP_ST_NOT_PD;
  }
  constexpr int NV = BP_N * BP_D + 2;
  double* yp = a.ypart + (size_t)tile * NV * a.nchains + c;
#pragma unroll
  for (int i = 0; i < BP_N * BP_D; ++i) yp[(size_t)i * a.nchains] = acc[i];
  yp[(size_t)(BP_N * BP_D) * a.nchains] = s.total();
  yp[(size_t)(BP_N * BP_D + 1) * a.nchains] = s.sb;
  if (status) atomicOr(a.status + c, status);
}

extern "C" __global__ void __launch_bounds__(128) bp_grouped(BpGroupArgs a) {
  bp_pdl_enter();
  const int c = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
  if (bp_skip(a.skip, c)) return;
  const int G = a.ngroups, GN = G * BP_N;
  double* ymom = bp_smem;                  // [GN][BP_D]
  double* ybar = ymom + (size_t)GN * BP_D; // [GN][BP_D]
  double* red = ybar + (size_t)GN * BP_D;  // [max(BP_N*BP_D, BP_NPART) * warps]
  double th[BP_DIM];
  if (a.u) {   // the transform of bp_prologue, by every thread (theta is kept for the priors)
    bp_transform(a.u + (size_t)c * BP_DIM, th);
    if (tid == 0) {
#pragma unroll
      for (int k = 0; k < BP_DIM; ++k) a.theta[(size_t)c * BP_DIM + k] = th[k];
    }
  } else {
#pragma unroll
    for (int k = 0; k < BP_DIM; ++k) th[k] = a.theta[(size_t)c * BP_DIM + k];
  }
#if BP_NOISE
  const double sigma = th[BP_P];
#else
  const double sigma = 0.0;
#endif
  int status = 0, apps = 0;
  constexpr int DN = BP_D * BP_N, DD = BP_D * BP_D;
  const int nwarp = T >> 5;
  // (layout sized by log_prob_dev; BP_GMS: the cooperative multi-step path with its D x D work arrays, n <= 3)
#if BP_N <= 3
#define BP_GMS 1
  const int WK = a.yscap > 0 ? DD : DN;   // (yscap == 0: the host left the multi-step path's arrays out)
#else
#define BP_GMS 0
  constexpr int WK = DN;
#endif
  double* Lm = red + (size_t)(BP_N * BP_D > BP_NPART ? BP_N * BP_D : BP_NPART) * nwarp;   // [D][D]
  double* Uk = Lm + DD;                         // [BP_WROWS + 1][D * n]   (multi-step path: [BP_WROWS + 1][D * D], powers of L)
  double* Wm = Uk + (size_t)(BP_WROWS + 1) * WK; // [BP_WROWS][D * n]      (multi-step path: [BP_WROWS][D * D])
  double* Zb = Wm + (size_t)BP_WROWS * WK;       // [2][D][D]
#if BP_GMS
  double* Fm = Zb + 2 * DD;                      // [D][D]   exp(h L) of the group at hand
  double* Qm = Fm + DD;                          // [D][D]   sum over sub-steps of z_(i+1) y_i^T
  double* Zv = Qm + DD;                          // [2][D * n] cotangent of the sub-step states
  double* Ys = Zv + 2 * DN;                      // [yscap][D * n] sub-step states Y_0 .. Y_s of every group
  int* gm = a.yscap > 0 ? (int*)(Ys + (size_t)a.yscap * DN) : (int*)(Zb + 2 * DD);   // [G] series length of each group
#else
  int* gm = (int*)(Zb + 2 * DD);                 // [G] series length of each group
#endif
  int* gs = gm + G;                              // [G] sub-steps of each group
  int* goff = gs + G;                            // [G] first block of the group in Ys
  // cooperative path iff every group has finite rates and fits one sub-step; cooperative multi-step path iff the rates are finite
  // and the sub-step states of all groups fit Ys
  bool okc = true, okm = true;
  for (int g = tid; g < G; g += T) {
    double A[BP_N * BP_N], H[BP_N * BP_S], xk[BP_KDEP], a1;
    const bool fin = bp_group_generator(th, a, g, xk, A, H, a1);
    BpPlan p;
    int dh = 1;
    const bool okp = fin && bp_make_plan(a1, a.gt[g], p, dh, BP_WCAP, BP_WTHETA_CAP);
    gm[g] = okp ? p.m : 0;
    gs[g] = okp ? p.s : 0;
    okc = okc && okp && p.s == 1;
    okm = okm && okp;
  }
  const bool coop = __syncthreads_and(okc) != 0;
  bool coopm = false;
#if BP_GMS
  if (!coop) {
    coopm = __syncthreads_and(okm) != 0;
    int tot = 0;
    for (int g = 0; g < G; ++g) tot += gs[g] + 1;
    coopm = coopm && tot <= a.yscap;
    __syncthreads();
    if (coopm && tid == 0) { int o = 0; for (int g = 0; g < G; ++g) { goff[g] = o; o += gs[g] + 1; } }
    __syncthreads();
  }
#endif
  if (coop) {
    // ---- phase A, cooperative: per level, Krylov blocks and the groups' moments ----
    for (int v = 0; v < a.nlevels; ++v) {
      int gv = 0, mlev = 0;
      for (int g = G - 1; g >= 0; --g) if (a.glev[g] == v) { gv = g; mlev = gm[g] > mlev ? gm[g] : mlev; }
      double A[BP_N * BP_N], H[BP_N * BP_S], xk[BP_KDEP], a1;
      bp_group_generator(th, a, gv, xk, A, H, a1);
      __syncthreads();   // previous level's readers of Lm / Uk are done
      bp_coop_build_L(A, H, Lm);
      for (int o = tid; o < DN; o += T) { const int d = o / BP_N, l = o - d * BP_N; Uk[o] = (d == l) ? 1.0 : 0.0; }
      __syncthreads();
      for (int j = 1; j <= mlev; ++j) {
        for (int o = tid; o < DN; o += T) {
          const int d = o / BP_N, l = o - d * BP_N;
          double u = 0.0;
#pragma unroll
          for (int rr = 0; rr < BP_D; ++rr) u += Lm[d * BP_D + rr] * Uk[(j - 1) * DN + rr * BP_N + l];
          Uk[j * DN + o] = u;
        }
        __syncthreads();
      }
      for (int g = 0; g < G; ++g) {
        if (a.glev[g] != v) continue;
        const double tg = a.gt[g];
        for (int o = tid; o < DN; o += T) {
          const int d = o / BP_N, l = o - d * BP_N;
          double y = Uk[o], cj = 1.0;
          for (int j = 1; j <= gm[g]; ++j) { cj = cj * (tg * bp_rk[j - 1]); y += cj * Uk[j * DN + o]; }
          ymom[(size_t)(g * BP_N + l) * BP_D + d] = y;
        }
      }
    }
  }
#if BP_GMS
  else if (coopm) {
    // ---- phase A, cooperative with sub-steps: per level the powers of L; per group F = exp(h L) as a D x D matrix and the
    // sub-step states Y_(i+1) = F Y_i from Y_0 = [e_1 .. e_n] (kept for phase A') ----
    for (int v = 0; v < a.nlevels; ++v) {
      int gv = 0, mlev = 0;
      for (int g = G - 1; g >= 0; --g) if (a.glev[g] == v) { gv = g; mlev = gm[g] > mlev ? gm[g] : mlev; }
      double A[BP_N * BP_N], H[BP_N * BP_S], xk[BP_KDEP], a1;
      bp_group_generator(th, a, gv, xk, A, H, a1);
      __syncthreads();
      bp_coop_build_L(A, H, Lm);
      for (int o = tid; o < DD; o += T) { const int d = o / BP_D, e = o - d * BP_D; Uk[o] = (d == e) ? 1.0 : 0.0; }
      __syncthreads();
      for (int j = 1; j <= mlev; ++j) {
        for (int o = tid; o < DD; o += T) {
          const int d = o / BP_D, e = o - d * BP_D;
          double u = 0.0;
#pragma unroll
          for (int rr = 0; rr < BP_D; ++rr) u += Lm[d * BP_D + rr] * Uk[(j - 1) * DD + rr * BP_D + e];
          Uk[j * DD + o] = u;
        }
        __syncthreads();
      }
      for (int g = 0; g < G; ++g) {
        if (a.glev[g] != v) continue;
        const double hg = a.gt[g] / gs[g];
        for (int o = tid; o < DD; o += T) {
          double y = Uk[o], cj = 1.0;
          for (int j = 1; j <= gm[g]; ++j) { cj = cj * (hg * bp_rk[j - 1]); y += cj * Uk[j * DD + o]; }
          Fm[o] = y;
        }
        double* Y = Ys + (size_t)goff[g] * DN;
        for (int o = tid; o < DN; o += T) { const int d = o / BP_N, l = o - d * BP_N; Y[o] = (d == l) ? 1.0 : 0.0; }
        __syncthreads();
        for (int i = 0; i < gs[g]; ++i) {
          for (int o = tid; o < DN; o += T) {
            const int d = o / BP_N, l = o - d * BP_N;
            double y = 0.0;
#pragma unroll
            for (int rr = 0; rr < BP_D; ++rr) y += Fm[d * BP_D + rr] * Y[i * DN + rr * BP_N + l];
            Y[(i + 1) * DN + o] = y;
          }
          __syncthreads();
        }
        for (int o = tid; o < DN; o += T) {
          const int d = o / BP_N, l = o - d * BP_N;
          ymom[(size_t)(g * BP_N + l) * BP_D + d] = Y[gs[g] * DN + o];
        }
        __syncthreads();   // Fm is rewritten for the next group
      }
    }
  }
#endif
  else {
  // ---- phase A: single-ancestor moments of every (group, ancestor type) ----
  for (int sx = tid; sx < GN; sx += T) {
    const int g = sx / BP_N, l = sx - g * BP_N;
    double A[BP_N * BP_N], H[BP_N * BP_S], xk[BP_KDEP], a1, ymu[BP_N], yS[BP_S];
    if (!bp_group_generator(th, a, g, xk, A, H, a1)) status |= BP_ST_RATE_NONFINITE;
#pragma unroll
    for (int i = 0; i < BP_N; ++i) ymu[i] = (i == l) ? 1.0 : 0.0;
#pragma unroll
    for (int q = 0; q < BP_S; ++q) yS[q] = 0.0;
    BpPlan p;
    int dh = 1;
    if (!bp_make_plan(a1, a.gt[g], p, dh)) { status |= BP_ST_RATE_NONFINITE; p.s = 0; }
    double cfw;
    for (int j = 0; j < p.s; ++j) bp_series_fwd<false>(A, H, p.h, p.m, ymu, yS, true, ymu, yS, false, BpRingShared(), 0, 0, cfw);
#pragma unroll
    for (int i = 0; i < BP_N; ++i) ymom[(size_t)sx * BP_D + i] = ymu[i];
#pragma unroll
    for (int q = 0; q < BP_S; ++q) ymom[(size_t)sx * BP_D + BP_N + q] = yS[q];
  }
  }
  __syncthreads();
  if (a.mode == 1) {   // phase A only: the moments go to global memory for bp_groupb; the status is reset before its CTAs OR into it
    double* gm = a.gmom + (size_t)c * GN * BP_D;
    for (int i = tid; i < GN * BP_D; i += T) gm[i] = ymom[i];
    if (tid == 0) a.status[c] = 0;
    __syncthreads();
    if (status) atomicOr(a.status + c, status);
    return;
  }
  // ---- phase B: observations, group by group ----
  double lp = 0.0, sb = 0.0;
  if (a.mode == 2) {
    // bp_groupb did phase B: sum its per-tile cotangents group by group in tile order; lp and sb by thread 0
    constexpr int NV = BP_N * BP_D + 2;
    // element e of tile tI of this chain: [C][ntb][NV] (bp_groupb) or [ntb][NV][C] (bp_groupc)
    const size_t st_t = a.ypt ? (size_t)NV * a.nchains : (size_t)NV, st_e = a.ypt ? (size_t)a.nchains : (size_t)1;
    const double* yp = a.ypart + (a.ypt ? (size_t)c : (size_t)c * a.ntb * NV);
    for (int i = tid; i < GN * BP_D; i += T) {
      const int g = i / (BP_N * BP_D), e = i - g * (BP_N * BP_D);
      double sacc = 0.0;
      for (int tI = a.gtile0[g]; tI < a.gtile0[g + 1]; ++tI) sacc += yp[(size_t)tI * st_t + e * st_e];
      ybar[i] = sacc;
    }
    if (tid < 2) {          // thread 0: lp, thread 1: d/dsigma_obs (both enter the block sums below from one thread each)
      double sacc = 0.0;
      for (int tI = 0; tI < a.ntb; ++tI) sacc += yp[(size_t)tI * st_t + (BP_N * BP_D + tid) * st_e];
      if (tid == 0) lp = sacc; else sb = sacc;
    }
    __syncthreads();
  } else {
    BpGroupBAcc sB;
    sB.init();
    for (int g = 0; g < G; ++g) {
      double mom[BP_N * BP_D], acc[BP_N * BP_D];
#pragma unroll
      for (int i = 0; i < BP_N * BP_D; ++i) { mom[i] = ymom[(size_t)g * BP_N * BP_D + i]; acc[i] = 0.0; }
      const int k1 = a.gstart[g + 1];
      for (int k = a.gstart[g] + tid; k < k1; k += T)
        if (!bp_groupb_obs(a, k, sigma, mom, acc, sB)) status |= BP_ST_NOT_PD;
      bp_block_sum(acc, red, ybar + (size_t)g * BP_N * BP_D);
    }
    lp = sB.total();
    sb = sB.sb;
  }
  // ---- phase A': reverse sweep of every series with the reduced cotangent ----
  double cb[BP_P];
#pragma unroll
  for (int k = 0; k < BP_P; ++k) cb[k] = 0.0;
  if (coop) {
    for (int v = 0; v < a.nlevels; ++v) {
      int gv = 0, mlev = 0;
      for (int g = G - 1; g >= 0; --g) if (a.glev[g] == v) { gv = g; mlev = gm[g] > mlev ? gm[g] : mlev; }
      double A[BP_N * BP_N], H[BP_N * BP_S], xk[BP_KDEP], a1;
      bp_group_generator(th, a, gv, xk, A, H, a1);
      __syncthreads();
      bp_coop_build_L(A, H, Lm);
      // W_(j+1)[d][l] = sum over the level's groups of c_(j+1)(t_g) * (packed-coordinate cotangent of y^(g,l))_d
      for (int o = tid; o < DN; o += T) {
        const int d = o / BP_N, l = o - d * BP_N;
        bool offdiag = d >= BP_N;
#pragma unroll
        for (int i = 0; i < BP_N; ++i) if (d == BP_N + bp_sidx(i, i)) offdiag = false;
        for (int j = 0; j < mlev; ++j) Wm[j * DN + o] = 0.0;
        for (int g = 0; g < G; ++g) {
          if (a.glev[g] != v) continue;
          const double tg = a.gt[g];
          double yb = ybar[(size_t)(g * BP_N + l) * BP_D + d];
          if (offdiag) yb += yb;
          double cj = 1.0;
          for (int j = 0; j < gm[g]; ++j) { cj = cj * (tg * bp_rk[j]); Wm[j * DN + o] += cj * yb; }
        }
      }
      __syncthreads();
      const int cur = bp_coop_lbar(Wm, mlev, Lm, Zb);
      if (tid == 0) {
        double gA[BP_N * BP_N], gH[BP_N * BP_S], rb[BP_E];
#pragma unroll
        for (int i = 0; i < BP_N * BP_N; ++i) gA[i] = 0.0;
#pragma unroll
        for (int i = 0; i < BP_N * BP_S; ++i) gH[i] = 0.0;
        bp_lbar_to_grad(Zb + cur * DD, gA, gH);
        bp_assemble_vjp(gA, gH, rb);
        bp_rates_vjp(th, xk, rb, cb);
        for (int g = 0; g < G; ++g) if (a.glev[g] == v) apps += BP_N * gm[g];
      }
    }
  }
#if BP_GMS
  else if (coopm) {
    // reverse of the sub-steps: z_s = cotangent of Y_s; for i = s - 1 .. 0: Q += z_(i+1) y_i^T, z_i = F^T z_(i+1); every sub-step
    // applies the same series F = sum_j c_j(h) L^j, so W_j = c_j(h) Q (D x D) and Lbar = the matrix polynomial of bp_coop_lbar_full
    for (int v = 0; v < a.nlevels; ++v) {
      int gv = 0, mlev = 0;
      for (int g = G - 1; g >= 0; --g) if (a.glev[g] == v) { gv = g; mlev = gm[g] > mlev ? gm[g] : mlev; }
      double A[BP_N * BP_N], H[BP_N * BP_S], xk[BP_KDEP], a1;
      bp_group_generator(th, a, gv, xk, A, H, a1);
      __syncthreads();
      bp_coop_build_L(A, H, Lm);
      for (int o = tid; o < DD; o += T) {
        const int d = o / BP_D, e = o - d * BP_D;
        Uk[o] = (d == e) ? 1.0 : 0.0;
        for (int j = 0; j < mlev; ++j) Wm[j * DD + o] = 0.0;
      }
      __syncthreads();
      for (int j = 1; j <= mlev; ++j) {
        for (int o = tid; o < DD; o += T) {
          const int d = o / BP_D, e = o - d * BP_D;
          double u = 0.0;
#pragma unroll
          for (int rr = 0; rr < BP_D; ++rr) u += Lm[d * BP_D + rr] * Uk[(j - 1) * DD + rr * BP_D + e];
          Uk[j * DD + o] = u;
        }
        __syncthreads();
      }
      for (int g = 0; g < G; ++g) {
        if (a.glev[g] != v) continue;
        const double hg = a.gt[g] / gs[g];
        for (int o = tid; o < DD; o += T) {
          double y = Uk[o], cj = 1.0;
          for (int j = 1; j <= gm[g]; ++j) { cj = cj * (hg * bp_rk[j - 1]); y += cj * Uk[j * DD + o]; }
          Fm[o] = y;
          Qm[o] = 0.0;
        }
        for (int o = tid; o < DN; o += T) {   // packed-coordinate cotangent of Y_s
          const int d = o / BP_N, l = o - d * BP_N;
          bool offdiag = d >= BP_N;
#pragma unroll
          for (int i = 0; i < BP_N; ++i) if (d == BP_N + bp_sidx(i, i)) offdiag = false;
          double yb = ybar[(size_t)(g * BP_N + l) * BP_D + d];
          if (offdiag) yb += yb;
          Zv[o] = yb;
        }
        __syncthreads();
        const double* Y = Ys + (size_t)goff[g] * DN;
        int cz = 0;
        for (int i = gs[g] - 1; i >= 0; --i) {
          for (int o = tid; o < DD; o += T) {
            const int d = o / BP_D, e = o - d * BP_D;
            double q = Qm[o];
#pragma unroll
            for (int l = 0; l < BP_N; ++l) q += Zv[cz * DN + d * BP_N + l] * Y[i * DN + e * BP_N + l];
            Qm[o] = q;
          }
          for (int o = tid; o < DN; o += T) {
            const int d = o / BP_N, l = o - d * BP_N;
            double z = 0.0;
#pragma unroll
            for (int rr = 0; rr < BP_D; ++rr) z += Fm[rr * BP_D + d] * Zv[cz * DN + rr * BP_N + l];
            Zv[(cz ^ 1) * DN + o] = z;
          }
          __syncthreads();
          cz ^= 1;
        }
        for (int o = tid; o < DD; o += T) {
          const double q = Qm[o];
          double cj = 1.0;
          for (int j = 0; j < gm[g]; ++j) { cj = cj * (hg * bp_rk[j]); Wm[j * DD + o] += cj * q; }
        }
        __syncthreads();   // Fm, Qm, Zv are rewritten for the next group
      }
      const int cur = bp_coop_lbar_full(Wm, mlev, Lm, Zb);
      if (tid == 0) {
        double gA[BP_N * BP_N], gH[BP_N * BP_S], rb[BP_E];
#pragma unroll
        for (int i = 0; i < BP_N * BP_N; ++i) gA[i] = 0.0;
#pragma unroll
        for (int i = 0; i < BP_N * BP_S; ++i) gH[i] = 0.0;
        bp_lbar_to_grad(Zb + cur * DD, gA, gH);
        bp_assemble_vjp(gA, gH, rb);
        bp_rates_vjp(th, xk, rb, cb);
        for (int g = 0; g < G; ++g) if (a.glev[g] == v) apps += BP_N * gs[g] * gm[g];
      }
    }
  }
#endif
  else {
  BpRingGlobal ring{a.ring + (size_t)c * BP_MSLOT * BP_D * GN};
  for (int sx = tid; sx < GN; sx += T) {
    const int g = sx / BP_N, l = sx - g * BP_N;
    double A[BP_N * BP_N], H[BP_N * BP_S], gA[BP_N * BP_N], gH[BP_N * BP_S], xk[BP_KDEP], a1, imu[BP_N], iS[BP_S], rb[BP_E];
    if (!bp_group_generator(th, a, g, xk, A, H, a1)) continue;
#pragma unroll
    for (int i = 0; i < BP_N * BP_N; ++i) gA[i] = 0.0;
#pragma unroll
    for (int i = 0; i < BP_N * BP_S; ++i) gH[i] = 0.0;
#pragma unroll
    for (int i = 0; i < BP_N; ++i) imu[i] = (i == l) ? 1.0 : 0.0;
#pragma unroll
    for (int q = 0; q < BP_S; ++q) iS[q] = 0.0;
    BpPlan p;
    int dh = 1;
    if (!bp_make_plan(a1, a.gt[g], p, dh)) continue;
    apps += p.s * p.m;
    const double* yb = ybar + (size_t)sx * BP_D;
    bp_series_fwd_bwd(A, H, p, imu, iS,
                      [&](double (&)[BP_N], double (&)[BP_S], double (&bmu)[BP_N], double (&bS)[BP_S]) {
#pragma unroll
                        for (int i = 0; i < BP_N; ++i) bmu[i] = yb[i];
#pragma unroll
                        for (int q = 0; q < BP_S; ++q) bS[q] = yb[BP_N + q];
                        return true;
                      },
                      ring, sx, GN, gA, gH);
    bp_assemble_vjp(gA, gH, rb);
    bp_rates_vjp(th, xk, rb, cb);
  }
  }
  double vals[BP_NPART];
  vals[0] = lp;
#pragma unroll
  for (int k = 0; k < BP_P; ++k) vals[1 + k] = cb[k];
  vals[BP_P + 1] = sb;
  vals[BP_P + 2] = (double)apps;
  if (a.u) {
    // the chain's status: OR over the CTA; then the work of bp_epilogue by one thread
    __shared__ int sst;
    if (tid == 0) sst = a.mode == 2 ? a.status[c] : 0;   // (mode 2: the bits phase A and bp_groupb left)
    __syncthreads();
    if (status) atomicOr(&sst, status);
    bp_block_sum(vals, red, a.part + (size_t)c * BP_NPART);   // (ends with a barrier; the sums are also in global memory)
    if (tid == 0) {
      double acc[BP_NPART];
#pragma unroll
      for (int i = 0; i < BP_NPART; ++i) acc[i] = __ldcg(a.part + (size_t)c * BP_NPART + i);
      bp_finish(c, acc, sst, a.u, a.theta, a.jacobian, a.lp_out, a.grad_out, a.status, a.apps_out);
    }
    return;
  }
  bp_block_sum(vals, red, a.part + (size_t)c * BP_NPART);
  if (status) atomicOr(a.status + c, status);
}

// ---- per-observation log_lik (LOO block), forward only -----------------------------------------
extern "C" __global__ void __launch_bounds__(BP_THREADS) bp_loglik(BpObsArgs a) {
  const int c = blockIdx.y;
  const int k = blockIdx.x * BP_THREADS + threadIdx.x;
  if (k >= a.nobs) return;
  double th[BP_DIM];
#pragma unroll
  for (int q = 0; q < BP_DIM; ++q) th[q] = a.theta[(size_t)c * BP_DIM + q];
#if BP_NOISE
  const double sigma = th[BP_P];
#else
  const double sigma = 0.0;
#endif
  double A[BP_N * BP_N], H[BP_N * BP_S], r[BP_E], xk[BP_KDEP], z0[BP_N], xo[BP_N];
#pragma unroll
  for (int q = 0; q < BP_KDEP; ++q) xk[q] = 0.0;
#if BP_XDEP
#pragma unroll
  for (int q = 0; q < BP_KDEP; ++q) xk[q] = a.xv[(size_t)q * a.npad + k];
#endif
  bp_rates(th, xk, r);
  bp_assemble(r, A, H);
  const double a1 = bp_norm1(A);
#pragma unroll
  for (int i = 0; i < BP_N; ++i) { z0[i] = a.z0[(size_t)i * a.npad + k]; xo[i] = a.xo[(size_t)i * a.npad + k]; }
  a.loglik[(size_t)c * a.nobs + a.perm[k]] = bp_obs_loglik(A, H, a1, z0, xo, a.t[k], sigma);
}

// ---- mean vector and covariance matrix of every (chain, observation): what calculate_moments returns
// (R/calculate_moments.R:79-86).  out: [C][N][BP_N + BP_N*BP_N] = mu then Sigma (row-major), original order.
extern "C" __global__ void __launch_bounds__(BP_THREADS) bp_moments(BpObsArgs a) {
  const int c = blockIdx.y;
  const int k = blockIdx.x * BP_THREADS + threadIdx.x;
  if (k >= a.nobs) return;
  double th[BP_DIM];
#pragma unroll
  for (int q = 0; q < BP_DIM; ++q) th[q] = a.theta[(size_t)c * BP_DIM + q];
  double A[BP_N * BP_N], H[BP_N * BP_S], r[BP_E], xk[BP_KDEP], ymu[BP_N], yS[BP_S];
#pragma unroll
  for (int q = 0; q < BP_KDEP; ++q) xk[q] = 0.0;
#if BP_XDEP
#pragma unroll
  for (int q = 0; q < BP_KDEP; ++q) xk[q] = a.xv[(size_t)q * a.npad + k];
#endif
  bp_rates(th, xk, r);
  bp_assemble(r, A, H);
  const double a1 = bp_norm1(A);
#pragma unroll
  for (int i = 0; i < BP_N; ++i) ymu[i] = a.z0[(size_t)i * a.npad + k];
#pragma unroll
  for (int q = 0; q < BP_S; ++q) yS[q] = 0.0;
  BpPlan p;
  int dhint = 1;
  const bool ok = bp_make_plan(a1, a.t[k], p, dhint);
  double cfw;
  if (ok) for (int j = 0; j < p.s; ++j) bp_series_fwd<false>(A, H, p.h, p.m, ymu, yS, true, ymu, yS, false, BpRingShared(), 0, 0, cfw);
  double* o = a.loglik + ((size_t)c * a.nobs + a.perm[k]) * (BP_N + BP_N * BP_N);
#pragma unroll
  for (int i = 0; i < BP_N; ++i) o[i] = ok ? ymu[i] : (0.0 / 0.0);
#pragma unroll
  for (int i = 0; i < BP_N; ++i)
#pragma unroll
    for (int j = 0; j < BP_N; ++j) o[BP_N + i * BP_N + j] = ok ? yS[bp_sidx(i, j)] : (0.0 / 0.0);
}
#else   // BP_KIND == 2
// ---- one-type template: thread per (chain, observation) -----------------------------------------
extern "C" __global__ void __launch_bounds__(BP_THREADS) bp_onetype(BpObsArgs a) {
  bp_pdl_enter();
  __shared__ double scratch[BP_NPART * (BP_THREADS / 32)];
  const int c = blockIdx.y;
  const int tid = threadIdx.x;
  if (bp_skip(a.skip, c)) return;
  double th[BP_DIM];
  if (a.u) {   // the transform of bp_prologue, by every thread (theta is kept for the priors)
    bp_transform(a.u + (size_t)c * BP_DIM, th);
    if (tid == 0) {
#pragma unroll
      for (int k = 0; k < BP_DIM; ++k) a.theta[(size_t)c * BP_DIM + k] = th[k];
    }
  } else {
#pragma unroll
    for (int k = 0; k < BP_DIM; ++k) th[k] = a.theta[(size_t)c * BP_DIM + k];
  }
  double lp = 0.0;
  double cb[BP_P], r[BP_E], rb[2] = {0.0, 0.0}, xk[BP_KDEP];
#pragma unroll
  for (int k = 0; k < BP_P; ++k) cb[k] = 0.0;
#pragma unroll
  for (int k = 0; k < BP_KDEP; ++k) xk[k] = 0.0;
  int status = 0;
#if !BP_XDEP
  bp_rates(th, xk, r);
  if (!isfinite(r[0]) || !isfinite(r[1])) status |= BP_ST_RATE_NONFINITE;
#endif
  const int k0 = blockIdx.x * a.tile;
  const int k1 = min(a.nobs, k0 + a.tile);
  if (!status) {
    for (int k = k0 + tid; k < k1; k += BP_THREADS) {
#if BP_XDEP
#pragma unroll
      for (int q = 0; q < BP_KDEP; ++q) xk[q] = a.xv[(size_t)q * a.npad + k];
      bp_rates(th, xk, r);
      if (!isfinite(r[0]) || !isfinite(r[1])) { status |= BP_ST_RATE_NONFINITE; continue; }
      rb[0] = 0.0; rb[1] = 0.0;
#endif
      bp_onetype_obs(r[0], r[1], a.t[k], a.z0[k], a.xo[k], lp, rb, status, false);
#if BP_XDEP
      bp_rates_vjp(th, xk, rb, cb);
#endif
    }
  }
#if !BP_XDEP
  bp_rates_vjp(th, xk, rb, cb);
#endif
  double vals[BP_NPART];
  vals[0] = lp;
#pragma unroll
  for (int k = 0; k < BP_P; ++k) vals[1 + k] = cb[k];
  vals[BP_P + 1] = 0.0;
  vals[BP_P + 2] = 0.0;
  if (a.u) {   // one tile per chain: the chain's status (OR over the CTA), its sums, then the work of bp_epilogue by one thread
    __shared__ int sst;
    if (tid == 0) sst = 0;
    __syncthreads();
    if (status) atomicOr(&sst, status);
    bp_block_reduce_store(vals, scratch, a.part + (size_t)c * BP_NPART);
    __syncthreads();
    if (tid == 0) {
      double acc[BP_NPART];
#pragma unroll
      for (int i = 0; i < BP_NPART; ++i) acc[i] = __ldcg(a.part + (size_t)c * BP_NPART + i);
      bp_finish(c, acc, sst, a.u, a.theta, a.jacobian, a.lp_out, a.grad_out, a.status, a.apps_out);
    }
    return;
  }
  bp_block_reduce_store(vals, scratch, a.part + ((size_t)c * a.ntiles + blockIdx.x) * BP_NPART);
  if (status) atomicOr(a.status + c, status);
}

extern "C" __global__ void __launch_bounds__(BP_THREADS) bp_loglik(BpObsArgs a) {
  const int c = blockIdx.y;
  const int k = blockIdx.x * BP_THREADS + threadIdx.x;
  if (k >= a.nobs) return;
  double th[BP_DIM], r[BP_E], xk[BP_KDEP], rb[2] = {0.0, 0.0};
#pragma unroll
  for (int q = 0; q < BP_DIM; ++q) th[q] = a.theta[(size_t)c * BP_DIM + q];
#pragma unroll
  for (int q = 0; q < BP_KDEP; ++q) xk[q] = 0.0;
#if BP_XDEP
#pragma unroll
  for (int q = 0; q < BP_KDEP; ++q) xk[q] = a.xv[(size_t)q * a.npad + k];
#endif
  bp_rates(th, xk, r);
  double lp = 0.0;
  int st = 0;
  bp_onetype_obs(r[0], r[1], a.t[k], a.z0[k], a.xo[k], lp, rb, st, true);
  a.loglik[(size_t)c * a.nobs + a.perm[k]] = st ? (0.0 / 0.0) : lp;
}
#endif  // BP_KIND

// ---- epilogue: one warp per chain -----------------------------------------------------------------
// sums the tile partials in a fixed order (lane i takes tiles i, i + 32, ..., then a shuffle tree), adds priors and the
// log-Jacobian, applies the chain rule of the bound transform, and rejects the chain (lp = -inf, grad = 0) when any status bit is set.
extern "C" __global__ void bp_epilogue(const double* __restrict__ u, const double* __restrict__ theta,
                                       const double* __restrict__ part, int ntiles, int nchains, int jacobian,
                                       double* __restrict__ lp_out, double* __restrict__ grad_out, int* __restrict__ status,
                                       double* __restrict__ apps_out, const int* __restrict__ skip) {
  bp_pdl_enter();
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= nchains || bp_skip(skip, c)) return;   // (warp-uniform)
  double acc[BP_NPART];
#pragma unroll
  for (int i = 0; i < BP_NPART; ++i) acc[i] = 0.0;
  for (int tI = lane; tI < ntiles; tI += 32) {
    const double* p = part + ((size_t)c * ntiles + tI) * BP_NPART;
#pragma unroll
    for (int i = 0; i < BP_NPART; ++i) acc[i] += p[i];
  }
#pragma unroll
  for (int i = 0; i < BP_NPART; ++i) acc[i] = bp_warp_sum(acc[i]);
  if (lane) return;
  bp_finish(c, acc, status[c], u, theta, jacobian, lp_out, grad_out, status, apps_out);
}

// ---- transformed parameters: the rates of every (draw, dependent-variable level) -------------------------------
// what the `transformed parameters` block stores per draw (multitype_template.stan:105-117, one_type_template.stan:19-27):
// r_e = f_e(Theta, function_var[v, ]) from the generated rate functions (so Stan's typing quirks are the same as on the
// likelihood path).  theta: [ndraws][BP_DIM] CONSTRAINED values; fvar: function_var, nlev x ncols column-major; out: [ndraws][nlev][BP_E].
extern "C" __global__ void bp_tparams(const double* __restrict__ theta, int ndraws, const double* __restrict__ fvar, int nlev, int ncols,
                                      double* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)ndraws * nlev) return;
  const long long dr = idx / nlev;
  const int v = (int)(idx - dr * nlev);
  double th[BP_DIM], xk[BP_KDEP], r[BP_E];
#pragma unroll
  for (int k = 0; k < BP_DIM; ++k) th[k] = theta[dr * BP_DIM + k];
#pragma unroll
  for (int q = 0; q < BP_KDEP; ++q) xk[q] = q < ncols ? fvar[v + (size_t)nlev * q] : 0.0;
  bp_rates(th, xk, r);
#pragma unroll
  for (int e = 0; e < BP_E; ++e) out[idx * BP_E + e] = r[e];
}
#endif  // !BP_HOSTSIM

#endif  // BP_KERNELS_CUH
