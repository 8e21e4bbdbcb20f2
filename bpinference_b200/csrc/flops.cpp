// flops.cpp — algorithmic operation counts of the shipped kernels (bp_kernels.cuh), by walking the same loops
// with the same structural-sparsity masks.  Convention of SURVEY.md §8(d): add/mul = 1, FMA = 2,
// div/sqrt/exp/log = 1.  tests/test_kernel_core_host.py recounts the same device functions with a counting
// scalar and pins these numbers to 1 %.
#include "bpcuda_internal.hpp"

namespace bpc {

namespace {
inline double acc_chain(int terms) { return terms > 0 ? 2.0 * terms - 1.0 : 0.0; }   // first product is a multiply
}

FlopModel flop_model(const ModelSpec& sp) {
  const int n = sp.ntypes, S = n * (n + 1) / 2, D = n + S;
  FlopModel f;
  if (sp.kind == 2) {
    // bp_onetype_obs: value + gradient w.r.t. (b, d)
    f.per_app_fwd = f.per_app_bwd = 0.0;
    f.per_obs = 70.0;
    f.per_obs_grouped = 70.0;
    f.per_app_w = 0.0;
    f.per_obs_w = 70.0;
    f.per_term_c = 0.0;
    f.per_obs_c = 70.0;
    f.per_obs_o = 70.0;
    return f;
  }
  auto sidx = [n](int i, int j) { if (i > j) std::swap(i, j); return i * n - i * (i - 1) / 2 + (j - i); };
  auto anz = [&](int i, int j) { return sp.a_nz.empty() ? true : (bool)sp.a_nz[i * n + j]; };
  auto hnz = [&](int l, int q) { return sp.h_nz.empty() ? true : (bool)sp.h_nz[l * S + q]; };
  // ---- bp_L_apply
  double L = 0;
  for (int j = 0; j < n; ++j) { int t = 0; for (int i = 0; i < n; ++i) t += anz(i, j); L += acc_chain(t); }
  for (int i = 0; i < n; ++i)
    for (int j = i; j < n; ++j) {
      int t = 0;
      for (int l = 0; l < n; ++l) t += hnz(l, sidx(i, j));
      if (i == j) {
        int td = 0;
        for (int k = 0; k < n; ++k) td += anz(k, i);
        L += acc_chain(t) + acc_chain(td) + (td > 0 ? (t > 0 ? 2 : 1) : 0);
      } else {
        for (int k = 0; k < n; ++k) t += anz(k, i) + anz(k, j);
        L += acc_chain(t);
      }
    }
  // bp_series_fwd: c update (2), accumulate switch (1), y += c u (2 per component)
  f.per_app_fwd = L + 3 + 2 * D;
  // ---- bp_series_bwd
  double B = 0;
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) if (anz(i, j)) B += (2 * n - 1) + 4;   // s2, two accumulations
  for (int l = 0; l < n; ++l) for (int q = 0; q < S; ++q) if (hnz(l, q)) B += 2;
  // bp_LT_apply
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j) { bool used = false; for (int l = 0; l < n; ++l) used = used || hnz(l, sidx(i, j)); if (used) B += 1; }
  for (int i = 0; i < n; ++i) { int t = 0; for (int j = 0; j < n; ++j) t += anz(i, j); for (int q = 0; q < S; ++q) t += hnz(i, q); B += acc_chain(t); }
  for (int i = 0; i < n; ++i)
    for (int j = i; j < n; ++j) {
      int t = 0;
      if (i == j) { for (int k = 0; k < n; ++k) t += anz(i, k); B += acc_chain(t) + (t > 0 ? 1 : 0); }
      else { for (int k = 0; k < n; ++k) t += anz(i, k) + anz(j, k); B += acc_chain(t); }
    }
  // c update (2), z = c b + t (2 per component)
  f.per_app_bwd = B + 2 + 2 * D;
  // ---- per observation outside the series (single sub-step): plan (2), 1/h (1), lp accumulation (1), z_m = c_m b (D),
  // square-root-free LDL^T factorisation with one logarithm, inverse factor, solves, cotangents, noise terms
  double fac = 0, linv = 0, quad = 0, bmu = 0, bS = 0;
  for (int j = 0; j < n; ++j) {
    fac += 3 * j + 1 + (n - 1 - j) * (2 * j + 1);
    for (int i = j + 1; i < n; ++i) linv += 2 * (i - j - 1);
    quad += 4 + 3 * j;
    bmu += 2 * (n - 1 - j);
    for (int i = 0; i <= j; ++i) bS += 2 * (n - 1 - j) + 3;
  }
  const double logdet = (n - 1) + 1, W = n * (n - 1) / 2.0;
  const double score = fac + logdet + linv + quad + 2 + bmu + W + bS + 1 + (sp.kind == 1 ? 6 * n : 0);   // incl. lp accumulation
  f.per_obs = 3 + D + score;
  // grouped path: contraction y = sum_l z0_l y^(l) (D (2n - 1)) and cotangent accumulation (2 n D) instead of the plan
  f.per_obs_grouped = score + D * (2 * n - 1) + 2 * n * D;
  // (n <= 3: bp_grouped scores through the adjugate like bp_fusedo -- 12 / 32 / 78 operations instead of the LDL^T count, plus the
  // two running sums; recounted in tests/test_kernel_core_host.py::test_octet_flop_constants_and_adjugate_score)
  static const double small_score[4] = {0.0, 12.0, 32.0, 78.0};
  if (n <= 3) f.per_obs_grouped = small_score[n] + 2 + (sp.kind == 1 ? 6.0 * n : 0.0) + D * (2 * n - 1) + 2 * n * D;
  // W path: per series term one row of the W update (2 D n); per observation the plan, the staged coefficients c_j (2 per
  // row), the packed-coordinate cotangent (S - n doublings) and the outer product b z0^T (D n)
  f.per_app_w = 2.0 * D * n;
  f.per_obs_w = 2 + score + 2.0 * (sp.wcap + 1) + (S - n) + D * n;
  // centred W path: per term of the short series the coefficient c_j(h) (2), w_j z0 (n), N_j (w_j z0) (2 D n) and one row of the
  // S update (2 D n) -- bp_fusedc accumulates these and the per-group work (N_j chain, Toeplitz product) itself; per
  // observation h (1), the score, the packed-coordinate cotangent and the outer product b z0^T.  The per-chain tables
  // (bp_ctable) and the matrix polynomial (bp_wpost) are not counted.
  f.per_term_c = 2.0 + n + 4.0 * D * n;
  f.per_obs_c = 1 + score + (S - n) + D * n;
  // octet path (bp_fusedo): the same terms; per observation h (1), the score through the adjugate (one reciprocal, no logarithm: the
  // determinants are multiplied up, bp_mvn_score_small), the running sums of q and det (2), the packed-coordinate cotangent (S - n
  // doublings), minus the coefficient update the first term does not need (2); NO outer product b z0^T -- the S' update runs on the
  // X rows the terms already count.  Recounted with a counting scalar: tests/test_kernel_core_host.py::test_octet_flop_constants.
  static const double small_obs[4] = {0.0, 13.0, 34.0, 80.0};
  f.per_obs_o = n <= 3 ? small_obs[n] + (sp.kind == 1 ? 6.0 * n : 0.0) : f.per_obs_c;
  return f;
}

}  // namespace bpc
