// flops.cpp — algorithmic operation counts of the shipped kernels (bp_kernels.cuh), by hand count of
// their loops.  Convention of SURVEY.md §8(d): add/mul = 1, FMA = 2, div/sqrt/exp/log = 1.
// tests/test_hostsim.py recounts the same device functions with a counting scalar and pins these formulas.
#include "bpcuda_internal.hpp"

namespace bpc {

FlopModel flop_model(const ModelSpec& sp) {
  const double n = sp.ntypes, S = n * (n + 1) / 2, D = n + S;
  FlopModel f;
  if (sp.kind == 2) {
    // bp_onetype_obs: value + gradient w.r.t. (b, d)
    f.per_app_fwd = f.per_app_bwd = 0.0;
    f.per_obs = 70.0;
    return f;
  }
  // bp_L_apply: mu part n(2n-1); per packed pair the H contraction (2n-1) plus the Lyapunov part
  // (diagonal 2n+1, off-diagonal 4n); bp_step_fwd adds al (1), w = al*l (D), y += w (D)
  const double L_apply = n * (2 * n - 1) + n * (2 * n - 1 + 2 * n + 1) + (S - n) * (2 * n - 1 + 4 * n);
  f.per_app_fwd = L_apply + 2 * D + 1;
  // bp_step_bwd: al, al+al (2); amu, aS (D); gA n^2 (2n+2); gH 2nS; bp_LT_apply: doubling (S-n),
  // mu part n(2n-1+2S), diagonal 2n^2, off-diagonal (S-n)(4n-1); Horner update 2D
  f.per_app_bwd = 2 + D + n * n * (2 * n + 2) + 2 * n * S + (S - n) + n * (2 * n - 1 + 2 * S) + 2 * n * n + (S - n) * (4 * n - 1) + 2 * D;
  // per observation: plan (4) and lp accumulation (1), Cholesky + inverse factor + solves + cotangents, noise terms
  double chol = 0, linv = 0, quad = 0, bmu = 0, bS = 0;
  for (int j = 0; j < sp.ntypes; ++j) {
    chol += 2 * j + 4 + (sp.ntypes - 1 - j) * (2 * j + 1);
    for (int i = j + 1; i < sp.ntypes; ++i) linv += 2 * (i - j) + 1;
    quad += 3 * (j + 1) + 2;
    bmu += 2 * (sp.ntypes - j);
    for (int i = 0; i <= j; ++i) bS += 2 * (sp.ntypes - j) + 3;
  }
  f.per_obs = 5 + chol + linv + quad + 2 + bmu + bS + (sp.kind == 1 ? 6 * n : 0);
  return f;
}

}  // namespace bpc
