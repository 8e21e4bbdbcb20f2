// codegen.cpp — model description -> CUDA translation unit (see codegen.hpp).
#include "codegen.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>

#include "rexpr.hpp"

namespace bpc {

namespace {

struct DistInfo { const char* name; int nparams; const char* constraints; };
// ALLOWED_DISTRIBUTIONS decoded from R/sysdata.rda ("name nparams constraints"; '+' = must be > 0)
const DistInfo kDists[] = {
    {"normal", 2, "_+"}, {"exp_mod_normal", 3, "_++"}, {"skew_normal", 3, "_+_"}, {"student_t", 3, "+_+"},
    {"cauchy", 2, "_+"}, {"double_exponential", 2, "_+"}, {"logistic", 2, "_+"}, {"gumbel", 2, "_+"},
    {"lognormal", 2, "_+"}, {"chi_square", 1, "+"}, {"inv_chi_square", 1, "+"}, {"scaled_inv_chi_square", 2, "++"},
    {"exponential", 1, "+"}, {"gamma", 2, "++"}, {"inv_gamma", 2, "++"}, {"weibull", 2, "++"}, {"frechet", 2, "++"},
    {"rayleigh", 1, "+"}, {"pareto", 2, "++"}, {"pareto_type_2", 3, "_++"}, {"beta", 2, "++"}, {"uniform", 2, "__"}};
constexpr int kNDists = sizeof(kDists) / sizeof(kDists[0]);

std::string lit(double v) {
  if (std::isnan(v)) return "(0.0/0.0)";
  if (std::isinf(v)) return v > 0 ? "(1.0/0.0)" : "(-1.0/0.0)";
  char b[64];
  std::snprintf(b, sizeof b, "%.17g", v);
  std::string s(b);
  if (s.find_first_of(".eEn") == std::string::npos) s += ".0";
  return s;
}

// "c0*r[e0] + c1*r[e1] ..." for a coefficient vector over events ("real(0.0)" when empty)
std::string lincomb(const std::vector<double>& coef, const char* var) {
  std::string o;
  for (size_t e = 0; e < coef.size(); ++e) {
    const double c = coef[e];
    if (c == 0.0) continue;
    char v[32];
    std::snprintf(v, sizeof v, "%s[%zu]", var, e);
    const double a = std::fabs(c);
    std::string term = (a == 1.0) ? std::string(v) : "real(" + lit(a) + ") * " + v;
    if (o.empty()) o = (c < 0 ? "-" : "") + term;
    else o += (c < 0 ? " - " : " + ") + term;
  }
  return o.empty() ? "real(0.0)" : o;
}

}  // namespace

double taylor_theta(int d) { return std::exp((-55.0 * std::log(2.0) + std::lgamma(d + 1.0)) / d); }

int prior_id(const std::string& name) {
  for (int i = 0; i < kNDists; ++i) if (name == kDists[i].name) return i;
  return -1;
}

void validate_prior(PriorSpec& p) {
  p.id = prior_id(p.name);
  if (p.id < 0) throw std::runtime_error("Invalid prior name!");
  const DistInfo& d = kDists[p.id];
  if (d.nparams != p.nparams) throw std::runtime_error("Incorrect number of parameters for prior distribution");
  for (int i = 0; i < d.nparams; ++i)
    if (p.params[i] <= 0 && d.constraints[i] == '+') throw std::runtime_error("Positive parameter required here!");
  for (int i = 0; i < d.nparams; ++i)
    if (!std::isfinite(p.params[i])) throw std::runtime_error("prior parameters must be finite numbers");
  if (p.name == "uniform" && p.params[1] <= p.params[0]) throw std::runtime_error("upper bound must be greater than lower in uniform distribution!");
  if (p.has_bounds) {
    if (std::isnan(p.lower) || std::isnan(p.upper)) throw std::runtime_error("bounds should be a list of 2 numbers!");
    if (p.lower >= p.upper) throw std::runtime_error("Error: lower bounder greater or equal to upper bound!");
    // (the reference pastes the numbers into `real<lower=%s, upper=%s>`, R/stan_utils.R:399: Inf / -Inf do not survive stanc)
    if (std::isinf(p.lower) || std::isinf(p.upper)) throw std::runtime_error("bounds must be finite numbers (use bounds = NA for an unbounded parameter)");
  }
}

std::string generate_cuda(ModelSpec& m) {
  const int n = m.ntypes, E = m.nevents, P = m.nparams;
  if (n < 1 || n > 5) throw std::runtime_error("ntypes must be between 1 and 5");
  if (E < 1 || (int)m.p_vec.size() != E || (int)m.e_mat.size() != E * n || (int)m.func_deps.size() != E)
    throw std::runtime_error("Dimensions of e_mat, p_vec, and func_deps do not agree");
  if (P <= 0) throw std::runtime_error("nparams should be an integer number of parameters");
  if (m.ndep < 0) throw std::runtime_error("ndep should be an integer number of depedent variables");
  for (int e = 0; e < E; ++e)
    if (m.p_vec[e] < 1 || m.p_vec[e] > n)
      throw std::runtime_error("p_vec must be a vector of parents for each bith event, where each entry is an positive integer corresponding to the parent type");
  for (double v : m.e_mat)
    if (!(v >= 0) || std::round(v) != v)
      throw std::runtime_error("e_mat must be a matrix of birth events, where each entry is an integer number of offspring of a given type");
  if ((int)m.priors.size() != P) throw std::runtime_error("incorrect number of priors for model!");
  if (m.kind < 0 || m.kind > 2) throw std::runtime_error("unknown template kind");
  if (m.kind == 2 && !(n == 1 && E == 2 && m.e_mat[0] == 2.0 && m.e_mat[1] == 0.0)) throw std::runtime_error("model is not a simple birth-death process!");
  if (m.kind == 1 && n == 1) throw std::runtime_error("observational noise models are not available for one-type processes");
  for (auto& p : m.priors) validate_prior(p);

  m.dim = P + (m.kind == 1 ? 1 : 0);
  m.kdep = m.ndep > 0 ? m.ndep : 1;
  m.mcap = 23;   // at most mcap + 1 = 24 applications of L per sub-step
  static const int threads_by_n[6] = {0, 256, 192, 128, 64, 32};
  m.threads = (m.kind == 2) ? 256 : threads_by_n[n];
  m.theta_cap = taylor_theta(m.mcap);
  // terms kept per thread in shared memory (ring); longer series are reversed in segments.  BPC_MSLOT / BPC_THREADS
  // override the defaults (tuning experiments).
  m.mslot = (m.mcap + 1) / 2;
  if (const char* ev = std::getenv("BPC_MSLOT")) { const int v = std::atoi(ev); if (v >= 2 && v <= m.mcap + 1) m.mslot = v; }
  if (const char* ev = std::getenv("BPC_THREADS")) { const int v = std::atoi(ev); if (m.kind != 2 && v >= 32 && v <= 256 && v % 32 == 0) m.threads = v; }
  {
    const size_t smem = (size_t)m.mslot * (n + n * (n + 1) / 2) * m.threads * 8 + 1024;
    m.minblocks = (int)std::max<size_t>(1, std::min<size_t>(8, (227 * 1024) / smem));
    while (m.minblocks > 1 && (size_t)m.minblocks * m.threads * 128 > 65536) --m.minblocks;   // never ask for < 128 registers
    if (const char* ev = std::getenv("BPC_MINBLOCKS")) { const int v = std::atoi(ev); if (v >= 1 && v <= m.minblocks) m.minblocks = v; }
  }

  // rates
  std::string rates_body, vjp_body;
  m.xdep = false;
  for (int e = 0; e < E; ++e) {
    NodePtr ast = parse_r(m.func_deps[e]);
    check_valid(*ast, P, m.ndep);
    RateProgram rp = lower_to_cuda(*ast, e, P, m.ndep);
    rates_body += "  // event " + std::to_string(e + 1) + ": " + deparse(*ast) + "\n" + rp.value_code;
    vjp_body += rp.vjp_code;
    m.xdep = m.xdep || rp.uses_x;
    for (auto& w : rp.warnings) m.warnings.push_back(w);
  }

  const int S = n * (n + 1) / 2;
  auto sidx = [n](int i, int j) { if (i > j) std::swap(i, j); return i * n - i * (i - 1) / 2 + (j - i); };
  // coefficient tensors of the generator: A[i][a] and H_l[a<=b] are linear in the rates
  std::vector<std::vector<double>> Ac(n * n, std::vector<double>(E, 0.0)), Hc(n * S, std::vector<double>(E, 0.0));
  for (int e = 0; e < E; ++e) {
    const int i = m.p_vec[e] - 1;
    for (int a = 0; a < n; ++a) {
      const double ea = m.e_mat[e + E * a];
      Ac[i * n + a][e] += ea;
      for (int b = a; b < n; ++b) Hc[i * S + sidx(a, b)][e] += ea * m.e_mat[e + E * b] - (a == b ? ea : 0.0);   // C_i
    }
    Ac[i * n + i][e] -= 1.0;
  }
  // H_l = C_l + diag(a_l) - a_l e_l^T - e_l a_l^T
  for (int l = 0; l < n; ++l)
    for (int a = 0; a < n; ++a)
      for (int b = a; b < n; ++b)
        for (int e = 0; e < E; ++e) {
          double h = 0.0;
          if (a == b) h += Ac[l * n + a][e];
          if (b == l) h -= Ac[l * n + a][e];
          if (a == l) h -= Ac[l * n + b][e];
          Hc[l * S + sidx(a, b)][e] += h;
        }

  // structural sparsity masks (bit i*n+j of A, bit l*S+q of H); all-ones for n = 5 where H would need more than 64 bits
  unsigned long long amask = 0ull, hmask = 0ull;
  m.a_nz.assign(n * n, true);
  m.h_nz.assign(n * S, true);
  if (n <= 4) {
    for (int i = 0; i < n * n; ++i) { bool nz = false; for (int e = 0; e < E; ++e) nz = nz || Ac[i][e] != 0.0; m.a_nz[i] = nz; if (nz) amask |= 1ull << i; }
    for (int i = 0; i < n * S; ++i) { bool nz = false; for (int e = 0; e < E; ++e) nz = nz || Hc[i][e] != 0.0; m.h_nz[i] = nz; if (nz) hmask |= 1ull << i; }
  } else {
    amask = ~0ull; hmask = ~0ull;
  }

  std::string s;
  char buf[768];
  s += "// generated by libbpcuda codegen for one bpinference model — do not edit\n";
  std::snprintf(buf, sizeof buf,
                "#define BP_N %d\n#define BP_E %d\n#define BP_P %d\n#define BP_DIM %d\n#define BP_KDEP %d\n#define BP_XDEP %d\n"
                "#define BP_NOISE %d\n#define BP_KIND %d\n#define BP_THREADS %d\n#define BP_MCAP %d\n#define BP_MSLOT %d\n#define BP_MINBLOCKS %d\n"
                "#define BP_THETA_CAP %s\n",
                n, E, P, m.dim, m.kdep, m.xdep ? 1 : 0, m.kind == 1 ? 1 : 0, m.kind, m.threads, m.mcap, m.mslot, m.minblocks,
                lit(m.theta_cap).c_str());
  s += buf;
  std::snprintf(buf, sizeof buf, "#define BP_AMASK 0x%llxull\n#define BP_HMASK 0x%llxull\n", amask, hmask);
  s += buf;
  {
    // W path (bp_fusedw): staging area per CTA and the CTAs per SM it allows
    const int D = n + S, wrt = (m.wcap + 1 + 7) / 8, wct = (D * n + 7) / 8;
    m.w_smem = (size_t)(m.threads / 32) * (wrt * 8 + wct * 8) * 36 * 8;
    m.w_post_threads = std::max(128, ((D * D + 31) / 32) * 32);
    int wmin = (int)std::max<size_t>(1, std::min<size_t>(3, (227 * 1024) / (m.w_smem + 1024)));
    if (const char* ev = std::getenv("BPC_WMINBLOCKS")) { const int v = std::atoi(ev); if (v >= 1 && v <= 8) wmin = v; }
    std::snprintf(buf, sizeof buf, "#define BP_WMINBLOCKS %d\n#define BP_WPOST_THREADS %d\n", wmin, m.w_post_threads);
    s += buf;
    // centred W path (bp_fusedc): per warp N_j and the T / B staging; the CTA's dense L
    const int crt = (m.wcap + 1 + 8 + 7) / 8, dn = D * n, nw = m.threads / 32;
    int csplit = 1;
    if (const char* ev = std::getenv("BPC_CSPLIT")) { const int v = std::atoi(ev); if (v >= 1 && v <= wct && wct % v == 0) csplit = v; }
    int cx = 1;   // X formulation: both products of a pass on the fp64 tensor path (bp_kernels.cuh, BP_CX)
    if (const char* ev = std::getenv("BPC_CX")) cx = std::atoi(ev) ? 1 : 0;
    const int xr = 8 * n, xdr = D > 8 ? D - 8 : 0;
    const size_t cwarp = cx ? (size_t)xr * 12 + (size_t)xr * (xdr ? xdr : 1) + (size_t)xr * 36 + (size_t)8 * 40
                            : (size_t)8 * (dn + (dn & 1)) + (size_t)(8 + wct / csplit * 8) * 36;
    std::snprintf(buf, sizeof buf, "#define BP_CX %d\n", cx);
    s += buf;
    m.c_smem = std::max<size_t>(nw * cwarp + D * D, (size_t)nw * crt * wct * 64) * 8;   // (also holds the end-of-kernel reduction of W)
    // (X formulation: per-lane accumulators of the ninth component need the registers of 3 CTAs per SM; the kernel is bound by the
    // shared-memory and FP64 pipes, not by occupancy)
    int cmin = (int)std::max<size_t>(1, std::min<size_t>(cx ? 3 : 4, (227 * 1024) / (m.c_smem + 1024)));
    std::snprintf(buf, sizeof buf, "#define BP_CSPLIT %d\n", csplit);
    s += buf;
    if (const char* ev = std::getenv("BPC_CMINBLOCKS")) { const int v = std::atoi(ev); if (v >= 1 && v <= 8) cmin = v; }
    std::snprintf(buf, sizeof buf, "#define BP_CMINBLOCKS %d\n", cmin);
    s += buf;
    // octet path (bp_fusedo): four CTAs of four warps per SM (128 registers per thread, no spills: every tile of S' has one owner warp,
    // 14 accumulator doubles per thread; tools/perf_c.py: 2.24 ms against 2.33 ms with three CTAs on cfg5U)
    int omin = 4;
    if (const char* ev = std::getenv("BPC_OMINBLOCKS")) { const int v = std::atoi(ev); if (v >= 1 && v <= 4) omin = v; }
    std::snprintf(buf, sizeof buf, "#define BP_OMINBLOCKS %d\n", omin);
    s += buf;
    int gcmin = 3;   // bp_groupc: CTAs per SM the register budget is cut for (3: 158 registers, no spills; 4: 128 registers, 128 bytes spilled)
    if (const char* ev = std::getenv("BPC_GC_MINBLOCKS")) { const int v = std::atoi(ev); if (v >= 1 && v <= 4) gcmin = v; }
    std::snprintf(buf, sizeof buf, "#define BP_GC_MINBLOCKS %d\n", gcmin);
    s += buf;
    {   // bp_fusedo2 / bp_fusedo9 are compiled only for a process that selects one of them (they add about a second to every model build)
      const char* o2 = std::getenv("BPC_O2");
      const char* o9 = std::getenv("BPC_O9");
      const char* alt = std::getenv("BPC_ALT_KERNELS");
      const bool on = (o2 && std::atoi(o2) == 1) || (o9 && std::atoi(o9) == 1) || (alt && std::atoi(alt) == 1);
      s += on ? "#define BP_ALT_KERNELS 1\n" : "#define BP_ALT_KERNELS 0\n";
    }
    int tmin = 4;    // bp_toep: CTAs per SM the register budget is cut for (n = 3: 30 accumulator doubles per thread; the (D, octets) grid of 512
                     // chains is one wave at 4; measurements at the kernel)
    if (const char* ev = std::getenv("BPC_TOEP_MINBLOCKS")) { const int v = std::atoi(ev); if (v >= 1 && v <= 16) tmin = v; }
    std::snprintf(buf, sizeof buf, "#define BP_TOEP_MINBLOCKS %d\n", tmin);
    s += buf;
  }
  s += "#ifdef BP_HOSTSIM\n#define BP_HD inline\n#define BP_GLOBAL_CONST static const\n#else\n"
       "#define BP_HD __device__ __forceinline__\n#define BP_GLOBAL_CONST __constant__ const\ntypedef double real;\n#endif\n";
  // Taylor tables
  // (tables cover the longer series of the W path: BP_WCAP >= BP_MCAP)
  std::snprintf(buf, sizeof buf, "#define BP_WCAP %d\n#define BP_WTHETA_CAP %s\n", m.wcap, lit(taylor_theta(m.wcap)).c_str());
  s += buf;
  s += "BP_GLOBAL_CONST double bp_th[BP_WCAP + 1] = {0.0";
  for (int d = 1; d <= m.wcap; ++d) s += ", " + lit(taylor_theta(d));
  s += "};\nBP_GLOBAL_CONST double bp_rk[BP_WCAP + 1] = {";
  for (int k = 0; k <= m.wcap; ++k) s += (k ? ", " : "") + lit(1.0 / (k + 1));
  s += "};\nBP_GLOBAL_CONST double bp_kp1[BP_WCAP + 1] = {";
  for (int k = 0; k <= m.wcap; ++k) s += (k ? ", " : "") + lit((double)(k + 1));
  s += "};\n";
  // prior table
  s += "BP_GLOBAL_CONST int bp_prior_id[BP_P] = {";
  for (int k = 0; k < P; ++k) s += (k ? ", " : "") + std::to_string(m.priors[k].id);
  s += "};\nBP_GLOBAL_CONST double bp_prior_par[BP_P][3] = {";
  for (int k = 0; k < P; ++k) s += std::string(k ? ", " : "") + "{" + lit(m.priors[k].params[0]) + ", " + lit(m.priors[k].params[1]) + ", " + lit(m.priors[k].params[2]) + "}";
  s += "};\nBP_GLOBAL_CONST int bp_has_bounds[BP_P] = {";
  for (int k = 0; k < P; ++k) s += (k ? ", " : "") + std::to_string(m.priors[k].has_bounds ? 1 : 0);
  s += "};\nBP_GLOBAL_CONST double bp_lo[BP_P] = {";
  for (int k = 0; k < P; ++k) s += (k ? ", " : "") + lit(m.priors[k].has_bounds ? m.priors[k].lower : 0.0);
  s += "};\nBP_GLOBAL_CONST double bp_hi[BP_P] = {";
  for (int k = 0; k < P; ++k) s += (k ? ", " : "") + lit(m.priors[k].has_bounds ? m.priors[k].upper : 0.0);
  s += "};\n\n";
  // rate functions
  s += "// rates r[e] = f_e(Theta, x): the generated `r_mat[e, p_vec[e]] = ...` lines of the reference (R/stan_utils.R:153-157)\n";
  s += "BP_HD void bp_rates(const real (&c)[BP_DIM], const real (&x)[BP_KDEP], real (&r)[BP_E]) {\n" + rates_body + "}\n";
  s += "// cb += (d r / d Theta)^T rb   (reverse mode through the same expressions)\n";
  s += "BP_HD void bp_rates_vjp(const real (&c)[BP_DIM], const real (&x)[BP_KDEP], const real (&rb)[BP_E], real (&cb)[BP_P]) {\n" + vjp_body + "}\n";
  if (m.kind != 2) {
    s += "// generator assembly (stanfiles/multitype_template.stan:23-43), linear in the rates, zero terms dropped\n";
    s += "BP_HD void bp_assemble(const real (&r)[BP_E], real (&A)[BP_N * BP_N], real (&H)[BP_N * (BP_N * (BP_N + 1) / 2)]) {\n";
    for (int i = 0; i < n * n; ++i) s += "  A[" + std::to_string(i) + "] = " + lincomb(Ac[i], "r") + ";\n";
    for (int i = 0; i < n * S; ++i) s += "  H[" + std::to_string(i) + "] = " + lincomb(Hc[i], "r") + ";\n";
    s += "}\n";
    s += "// rb[e] = <gA, dA/dr_e> + <gH, dH/dr_e>  (full-matrix inner products: packed off-diagonals count twice)\n";
    s += "BP_HD void bp_assemble_vjp(const real (&gA)[BP_N * BP_N], const real (&gH)[BP_N * (BP_N * (BP_N + 1) / 2)], real (&rb)[BP_E]) {\n";
    for (int e = 0; e < E; ++e) {
      std::string o;
      auto add = [&](double c, const std::string& v) {
        if (c == 0.0) return;
        const double a = std::fabs(c);
        std::string term = (a == 1.0) ? v : "real(" + lit(a) + ") * " + v;
        if (o.empty()) o = (c < 0 ? "-" : "") + term;
        else o += (c < 0 ? " - " : " + ") + term;
      };
      for (int i = 0; i < n * n; ++i) add(Ac[i][e], "gA[" + std::to_string(i) + "]");
      for (int l = 0; l < n; ++l)
        for (int a = 0; a < n; ++a)
          for (int b = a; b < n; ++b) add(Hc[l * S + sidx(a, b)][e] * (a == b ? 1.0 : 2.0), "gH[" + std::to_string(l * S + sidx(a, b)) + "]");
      s += "  rb[" + std::to_string(e) + "] = " + (o.empty() ? std::string("real(0.0)") : o) + ";\n";
    }
    s += "}\n";
  }
  s += "\n#include \"bp_kernels.cuh\"\n";
  return s;
}

}  // namespace bpc
