// bpcuda.cpp — host side of libbpcuda: handles, NVRTC compilation, data repacking, launches.
// The C ABI is include/bpcuda.h; each entry point there cites the reference interface it replaces.
// There is NO CPU fallback: every evaluate/sample entry point fails with BPC_E_CUDA without a device.
#include "bpcuda.h"

#include <cuda_runtime.h>
#include <nvrtc.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <string>
#include <vector>

#include "bpcuda_internal.hpp"
#include "codegen.hpp"
#include "rexpr.hpp"

extern "C" const char bp_kernels_cuh_text[];   // bp_kernels.cuh embedded at build time (embed_header.py)

namespace bpc {

thread_local std::string g_last_error;
thread_local long long g_launches = 0;

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

int cuda_fail(cudaError_t e, const char* what) {
  return fail(BPC_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

#define BPC_CUDA(call)                                      \
  do {                                                      \
    cudaError_t e__ = (call);                               \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call);   \
  } while (0)

// ---------------------------------------------------------------------------------------------
// Modules already compiled by this process, by generated source (it holds every macro of the build): creating the same model again
// (a second stan_model() of the same description, the engine a sampler test rebuilds) costs a map lookup instead of ~8 s of NVRTC.
static std::mutex g_cubin_mu;
static std::map<std::string, std::pair<std::vector<char>, std::string>> g_cubin_cache;

static int compile_model(bpc_model* m) {
  {
    std::lock_guard<std::mutex> lock(g_cubin_mu);
    auto it = g_cubin_cache.find(m->source);
    if (it != g_cubin_cache.end()) { m->cubin = it->second.first; m->compile_log = it->second.second; return BPC_OK; }
  }
  nvrtcProgram prog;
  const char* hdr_src[] = {bp_kernels_cuh_text};
  const char* hdr_name[] = {"bp_kernels.cuh"};
  if (nvrtcCreateProgram(&prog, m->source.c_str(), "bp_model.cu", 1, hdr_src, hdr_name) != NVRTC_SUCCESS)
    return fail(BPC_E_COMPILE, "nvrtcCreateProgram failed");
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "--fmad=true"};
  const nvrtcResult rc = nvrtcCompileProgram(prog, 4, opts);
  size_t logsz = 0;
  nvrtcGetProgramLogSize(prog, &logsz);
  std::string log(logsz, '\0');
  if (logsz > 1) nvrtcGetProgramLog(prog, &log[0]);
  if (rc != NVRTC_SUCCESS) {
    nvrtcDestroyProgram(&prog);
    return fail(BPC_E_COMPILE, "NVRTC compilation failed:\n" + log);
  }
  size_t sz = 0;
  if (nvrtcGetCUBINSize(prog, &sz) != NVRTC_SUCCESS || sz == 0) {
    nvrtcDestroyProgram(&prog);
    return fail(BPC_E_COMPILE, "NVRTC produced no cubin");
  }
  m->cubin.resize(sz);
  nvrtcGetCUBIN(prog, m->cubin.data());
  nvrtcDestroyProgram(&prog);
  m->compile_log = log;
  {
    std::lock_guard<std::mutex> lock(g_cubin_mu);
    if (g_cubin_cache.size() >= 64) g_cubin_cache.clear();   // (bounded: a process that builds hundreds of different models keeps the last ones)
    g_cubin_cache.emplace(m->source, std::make_pair(m->cubin, m->compile_log));
  }
  return BPC_OK;
}

int model_load(bpc_model* m, int device, Loaded** out) {
  std::lock_guard<std::mutex> lock(m->load_mu);
  auto it = m->loaded.find(device);
  if (it != m->loaded.end()) { *out = &it->second; return BPC_OK; }
  BPC_CUDA(cudaSetDevice(device));
  Loaded L;
  BPC_CUDA(cudaLibraryLoadData(&L.lib, m->cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
  BPC_CUDA(cudaLibraryGetKernel(&L.prologue, L.lib, "bp_prologue"));
  BPC_CUDA(cudaLibraryGetKernel(&L.epilogue, L.lib, "bp_epilogue"));
  BPC_CUDA(cudaLibraryGetKernel(&L.loglik, L.lib, "bp_loglik"));
  BPC_CUDA(cudaLibraryGetKernel(&L.tparams, L.lib, "bp_tparams"));
  if (m->spec.kind == BPC_SIMPLE_BD) {
    BPC_CUDA(cudaLibraryGetKernel(&L.obs, L.lib, "bp_onetype"));
    L.obs_smem = 0;
  } else {
    BPC_CUDA(cudaLibraryGetKernel(&L.obs, L.lib, "bp_fused"));
    const int n = m->spec.ntypes;
    const int D = n + n * (n + 1) / 2;
    L.obs_smem = (size_t)m->spec.mslot * D * m->spec.threads * sizeof(double);
    BPC_CUDA(cudaKernelSetAttributeForDevice(L.obs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.obs_smem, device));
    if (cudaLibraryGetKernel(&L.grouped, L.lib, "bp_grouped") != cudaSuccess) { L.grouped = nullptr; cudaGetLastError(); }
    if (cudaLibraryGetKernel(&L.groupb, L.lib, "bp_groupb") != cudaSuccess) { L.groupb = nullptr; cudaGetLastError(); }
    if (cudaLibraryGetKernel(&L.groupc, L.lib, "bp_groupc") != cudaSuccess) { L.groupc = nullptr; cudaGetLastError(); }
    BPC_CUDA(cudaLibraryGetKernel(&L.moments, L.lib, "bp_moments"));
    if (!m->spec.xdep) {
      BPC_CUDA(cudaLibraryGetKernel(&L.fusedw, L.lib, "bp_fusedw"));
      BPC_CUDA(cudaLibraryGetKernel(&L.wpost, L.lib, "bp_wpost"));
      BPC_CUDA(cudaKernelSetAttributeForDevice(L.fusedw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m->spec.w_smem, device));
      if (n <= 3) {
        BPC_CUDA(cudaLibraryGetKernel(&L.fusedc, L.lib, "bp_fusedc"));
        BPC_CUDA(cudaLibraryGetKernel(&L.ctable, L.lib, "bp_ctable"));
        BPC_CUDA(cudaKernelSetAttributeForDevice(L.fusedc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m->spec.c_smem, device));
        BPC_CUDA(cudaLibraryGetKernel(&L.otable, L.lib, "bp_otable"));
        BPC_CUDA(cudaLibraryGetKernel(&L.fusedo, L.lib, "bp_fusedo"));
        BPC_CUDA(cudaLibraryGetKernel(&L.toep, L.lib, "bp_toep"));
        // bp_fusedo: N of the octet [8 n][8 D + 4], X [2][8 n][36], Y [8 D][40], two chunks of the table [2][2 n + 1][32]  (BP_OSMEM in bp_kernels.cuh)
        L.o_smem = ((size_t)8 * n * (8 * D + 4) + 2 * (size_t)8 * n * 36 + (size_t)8 * D * 40 + 2 * (size_t)(2 * n + 1) * 32) * sizeof(double);
        BPC_CUDA(cudaKernelSetAttributeForDevice(L.fusedo, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.o_smem, device));
        const char* o9 = std::getenv("BPC_O9");
        // (measured on cfg5U: 1.991 ms per step against bp_fusedo's 1.965 -- the barriers are not what limits the passes; kept as the
        // measured alternative, BPC_O9=1 selects it)
        if (n == 3 && o9 && std::atoi(o9) == 1 && cudaLibraryGetKernel(&L.fusedo9, L.lib, "bp_fusedo9") == cudaSuccess) {
          // bp_fusedo9: N [24][76], X [2][24][36], Y tiles 0..7 [64][40], tile 8 [2][8][40], one chunk of the table [7][32]  (BP_O9SMEM)
          L.o9_smem = ((size_t)24 * 76 + 2 * 24 * 36 + 64 * 40 + 2 * 8 * 40 + 7 * 32) * sizeof(double);
          BPC_CUDA(cudaKernelSetAttributeForDevice(L.fusedo9, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.o9_smem, device));
        } else { L.fusedo9 = nullptr; cudaGetLastError(); }
        // bp_fusedo2 (warp-specialised: four tensor warps + four score warps, X and Y double-buffered): BPC_O2=1 selects it
        const char* o2 = std::getenv("BPC_O2");
        if (o2 && std::atoi(o2) == 1 && cudaLibraryGetKernel(&L.fusedo2, L.lib, "bp_fusedo2") == cudaSuccess) {
          L.o2_smem = ((size_t)8 * n * (8 * D + 4) + 4 * (size_t)8 * n * 36 + 3 * (size_t)8 * D * 40) * sizeof(double);   // BP_O2SMEM
          BPC_CUDA(cudaKernelSetAttributeForDevice(L.fusedo2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.o2_smem, device));
        } else { L.fusedo2 = nullptr; cudaGetLastError(); }
      }
    }
  }
  auto ins = m->loaded.emplace(device, L);
  *out = &ins.first->second;
  return BPC_OK;
}

}  // namespace bpc

using namespace bpc;

// =============================================================================================
namespace bpc {

// Cuts the sorted table (d->h_t, longest first) into runs of at most gmax chunks of 32 observations whose durations lie within
// cg_hcap of a point tau_g of the fine time grid.  store: build and upload the arrays of G; otherwise only count (G.n, G.maxlen).
// Returns 0, -1 when a single chunk is wider than cg_hcap (no centred path for this data), or a BPC error code (> 0).
int build_grouping(bpc_data* d, int gmax, CGrouping& G, bool store) {
  const int N = d->N, nchunks = (N + 31) / 32, FG = 64 * 64;
  const double delta = d->cdelta, hcap = d->cg_hcap;
  const std::vector<double>& ht = d->h_t;
  const ModelSpec& s = d->model->spec;
  std::vector<int> c0, idx;
  std::vector<double> tau, hm;
  auto centre = [&](double hi, double lo, int& ig, double& tg, double& hg) {
    ig = (int)std::llround(0.5 * (hi + lo) / delta);
    ig = std::max(0, std::min(FG, ig));
    tg = ig * delta;
    hg = std::max(std::fabs(hi - tg), std::fabs(lo - tg));
  };
  int ng = 0, maxlen = 0;
  for (int ch = 0; ch < nchunks;) {
    const double hi = ht[(size_t)ch * 32];
    int end = ch + 1, ig;
    double tg, hg;
    centre(hi, ht[std::min(N, end * 32) - 1], ig, tg, hg);
    if (!(hg <= hcap)) return -1;
    while (end < nchunks && end - ch < gmax) {
      int ig2;
      double tg2, hg2;
      centre(hi, ht[std::min(N, (end + 1) * 32) - 1], ig2, tg2, hg2);
      if (!(hg2 <= hcap)) break;
      ig = ig2; tg = tg2; hg = hg2; ++end;
    }
    if (store) { c0.push_back(ch); idx.push_back(ig); tau.push_back(tg); hm.push_back(hg); }
    ++ng;
    maxlen = std::max(maxlen, end - ch);
    ch = end;
  }
  G.n = ng; G.maxlen = maxlen; G.gmax = gmax;
  if (!store) return 0;
  c0.push_back(nchunks);
  auto up = [&](DevBuf& b, const void* src, size_t bytes) -> int {
    int rc = b.alloc(bytes);
    if (rc) return rc;
    if (bytes && cudaMemcpy(b.p, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess) return fail(BPC_E_CUDA, "cudaMemcpy (groups)");
    return BPC_OK;
  };
  int rc;
  if ((rc = up(G.chunk0, c0.data(), c0.size() * 4))) return rc;
  if ((rc = up(G.idx, idx.data(), idx.size() * 4))) return rc;
  if ((rc = up(G.tau, tau.data(), tau.size() * 8))) return rc;
  if ((rc = up(G.hmax, hm.data(), hm.size() * 8))) return rc;
  // c_i(tau_g) = tau_g^i / i! at index i + 8, i = 0 .. wcap + 1 (the A operand of the Toeplitz products)
  const int ctab = ((s.wcap + 1 + 8 + 7) / 8) * 8 + 16;
  std::vector<double> coef((size_t)G.n * ctab, 0.0);
  for (int gI = 0; gI < G.n; ++gI) {
    double ci = 1.0;
    for (int i = 0; i <= s.wcap + 1; ++i) {
      if (i > 0) ci = ci * (tau[gI] * (1.0 / i));
      coef[(size_t)gI * ctab + i + 8] = ci;
    }
  }
  return up(G.coef, coef.data(), coef.size() * 8);
}

}  // namespace bpc

extern "C" {

int bpc_abi_version(void) { return BPC_ABI_VERSION; }
const char* bpc_last_error(void) { return g_last_error.c_str(); }

int bpc_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

long long bpc_launch_count(int reset) {
  const long long v = g_launches;
  if (reset) g_launches = 0;
  return v;
}

// ---- expression front end --------------------------------------------------------------------
int bpc_check_expr(const char* expr, int max_params, int max_dep) {
  try {
    NodePtr n = parse_r(expr ? expr : "");
    check_valid(*n, max_params, max_dep);
    return BPC_OK;
  } catch (const std::exception& e) { return fail(BPC_E_EXPR, e.what()); }
}

static int copy_out(const std::string& s, char* out, int cap) {
  if (!out || cap <= (int)s.size()) return fail(BPC_E_INVALID, "output buffer too small");
  std::memcpy(out, s.c_str(), s.size() + 1);
  return BPC_OK;
}

int bpc_expr_to_stan(const char* expr, int max_params, int max_dep, char* out, int out_cap) {
  try {
    NodePtr n = parse_r(expr ? expr : "");
    return copy_out(to_stan(*n, max_params, max_dep), out, out_cap);
  } catch (const std::exception& e) { return fail(BPC_E_EXPR, e.what()); }
}

int bpc_expr_deparse(const char* expr, char* out, int out_cap) {
  try {
    NodePtr n = parse_r(expr ? expr : "");
    return copy_out(deparse(*n), out, out_cap);
  } catch (const std::exception& e) { return fail(BPC_E_EXPR, e.what()); }
}

static PriorSpec to_spec(const bpc_prior& p) {
  PriorSpec s;
  s.name = p.name ? p.name : "";
  s.nparams = p.nparams;
  for (int i = 0; i < 3; ++i) s.params[i] = (i < p.nparams) ? p.params[i] : 0.0;
  s.has_bounds = p.has_bounds != 0;
  s.lower = p.lower;
  s.upper = p.upper;
  return s;
}

int bpc_check_prior(const bpc_prior* prior) {
  if (!prior) return fail(BPC_E_INVALID, "null prior");
  if (prior->nparams < 0 || prior->nparams > 3) return fail(BPC_E_PRIOR, prior_id(prior->name ? prior->name : "") < 0 ? "Invalid prior name!" : "Incorrect number of parameters for prior distribution");
  try {
    PriorSpec s = to_spec(*prior);
    validate_prior(s);
    return BPC_OK;
  } catch (const std::exception& e) { return fail(BPC_E_PRIOR, e.what()); }
}

// ---- model -------------------------------------------------------------------------------------
int bpc_model_create(const bpc_model_desc* d, bpc_model** out) {
  if (!d || !out) return fail(BPC_E_INVALID, "null argument");
  *out = nullptr;
  auto m = std::make_unique<bpc_model>();
  ModelSpec& s = m->spec;
  s.ntypes = d->ntypes; s.nevents = d->nevents; s.nparams = d->nparams; s.ndep = d->ndep; s.kind = d->kind;
  if (d->ntypes < 1 || d->nevents < 1 || !d->e_mat || !d->p_vec || !d->func_deps) return fail(BPC_E_INVALID, "Dimensions of e_mat, p_vec, and func_deps do not agree");
  if (d->nparams <= 0) return fail(BPC_E_INVALID, "nparams should be an integer number of parameters");
  if (!d->priors) return fail(BPC_E_INVALID, "incorrect number of priors for model!");
  s.e_mat.assign(d->e_mat, d->e_mat + (size_t)d->nevents * d->ntypes);
  s.p_vec.assign(d->p_vec, d->p_vec + d->nevents);
  for (int e = 0; e < d->nevents; ++e) s.func_deps.emplace_back(d->func_deps[e] ? d->func_deps[e] : "");
  for (int k = 0; k < d->nparams; ++k) {
    if (d->priors[k].nparams < 0 || d->priors[k].nparams > 3) return fail(BPC_E_PRIOR, "Incorrect number of parameters for prior distribution");
    s.priors.push_back(to_spec(d->priors[k]));
  }
  try {
    m->source = generate_cuda(s);
  } catch (const ExprError& e) {
    return fail(BPC_E_EXPR, e.what());
  } catch (const std::exception& e) {
    const std::string w = e.what();
    const bool prior = w.find("prior") != std::string::npos || w.find("Positive parameter") != std::string::npos || w.find("bound") != std::string::npos;
    return fail(prior ? BPC_E_PRIOR : BPC_E_INVALID, w);
  }
  for (auto& w : s.warnings) m->warnings += w + "\n";
  const int rc = compile_model(m.get());
  if (rc) return rc;
  *out = m.release();
  return BPC_OK;
}

void bpc_model_free(bpc_model* m) {
  if (!m) return;
  for (auto& kv : m->loaded) {
    cudaSetDevice(kv.first);
    cudaLibraryUnload(kv.second.lib);
  }
  delete m;
}

int bpc_model_dim(const bpc_model* m) { return m ? m->spec.dim : -1; }
const char* bpc_model_cuda_source(const bpc_model* m) { return m ? m->source.c_str() : ""; }
const char* bpc_model_warnings(const bpc_model* m) { return m ? m->warnings.c_str() : ""; }
int bpc_model_cubin(const bpc_model* m, const void** data, unsigned long long* size) {
  if (!m || !data || !size) return fail(BPC_E_INVALID, "null argument");
  *data = m->cubin.data();
  *size = m->cubin.size();
  return BPC_OK;
}

// ---- constrain / unconstrain (host arithmetic on a handful of numbers; not the hot path) ---------
int bpc_constrain(const bpc_model* m, const double* u, int nchains, double* theta) {
  if (!m || !u || !theta || nchains < 0) return fail(BPC_E_INVALID, "null argument");
  const ModelSpec& s = m->spec;
  for (int c = 0; c < nchains; ++c) {
    for (int k = 0; k < s.nparams; ++k) {
      const double uk = u[(size_t)c * s.dim + k];
      const PriorSpec& p = s.priors[k];
      double th = uk;
      if (p.has_bounds) {
        const double sg = uk >= 0 ? 1.0 / (1.0 + std::exp(-uk)) : std::exp(uk) / (1.0 + std::exp(uk));
        th = p.lower + (p.upper - p.lower) * sg;
      }
      theta[(size_t)c * s.dim + k] = th;
    }
    if (s.kind == BPC_MULTITYPE_NOISE) theta[(size_t)c * s.dim + s.nparams] = std::exp(u[(size_t)c * s.dim + s.nparams]);
  }
  return BPC_OK;
}

int bpc_unconstrain(const bpc_model* m, const double* theta, int nchains, double* u) {
  if (!m || !u || !theta || nchains < 0) return fail(BPC_E_INVALID, "null argument");
  const ModelSpec& s = m->spec;
  for (int c = 0; c < nchains; ++c) {
    for (int k = 0; k < s.nparams; ++k) {
      const double th = theta[(size_t)c * s.dim + k];
      const PriorSpec& p = s.priors[k];
      double uk = th;
      if (p.has_bounds) {
        const double z = (th - p.lower) / (p.upper - p.lower);
        uk = std::log(z / (1.0 - z));
      }
      u[(size_t)c * s.dim + k] = uk;
    }
    if (s.kind == BPC_MULTITYPE_NOISE) u[(size_t)c * s.dim + s.nparams] = std::log(theta[(size_t)c * s.dim + s.nparams]);
  }
  return BPC_OK;
}

// ---- data ----------------------------------------------------------------------------------------
int bpc_data_create(const bpc_model* m, const bpc_stan_data* d, int device, int grouping, bpc_data** out) {
  if (!m || !d || !out) return fail(BPC_E_INVALID, "null argument");
  *out = nullptr;
  const ModelSpec& s = m->spec;
  const int N = d->ndatapts, n = s.ntypes;
  const bool simple = s.kind == BPC_SIMPLE_BD;
  if (N < 0 || !d->var_idx || (N > 0 && (!d->pop_vec || !d->init_pop || !d->times)))
    return fail(BPC_E_INVALID, "times, initial populations, and final populations must all have same number of rows!");
  if (d->ndep_levels < 1 || d->ndep < 1 || !d->function_var) return fail(BPC_E_INVALID, "function_var must have at least one level and one column (the dummy matrix(0,1,1) when ndep == 0)");
  if (s.ndep > 0 && d->ndep < s.ndep) return fail(BPC_E_INVALID, "c_mat does not contain enough dependent variables for the model");
  if (!simple && (!d->times_idx || d->ntimes_unique < 1)) return fail(BPC_E_INVALID, "times_idx / ntimes_unique missing for a multitype data list");
  for (int k = 0; k < N; ++k) {
    if (d->var_idx[k] < 1 || d->var_idx[k] > d->ndep_levels) return fail(BPC_E_INVALID, "var_idx out of range [1, ndep_levels]");
    if (!simple && (d->times_idx[k] < 1 || d->times_idx[k] > d->ntimes_unique)) return fail(BPC_E_INVALID, "times_idx out of range [1, ntimes_unique]");
  }
  // durations must be finite and >= 0 (they are sorted below: a NaN would break the ordering), populations and dependent variables finite
  for (int k = 0, nt = simple ? N : d->ntimes_unique; k < nt; ++k)
    if (!std::isfinite(d->times[k]) || d->times[k] < 0.0) return fail(BPC_E_INVALID, "times must be finite and non-negative");
  for (size_t k = 0, nn = (size_t)N * (simple ? 1 : n); k < nn; ++k)
    if (!std::isfinite(d->pop_vec[k]) || !std::isfinite(d->init_pop[k])) return fail(BPC_E_INVALID, "pop_vec and init_pop must be finite");
  for (size_t k = 0, nn = (size_t)d->ndep_levels * d->ndep; k < nn; ++k)
    if (!std::isfinite(d->function_var[k])) return fail(BPC_E_INVALID, "function_var must be finite");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(BPC_E_CUDA, "no CUDA device available: libbpcuda has no CPU fallback"); }
  if (device < 0 || device >= ndev) return fail(BPC_E_INVALID, "device index out of range");
  BPC_CUDA(cudaSetDevice(device));

  auto dd = std::make_unique<bpc_data>();
  dd->model = m;
  dd->device = device;
  dd->N = N;
  // per-observation duration and (level, duration) group
  std::vector<double> t(N);
  std::vector<long long> gkey(N);
  for (int k = 0; k < N; ++k) {
    t[k] = simple ? d->times[k] : d->times[d->times_idx[k] - 1];
    gkey[k] = simple ? k : (long long)(d->var_idx[k] - 1) * d->ntimes_unique + (d->times_idx[k] - 1);
  }
  {
    std::vector<long long> u(gkey);
    std::sort(u.begin(), u.end());
    dd->ngroups = (int)(std::unique(u.begin(), u.end()) - u.begin());
  }
  // sort: longest duration first (similar series length within a warp; heavy tiles launch first)
  std::vector<int> perm(N);
  std::iota(perm.begin(), perm.end(), 0);
  // (ties broken by group so that the observations of one (level, duration) pair are contiguous: the grouped kernel's ranges)
  std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return t[a] != t[b] ? t[a] > t[b] : gkey[a] < gkey[b]; });
  const int npad = ((N + 31) / 32) * 32;
  dd->npad = npad;
  const int kd = s.kdep;
  std::vector<double> hz((size_t)n * npad, 1.0), hx((size_t)n * npad, 1.0), ht(npad, 0.0), hv((size_t)kd * npad, 0.0);
  std::vector<int> hp(npad, 0);
  for (int j = 0; j < N; ++j) {
    const int k = perm[j];
    for (int i = 0; i < n; ++i) {
      hz[(size_t)i * npad + j] = d->init_pop[k + (size_t)N * i];
      hx[(size_t)i * npad + j] = d->pop_vec[k + (size_t)N * i];
    }
    ht[j] = t[k];
    const int v = d->var_idx[k] - 1;
    for (int q = 0; q < kd; ++q) hv[(size_t)q * npad + j] = (q < d->ndep) ? d->function_var[v + (size_t)d->ndep_levels * q] : 0.0;
    hp[j] = k;
  }
  auto up = [&](DevBuf& b, const void* src, size_t bytes) -> int {
    int rc = b.alloc(bytes);
    if (rc) return rc;
    if (bytes) BPC_CUDA(cudaMemcpy(b.p, src, bytes, cudaMemcpyHostToDevice));
    return BPC_OK;
  };
  int rc;
  if ((rc = up(dd->z0, hz.data(), hz.size() * 8))) return rc;
  if ((rc = up(dd->xo, hx.data(), hx.size() * 8))) return rc;
  if ((rc = up(dd->t, ht.data(), ht.size() * 8))) return rc;
  if ((rc = up(dd->xv, hv.data(), hv.size() * 8))) return rc;
  if ((rc = up(dd->perm, hp.data(), hp.size() * 4))) return rc;
  if ((rc = up(dd->fvar, d->function_var, (size_t)d->ndep_levels * d->ndep * 8))) return rc;
  dd->nlev = d->ndep_levels;
  dd->ncols = d->ndep;
  // groups = runs of equal (level, duration) in the sorted order
  dd->grouped = false;
  dd->wmode = false;
  if (!simple && N > 0) {
    std::vector<int> gs;
    std::vector<double> gtv;
    std::vector<int> glev;
    for (int j = 0; j < N; ++j)
      if (j == 0 || gkey[perm[j]] != gkey[perm[j - 1]]) { gs.push_back(j); gtv.push_back(t[perm[j]]); glev.push_back(d->var_idx[perm[j]] - 1); }
    const int G = (int)gs.size();
    gs.push_back(N);
    const long long GN = (long long)G * n;
    const int D = n + n * (n + 1) / 2;
    const size_t smem = (size_t)(2 * GN * D + 4 * std::max(n * D, s.nparams + 3) + 3 * D * D + (2 * (s.wcap + 1) + 1) * D * n) * 8 + (size_t)G * 4 + 16;
    const bool fits = GN <= 4096 && smem <= 200 * 1024;
    if (grouping == 2 && !fits) return fail(BPC_E_UNSUPPORTED, "too many (level, duration) groups for the grouped kernel");
    if (fits && (grouping == 2 || (grouping == 0 && GN * 8 <= N))) {
      std::vector<double> gxv((size_t)kd * G, 0.0);
      for (int g = 0; g < G; ++g)
        for (int q = 0; q < kd; ++q) gxv[(size_t)q * G + g] = (q < d->ndep) ? d->function_var[glev[g] + (size_t)d->ndep_levels * q] : 0.0;
      if ((rc = up(dd->gstart, gs.data(), gs.size() * 4))) return rc;
      if ((rc = up(dd->gt, gtv.data(), gtv.size() * 8))) return rc;
      if ((rc = up(dd->gx, gxv.data(), gxv.size() * 8))) return rc;
      // compact level ids (groups of one level share the generator)
      std::vector<int> lev_ids, gl(G);
      for (int g = 0; g < G; ++g) {
        int id = -1;
        for (size_t q = 0; q < lev_ids.size(); ++q) if (lev_ids[q] == glev[g]) id = (int)q;
        if (id < 0) { id = (int)lev_ids.size(); lev_ids.push_back(glev[g]); }
        gl[g] = id;
      }
      if ((rc = up(dd->glev, gl.data(), gl.size() * 4))) return rc;
      dd->nlevels = (int)lev_ids.size();
      dd->grouped = true;
      dd->ngroups = G;
      dd->group_threads = N <= 2048 ? 32 : (N <= 8192 ? 64 : 128);
      // many observations per chain: phase B on its own grid (bp_groupb), tiles of <= 4096 observations that do not straddle a group
      // (BPC_GROUPB=0 keeps the single launch; small data, which is what a sampler pass on the reference's examples is made of, too)
      const char* gb = std::getenv("BPC_GROUPB");
      if (N >= 16384 && !(gb && std::atoi(gb) == 0)) {
        std::vector<int> tg, ts, gt0;
        int tsz = 4096;   // (cfg5G, 512 chains x 100 000 observations: 512 -> 1.29 ms, 1024 -> 1.03, 2048 -> 0.90, 4096 -> 0.85)
        if (const char* ev = std::getenv("BPC_GROUPB_TILE")) { const int v = std::atoi(ev); if (v >= 128 && v <= (1 << 20)) tsz = v; }
        for (int g = 0; g < G; ++g) {
          gt0.push_back((int)tg.size());
          for (int k = gs[g]; k < gs[g + 1]; k += tsz) { tg.push_back(g); ts.push_back(k); }
        }
        gt0.push_back((int)tg.size());
        ts.push_back(N);
        dd->ntb = (int)tg.size();
        if ((rc = up(dd->tgroup, tg.data(), tg.size() * 4))) return rc;
        if ((rc = up(dd->tstart, ts.data(), ts.size() * 4))) return rc;
        if ((rc = up(dd->gtile0, gt0.data(), gt0.size() * 4))) return rc;
        // the same for bp_groupc (one thread per chain, from 128 chains): tiles of <= 256 observations (BP_GC_TILE)
        tg.clear(); ts.clear(); gt0.clear();
        int tsc = 256;
        if (const char* ev = std::getenv("BPC_GROUPC_TILE")) { const int v = std::atoi(ev); if (v >= 32 && v <= 512) tsc = v; }
        for (int g = 0; g < G; ++g) {
          gt0.push_back((int)tg.size());
          for (int k = gs[g]; k < gs[g + 1]; k += tsc) { tg.push_back(g); ts.push_back(k); }
        }
        gt0.push_back((int)tg.size());
        ts.push_back(N);
        dd->ntc = (int)tg.size();
        if ((rc = up(dd->tgroup_c, tg.data(), tg.size() * 4))) return rc;
        if ((rc = up(dd->tstart_c, ts.data(), ts.size() * 4))) return rc;
        if ((rc = up(dd->gtile0_c, gt0.data(), gt0.size() * 4))) return rc;
      }
    }
  }
  // W path (bp_fusedw + bp_wpost): rates independent of per-observation variables, ungrouped data, n <= 3
  dd->wmode = !simple && N > 0 && !dd->grouped && !s.xdep && n <= 3 && !std::getenv("BPC_NO_WPATH");
  // centred W path (bp_ctable + bp_fusedc): the sorted table is cut into groups of consecutive 32-observation chunks whose
  // durations lie within hcap of a point tau_g of the fine time grid (spacing tmax / 4096).  hcap is chosen so that the short
  // series in h = t - tau_g has degree at most BP_CJ - 1 = 7 for EVERY chain the W path accepts
  // (2 ||A|| tmax <= theta(wcap)  =>  2 ||A|| hcap <= theta(7)).  Data whose chunks are wider than that (few observations spread
  // over a long range) stay on bp_fusedw.
  if (dd->wmode && ht[0] > 0.0 && !std::getenv("BPC_NO_CPATH")) {
    const double tmax = ht[0];
    const int FG = 64 * 64;
    const double delta = tmax / FG;
    const double hcap = 0.999 * tmax * taylor_theta(7) / taylor_theta(s.wcap);
    // runs of at most gmax chunks: 16 for bp_fusedc (tools/perf_c.py: 8 -> 3.03 ms, 12 -> 3.13, 16 -> 2.78, 20 -> 2.99 on cfg5U; wider groups
    // need a longer short series); the octet path, whose passes always carry all BP_CJ terms, takes wider ones (8 -> 2.37 ms, 16 -> 2.23,
    // 24 -> 2.17, 32 -> 2.17: bp_fusedo's time does not depend on the width, bp_toep's is proportional to the number of groups; a search
    // for the width that fills the last wave of CTAs best found nothing to gain).  choose_tiling picks the grouping by chain count.
    dd->h_t.assign(ht.begin(), ht.begin() + N);
    dd->cg_hcap = hcap;
    dd->cdelta = delta;
    dd->cg_fixed = 0;
    if (const char* ev = std::getenv("BPC_CGROUP_CHUNKS")) { const int v = std::atoi(ev); if (v >= 1 && v <= 64) dd->cg_fixed = v; }
    int rcg = bpc::build_grouping(dd.get(), dd->cg_fixed ? dd->cg_fixed : 16, dd->cgs[0], true);
    if (rcg > 0) return rcg;
    if (rcg == 0) rcg = bpc::build_grouping(dd.get(), dd->cg_fixed ? dd->cg_fixed : 32, dd->cgs[1], true);
    if (rcg > 0) return rcg;
    if (rcg == 0) {
      dd->cgroups = dd->cgs[0].n;
      dd->cmode = true;
    }
    std::vector<double>().swap(dd->h_t);
  }
  BPC_CUDA(cudaStreamCreateWithFlags(&dd->stream, cudaStreamNonBlocking));
  *out = dd.release();
  return BPC_OK;
}

void bpc_data_free(bpc_data* d) {
  if (!d) return;
  cudaSetDevice(d->device);
  for (auto& ev : d->timing_events) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
  if (d->ws_event) cudaEventDestroy(d->ws_event);
  if (d->stream) cudaStreamDestroy(d->stream);
  delete d;
}

int bpc_data_ngroups(const bpc_data* d) { return d ? d->ngroups : -1; }
int bpc_data_is_grouped(const bpc_data* d) { return d ? (d->grouped ? 1 : 0) : -1; }

}  // extern "C"

// ---- evaluate ------------------------------------------------------------------------------------
namespace bpc {

// tile (observations per CTA) so that the grid has a few waves of CTAs when the problem allows it
static void choose_tiling(bpc_data* d, int nchains) {
  if (d->grouped) { d->tile = d->N; d->ntiles = 1; return; }
  const int T = d->model->spec.threads;
  // octet path: eight chains per tensor tile once there are at least two octets (BPC_OPATH=0 keeps bp_fusedc, =1 forces the octet path from 8 chains)
  { const char* ev = std::getenv("BPC_OPATH"); d->omode = d->cmode && nchains >= 8 && (ev ? std::atoi(ev) != 0 : nchains >= 16); }
  if (d->cmode) { d->cgsel = d->omode ? 1 : 0; d->cgroups = d->cgs[d->cgsel].n; }
  if (d->omode) {
    // octet path: CTA = cgpc consecutive groups of one octet of chains; one group per CTA unless that makes more than 16 waves of
    // 4 CTAs per SM (tools/perf_c.py on cfg5U: 1 -> 2.215 ms, 2 -> 2.230, 3 -> 2.236, 4 -> 2.256)
    d->noct = (nchains + 7) / 8;
    long long gpc = ((long long)d->cgroups * d->noct) / (148LL * 4 * 16);
    long long cap = 1;
    if (const char* ev = std::getenv("BPC_OGROUPS_PER_CTA")) { const long long v = std::atoll(ev); if (v >= 1 && v <= 4096) cap = v; }
    gpc = std::max<long long>(1, std::min(cap, gpc));
    d->cgpc = (int)gpc;
    d->ntiles = std::max(1, (d->cgroups + d->cgpc - 1) / d->cgpc);
    d->tile = ((d->N + d->ntiles - 1) / d->ntiles + 31) / 32 * 32;   // bp_fused's tiles (chains that need sub-steps) over the same ntiles
    return;
  }
  if (d->cmode) {
    // CTA = cgpc consecutive groups (a multiple of the warps per CTA); aim at >= 4 waves of 3 CTAs per SM
    const int nw = T / 32;
    long long gpc = ((long long)d->cgroups * nchains) / (148LL * 3 * 4);
    long long cap = 8LL * nw;
    if (const char* ev = std::getenv("BPC_CGROUPS_PER_CTA")) { const long long v = std::atoll(ev); if (v >= 1 && v <= 4096) cap = v; }
    gpc = std::max<long long>(nw, std::min(cap, gpc));
    gpc = (gpc / nw) * nw;
    d->cgpc = (int)gpc;
    d->ntiles = std::max(1, (d->cgroups + d->cgpc - 1) / d->cgpc);
    d->tile = ((d->N + d->ntiles - 1) / d->ntiles + 31) / 32 * 32;   // bp_fused's tiles (chains that need sub-steps) over the same ntiles
    return;
  }
  if (d->model->spec.kind == BPC_SIMPLE_BD && d->N <= 8 * T) {   // small data: one tile per chain, and the kernel does its own input / output
    d->tile = ((d->N + T - 1) / T) * T;
    d->ntiles = 1;
    return;
  }
  const long long total = (long long)d->N * nchains;
  long long per_thread = total / ((long long)148 * 8 * T);
  long long cap = d->wmode ? 32 : 8;
  if (const char* ev = std::getenv("BPC_OBS_PER_THREAD")) { const long long v = std::atoll(ev); if (v >= 1 && v <= 1024) cap = v; }
  per_thread = std::max(1LL, std::min(cap, per_thread));
  d->tile = (int)(T * per_thread);
  d->ntiles = std::max(1, (d->N + d->tile - 1) / d->tile);
}

static int ensure_workspace(bpc_data* d, int nchains) {
  const ModelSpec& s = d->model->spec;
  if (nchains != d->ws_chains) {
    choose_tiling(d, nchains);
    int rc;
    if ((rc = d->theta.alloc((size_t)nchains * s.dim * 8))) return rc;
    if ((rc = d->part.alloc((size_t)nchains * d->ntiles * (s.nparams + 3) * 8))) return rc;
    if ((rc = d->apps.alloc((size_t)nchains * 8))) return rc;
    if (d->grouped) {
      const int n = s.ntypes, D = n + n * (n + 1) / 2;
      if ((rc = d->ring.alloc((size_t)nchains * s.mslot * D * d->ngroups * n * 8))) return rc;
      if (d->ntb > 0) {
        if ((rc = d->gmom.alloc((size_t)nchains * d->ngroups * n * D * 8))) return rc;
        if ((rc = d->ypart.alloc((size_t)nchains * std::max(d->ntb, d->ntc) * (n * D + 2) * 8))) return rc;
      }
    }
    if (d->wmode) {
      const int n = s.ntypes, D = n + n * (n + 1) / 2;
      const size_t wfrag = (size_t)((s.wcap + 1 + (d->cmode ? 8 : 0) + 7) / 8) * ((D * n + 7) / 8) * 64;
      if (!d->omode && (rc = d->wpart.alloc((size_t)nchains * d->ntiles * wfrag * 8))) return rc;
      if (d->cmode && (rc = d->fb.alloc((size_t)(nchains + 1) * 4))) return rc;
      if (d->omode) {
        const size_t oct = (size_t)d->noct * d->cgroups * (8 * n) * (8 * D) * 8;
        if ((rc = d->osg.alloc(oct))) return rc;
        if ((rc = d->ctab_l.alloc((size_t)d->noct * 8 * (D * D + 1) * 8))) return rc;
        if ((rc = d->cg_jm.alloc((size_t)d->noct * 8 * d->cgroups * 4))) return rc;
        if ((rc = d->cg_flops.alloc((size_t)d->noct * 8 * d->cgroups * 8))) return rc;
        if ((rc = d->wplain.alloc((size_t)nchains * (((s.wcap + 1 + 8 + 7) / 8) * 8) * D * n * 8))) return rc;
        if ((rc = d->n0tab.alloc((size_t)nchains * d->cgroups * D * n * 8))) return rc;
      } else if (d->cmode) {
        if ((rc = d->csg.alloc((size_t)nchains * d->cgroups * 8 * ((D * n + 7) / 8) * 8 * 8))) return rc;   // one 8-row S matrix per group
        if ((rc = d->ctab_m1.alloc((size_t)nchains * 65 * D * n * 8))) return rc;
        if ((rc = d->ctab_e2.alloc((size_t)nchains * 64 * D * D * 8))) return rc;
      }
    }
    d->ws_chains = nchains;
    ++d->ws_gen;
  }
  return BPC_OK;
}

// The workspace of a data handle is shared by every entry point that takes the handle.  Every use ends by recording ws_event on
// the stream it ran on (release_workspace) and begins by making its stream wait for that event (order_workspace): sequential use from
// one thread on different streams is ordered, also when the earlier stream has been destroyed in the meantime (an event outlives
// it).  Handles are single-owner and not thread-safe (include/bpcuda.h).  Neither is called while the stream is being captured.
int order_workspace(bpc_data* d, cudaStream_t st) {
  if (d->ws_event) BPC_CUDA(cudaStreamWaitEvent(st, d->ws_event, 0));
  return BPC_OK;
}

int release_workspace(bpc_data* d, cudaStream_t st) {
  if (!d->ws_event) BPC_CUDA(cudaEventCreateWithFlags(&d->ws_event, cudaEventDisableTiming));
  BPC_CUDA(cudaEventRecord(d->ws_event, st));
  return BPC_OK;
}

struct GroupArgs {   // must match BpGroupArgs in bp_kernels.cuh
  double* theta; const double* z0; const double* xo; const int* gstart; const double* gt; const double* gx; const int* glev;
  double* ring; double* part; int* status;
  int nobs, npad, ngroups, nchains, nlevels;
  const double* u; int jacobian; double* lp_out; double* grad_out; double* apps_out;
  int yscap;
  const int* skip;
  int mode, ntb; const int* tgroup; const int* tstart; const int* gtile0; double* gmom; double* ypart; int ypt;
};

struct ObsArgs {   // must match BpObsArgs in bp_kernels.cuh
  double* theta; const double* z0; const double* xo; const double* t; const double* xv;
  double* part; int* status; double* loglik; const int* perm;
  int nobs, npad, tile, ntiles, nchains;
  double* wpart; int wmode;
  int cgroups, cgpc; const int* cg_chunk0; const int* cg_idx; const double* cg_tau; const double* cg_hmax; const double* cg_coef; double cdelta;
  double* ctab_m1; double* ctab_e2; double* csg;
  int noct; int* cg_jm; double* cg_flops; double* osg; double* wplain;
  int* fb; double* ctab_l;
  const double* u; int jacobian; double* lp_out; double* grad_out; double* apps_out;
  double* n0tab;
  const int* skip;
  int operm;
};

int log_prob_dev(bpc_model* m, bpc_data* d, const double* u_dev, int nchains, int jacobian, double* lp_dev,
                 double* grad_dev, int* status_dev, cudaStream_t st, const int* skip_dev) {
  if (nchains <= 0) return BPC_OK;
  Loaded* L;
  int rc = model_load(m, d->device, &L);
  if (rc) return rc;
  BPC_CUDA(cudaSetDevice(d->device));
  if ((rc = ensure_workspace(d, nchains))) return rc;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(st, &cap);
  if (cap == cudaStreamCaptureStatusNone && (rc = order_workspace(d, st))) return rc;
  const ModelSpec& s = m->spec;
  double* theta = (double*)d->theta.p;
  double* part = (double*)d->part.p;
  double* apps = (double*)d->apps.p;
  // grouped path, and the one-type template with one tile per chain: one CTA per chain does the transform, the evaluation and the
  // priors / chain rule itself (one launch)
  const bool own_io_1t = s.kind == BPC_SIMPLE_BD && d->ntiles == 1;
  const bool own_io = (d->grouped && L->grouped) || own_io_1t;
  const bool own_in = !own_io && d->wmode && d->omode;   // octet path: bp_otable does the transform
  // programmatic dependent launch between the kernels of one evaluation (MEASURED: cfg5U 1.967 -> 1.962 ms per step, cfg5G 0.637 ->
  // 0.633; a single-launch evaluation has nothing to gain, and replayed from the sampler's CUDA graph the programmatic edges made the
  // cfg2 passes 14 % SLOWER, so captured launches stay ordinary)
  const bool pdl = cap == cudaStreamCaptureStatusNone && pdl_enabled() && (!own_io || (d->grouped && d->ntb > 0));
  // BPC_KTIME=1 (diagnostic): an event pair around every launch of this evaluation, the times of the kernels in launch order on stderr
  // after a synchronize (warm numbers of the step's kernels in their real sequence, next to the cold per-kernel times ncu reports)
  static const bool ktime = [] { const char* e = std::getenv("BPC_KTIME"); return e && std::atoi(e) == 1; }();
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> kev;
  const bool kt = ktime && cap == cudaStreamCaptureStatusNone;
  auto launch = [pdl, kt, &kev](const void* f, dim3 grid, dim3 block, void** args, size_t smem, cudaStream_t stream) {
    cudaEvent_t a = nullptr, b = nullptr;
    if (kt) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, stream); }
    const cudaError_t e = pdl ? pdl_launch(f, grid, block, args, smem, stream) : cudaLaunchKernel(f, grid, block, args, smem, stream);
    if (kt) { cudaEventRecord(b, stream); kev.emplace_back(a, b); }
    return e;
  };
  if (!own_io && !own_in) {
    void* args[] = {(void*)&u_dev, (void*)&nchains, (void*)&theta, (void*)&status_dev};
    BPC_CUDA(launch((const void*)L->prologue, dim3((nchains + 127) / 128), dim3(128), args, 0, st));
  }
  {
    ObsArgs a{theta, (const double*)d->z0.p, (const double*)d->xo.p, (const double*)d->t.p, (const double*)d->xv.p,
              part, status_dev, nullptr, (const int*)d->perm.p, d->N, d->npad, d->tile, d->ntiles, nchains,
              (double*)d->wpart.p, d->wmode ? (d->omode ? 3 : (d->cmode ? 2 : 1)) : 0,
              d->cgroups, d->cgpc, (const int*)d->cgs[d->cgsel].chunk0.p, (const int*)d->cgs[d->cgsel].idx.p, (const double*)d->cgs[d->cgsel].tau.p,
              (const double*)d->cgs[d->cgsel].hmax.p, (const double*)d->cgs[d->cgsel].coef.p, d->cdelta, (double*)d->ctab_m1.p, (double*)d->ctab_e2.p, (double*)d->csg.p,
              d->noct, (int*)d->cg_jm.p, (double*)d->cg_flops.p, (double*)d->osg.p, (double*)d->wplain.p,
              (int*)d->fb.p, (double*)d->ctab_l.p,
              (own_io_1t || own_in) ? u_dev : nullptr, jacobian, lp_dev, grad_dev, apps, (double*)d->n0tab.p, skip_dev, 0};
    void* args[] = {(void*)&a};
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (d->timing) {
      BPC_CUDA(cudaEventCreate(&e0));
      BPC_CUDA(cudaEventCreate(&e1));
      BPC_CUDA(cudaEventRecord(e0, st));
    }
    if (d->grouped) {
      if (!L->grouped) return fail(BPC_E_UNSUPPORTED, "module has no grouped kernel");
      const int n = s.ntypes, D = n + n * (n + 1) / 2;
      // CTA size: as many threads as keep all chains resident in one wave (the kernel needs ~240 registers per thread),
      // never more than the observations can use
      int T = d->group_threads;
      while (T < 128 && (long long)nchains * 2 <= 148LL * (65536 / (256 * T * 2)) * 2 && d->N > 2 * T) T *= 2;
      const size_t wrows = s.wcap + 1;
      // (layout of bp_grouped; n <= 3: the cooperative multi-step path with D x D work arrays and yscap blocks for the sub-step states)
      const bool gms = n <= 3;
      const size_t base0 = (size_t)(2 * d->ngroups * n * D + (T / 32) * std::max(n * D, s.nparams + 3) + 3 * D * D + (2 * wrows + 1) * D * n) * 8 +
                           (size_t)d->ngroups * 12 + 16;
      const size_t extra = (size_t)((2 * wrows + 1) * D * (D - n) + 2 * D * D + 2 * D * n) * 8;   // D x D work arrays instead of D x n, F, Q, z
      // sub-step states: 4 G + 12 blocks when everything stays within 64 KB per CTA (data with many groups: the path is left out)
      int yscap = 0;
      if (gms && base0 + extra + (size_t)(2 * d->ngroups + 2) * D * n * 8 <= 64 * 1024)
        yscap = (int)std::min<size_t>(4 * (size_t)d->ngroups + 12, (64 * 1024 - base0 - extra) / ((size_t)D * n * 8));
      const size_t smem = base0 + (yscap > 0 ? extra + (size_t)yscap * D * n * 8 : 0);
      if (smem > L->grouped_smem) {
        BPC_CUDA(cudaKernelSetAttributeForDevice(L->grouped, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem, d->device));
        L->grouped_smem = smem;
      }
      GroupArgs ga{theta, (const double*)d->z0.p, (const double*)d->xo.p, (const int*)d->gstart.p, (const double*)d->gt.p,
                   (const double*)d->gx.p, (const int*)d->glev.p, (double*)d->ring.p, part, status_dev, d->N, d->npad, d->ngroups, nchains,
                   d->nlevels, u_dev, jacobian, lp_dev, grad_dev, apps, yscap, skip_dev,
                   0, d->ntb, (const int*)d->tgroup.p, (const int*)d->tstart.p, (const int*)d->gtile0.p, (double*)d->gmom.p, (double*)d->ypart.p, 0};
      void* gargs[] = {(void*)&ga};
      if (d->ntb > 0 && L->groupb && u_dev) {
        // phases A | B | A' as three launches: phase B gets (tiles x chains) CTAs instead of one per chain -- or, from 128 chains,
        // (small tiles x blocks of 128 chains) CTAs with one thread per chain (BPC_GROUPC=0 keeps bp_groupb)
        const char* gc = std::getenv("BPC_GROUPC");
        const bool per_chain = nchains >= 128 && L->groupc && d->ntc > 0 && !(gc && std::atoi(gc) == 0);
        if (per_chain) { ga.ntb = d->ntc; ga.tgroup = (const int*)d->tgroup_c.p; ga.tstart = (const int*)d->tstart_c.p; ga.gtile0 = (const int*)d->gtile0_c.p; ga.ypt = 1; }
        ga.mode = 1;
        BPC_CUDA(launch((const void*)L->grouped, dim3(nchains), dim3(T), gargs, smem, st));
        if (per_chain) BPC_CUDA(launch((const void*)L->groupc, dim3(d->ntc, (nchains + 127) / 128), dim3(128), gargs, 0, st));
        else BPC_CUDA(launch((const void*)L->groupb, dim3(d->ntb, nchains), dim3(128), gargs, 0, st));
        ga.mode = 2;
        BPC_CUDA(launch((const void*)L->grouped, dim3(nchains), dim3(T), gargs, smem, st));
        g_launches += 2;
      } else {
        BPC_CUDA(launch((const void*)L->grouped, dim3(nchains), dim3(T), gargs, smem, st));
      }
    } else {
      if (d->wmode) {
        // chains whose series need one sub-step: W accumulation on the fp64 tensor path + one per-chain matrix polynomial;
        // bp_fused (below) only works on the chains that need sub-steps
        if (d->cmode) BPC_CUDA(cudaMemsetAsync(d->fb.p, 0, sizeof(int), st));
        if (d->omode) {
          // octet path: per chain the transform, L, series lengths and N_0(tau_g) of every group; the fused passes of 8 chains at a
          // time; the Toeplitz products
          const int n = s.ntypes, D = n + n * (n + 1) / 2;
          BPC_CUDA(launch((const void*)L->otable, dim3(nchains), dim3(256), args, 0, st));
          if (L->fusedo2) {   // BPC_O2=1: the warp-specialised kernel
            BPC_CUDA(launch((const void*)L->fusedo2, dim3(d->ntiles, d->noct), dim3(256), args, L->o2_smem, st));
          } else if (L->fusedo9) {   // n = 3, BPC_O9=1: the two-barrier layout
            a.operm = 1;
            BPC_CUDA(launch((const void*)L->fusedo9, dim3(d->ntiles, d->noct), dim3(128), args, L->o9_smem, st));
          } else {
            BPC_CUDA(launch((const void*)L->fusedo, dim3(d->ntiles, d->noct), dim3(128), args, L->o_smem, st));
          }
          BPC_CUDA(launch((const void*)L->toep, dim3(D, d->noct), dim3(128), args, 0, st));
          g_launches += 1;   // (bp_otable stands in for bp_prologue in the count below)
        } else if (d->cmode) {
          // centred W path: per-chain tables of exp(tau L), then short series around the groups' centres
          BPC_CUDA(launch((const void*)L->ctable, dim3(nchains), dim3(256), args, 0, st));
          BPC_CUDA(launch((const void*)L->fusedc, dim3(d->ntiles, nchains), dim3(s.threads), args, s.c_smem, st));
          g_launches += 1;
        } else {
          BPC_CUDA(launch((const void*)L->fusedw, dim3(d->ntiles, nchains), dim3(s.threads), args, s.w_smem, st));
        }
        // octet path: the sub-step fallback goes first, so that bp_wpost can also do the epilogue's work (one launch less)
        if (d->omode) {
          const int fbr = std::min(nchains, std::max(8, (148 * 4 + d->ntiles - 1) / std::max(1, d->ntiles)));
          BPC_CUDA(launch((const void*)L->obs, dim3(d->ntiles, fbr), dim3(s.threads), args, L->obs_smem, st));
        }
        BPC_CUDA(launch((const void*)L->wpost, dim3(nchains), dim3(s.w_post_threads), args, 0, st));
        g_launches += 2;
      }
      // (centred / octet W paths: bp_ctable / bp_otable listed the chains that need sub-steps; a few rows of CTAs walk that list -- at
      // least 8, and enough to fill the device when the data has few tiles: far from the mode, e.g. during the ascent from wide starting
      // values on a subsample, EVERY chain is on the list; rows that find nothing to do return at once)
      const int fb_rows = std::min(nchains, std::max(8, (148 * 4 + d->ntiles - 1) / std::max(1, d->ntiles)));
      if (!(d->wmode && d->omode))
        BPC_CUDA(launch((const void*)L->obs, dim3(d->ntiles, d->cmode ? fb_rows : nchains), dim3(s.threads), args, L->obs_smem, st));
    }
    if (d->timing) {
      BPC_CUDA(cudaEventRecord(e1, st));
      d->timing_events.emplace_back(e0, e1);
    }
  }
  const bool folded = !own_io && d->wmode && d->omode;   // octet path: bp_wpost did the epilogue
  if (!own_io && !folded) {
    int ntiles = d->ntiles;
    void* args[] = {(void*)&u_dev, (void*)&theta, (void*)&part, (void*)&ntiles, (void*)&nchains, (void*)&jacobian,
                    (void*)&lp_dev, (void*)&grad_dev, (void*)&status_dev, (void*)&apps, (void*)&skip_dev};
    BPC_CUDA(launch((const void*)L->epilogue, dim3((nchains + 3) / 4), dim3(128), args, 0, st));
  }
  g_launches += own_io ? 1 : (folded ? 2 : 3);
  if (kt) {
    cudaStreamSynchronize(st);
    std::fprintf(stderr, "[ktime]");
    for (auto& ev : kev) { float ms = 0.f; cudaEventElapsedTime(&ms, ev.first, ev.second); std::fprintf(stderr, " %.1f", ms * 1e3); cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    std::fprintf(stderr, " us\n");
  }
  if (cap == cudaStreamCaptureStatusNone && (rc = release_workspace(d, st))) return rc;
  return BPC_OK;
}

}  // namespace bpc

extern "C" {

int bpc_data_path(bpc_data* d, int nchains, int* path, int* ntiles, unsigned long long* scratch_bytes) {
  if (!d || !path || !ntiles || nchains < 1) return fail(BPC_E_INVALID, "bad argument");
  const ModelSpec& s = d->model->spec;
  const int keep = d->ws_chains;
  d->ws_chains = -1;           // tiling only; no allocation
  choose_tiling(d, nchains);
  d->ws_chains = -1;
  (void)keep;
  *path = s.kind == BPC_SIMPLE_BD ? 3 : (d->grouped ? 1 : (d->wmode ? (d->omode ? 5 : (d->cmode ? 4 : 2)) : 0));
  *ntiles = d->ntiles;
  if (scratch_bytes) {
    const int n = s.ntypes, D = n + n * (n + 1) / 2;
    unsigned long long b = 2ull * nchains * d->ntiles * (s.nparams + 3) * 8;   // per-CTA partial sums, written then read
    if (d->omode) {
      // S' of every (octet, group), W, N_0 of every group and L of every chain, written then read
      b += 2ull * ((unsigned long long)d->noct * d->cgroups * (8 * n) * (8 * D) + (unsigned long long)nchains * ((((s.wcap + 1 + 8 + 7) / 8) * 8) * D * n + (unsigned long long)d->cgroups * D * n + D * D + 1)) * 8;
    } else {
      if (d->wmode) b += 2ull * nchains * d->ntiles * (unsigned long long)(((s.wcap + 1 + (d->cmode ? 8 : 0) + 7) / 8) * ((D * n + 7) / 8) * 64) * 8;
      if (d->cmode) b += 2ull * nchains * ((unsigned long long)d->cgroups * 8 * ((D * n + 7) / 8) * 8 + 65ull * D * n + 64ull * D * D) * 8;   // S^g per group and the per-chain tables, written then read
    }
    *scratch_bytes = b;
  }
  return BPC_OK;
}

int bpc_log_prob_dev(bpc_model* m, bpc_data* d, const double* upars_dev, int nchains, int jacobian, double* lp_dev,
                     double* grad_dev, int* status_dev, void* stream) {
  if (!m || !d || !upars_dev || !lp_dev || !status_dev) return fail(BPC_E_INVALID, "null argument");
  if (d->model != m) return fail(BPC_E_INVALID, "data handle belongs to another model");
  return log_prob_dev(m, d, upars_dev, nchains, jacobian, lp_dev, grad_dev, status_dev, (cudaStream_t)stream, nullptr);
}

int bpc_log_prob(bpc_model* m, bpc_data* d, const double* upars, int nchains, int jacobian, double* lp, double* grad,
                 int* status) {
  if (!m || !d || !upars || !lp) return fail(BPC_E_INVALID, "null argument");
  if (d->model != m) return fail(BPC_E_INVALID, "data handle belongs to another model");
  if (nchains <= 0) return BPC_OK;
  BPC_CUDA(cudaSetDevice(d->device));
  const int dim = m->spec.dim;
  int rc;
  // the caller's buffers are ordinary (pageable) memory: stage through pinned buffers of the handle, ONE asynchronous copy each way
  // (lp, gradient and status packed in one device buffer) -- three pageable device-to-host copies cost ~35 us more per call
  const size_t nin = (size_t)nchains * dim * 8, nlp = (size_t)nchains * 8, ngr = grad ? nin : 0, nst = (size_t)nchains * 4;
  if ((rc = d->io_u.alloc(nin))) return rc;
  if ((rc = d->io_out.alloc(nlp + ngr + nst))) return rc;
  if ((rc = d->h_in.alloc(nin))) return rc;
  if ((rc = d->h_out.alloc(nlp + ngr + nst))) return rc;
  cudaStream_t st = d->stream;
  std::memcpy(d->h_in.p, upars, nin);
  BPC_CUDA(cudaMemcpyAsync(d->io_u.p, d->h_in.p, nin, cudaMemcpyHostToDevice, st));
  char* out = (char*)d->io_out.p;
  rc = log_prob_dev(m, d, (const double*)d->io_u.p, nchains, jacobian, (double*)out, grad ? (double*)(out + nlp) : nullptr,
                    (int*)(out + nlp + ngr), st, nullptr);
  if (rc) return rc;
  BPC_CUDA(cudaMemcpyAsync(d->h_out.p, out, nlp + ngr + nst, cudaMemcpyDeviceToHost, st));
  BPC_CUDA(cudaStreamSynchronize(st));
  const char* h = (const char*)d->h_out.p;
  std::memcpy(lp, h, nlp);
  if (grad) std::memcpy(grad, h + nlp, ngr);
  if (status) std::memcpy(status, h + nlp + ngr, nst);
  return BPC_OK;
}

int bpc_log_lik(bpc_model* m, bpc_data* d, const double* upars, int nchains, double* log_lik) {
  if (!m || !d || !upars || !log_lik) return fail(BPC_E_INVALID, "null argument");
  if (d->model != m) return fail(BPC_E_INVALID, "data handle belongs to another model");
  if (nchains <= 0 || d->N == 0) return BPC_OK;
  Loaded* L;
  int rc = model_load(m, d->device, &L);
  if (rc) return rc;
  BPC_CUDA(cudaSetDevice(d->device));
  if ((rc = ensure_workspace(d, nchains))) return rc;
  const ModelSpec& s = m->spec;
  const int dim = s.dim;
  if ((rc = d->io_u.alloc((size_t)nchains * dim * 8))) return rc;
  if ((rc = d->io_status.alloc((size_t)nchains * 4))) return rc;
  DevBuf ll;
  if ((rc = ll.alloc((size_t)nchains * d->N * 8))) return rc;
  cudaStream_t st = d->stream;
  if ((rc = order_workspace(d, st))) return rc;
  BPC_CUDA(cudaMemcpyAsync(d->io_u.p, upars, (size_t)nchains * dim * 8, cudaMemcpyHostToDevice, st));
  const double* u_dev = (const double*)d->io_u.p;
  double* theta = (double*)d->theta.p;
  int* status_dev = (int*)d->io_status.p;
  {
    void* args[] = {(void*)&u_dev, (void*)&nchains, (void*)&theta, (void*)&status_dev};
    BPC_CUDA(cudaLaunchKernel((const void*)L->prologue, dim3((nchains + 127) / 128), dim3(128), args, 0, st));
  }
  {
    ObsArgs a{theta, (const double*)d->z0.p, (const double*)d->xo.p, (const double*)d->t.p, (const double*)d->xv.p,
              nullptr, status_dev, (double*)ll.p, (const int*)d->perm.p, d->N, d->npad, d->tile, d->ntiles, nchains, nullptr, 0};
    void* args[] = {(void*)&a};
    BPC_CUDA(cudaLaunchKernel((const void*)L->loglik, dim3((d->N + s.threads - 1) / s.threads, nchains), dim3(s.threads), args, 0, st));
  }
  g_launches += 2;
  BPC_CUDA(cudaMemcpyAsync(log_lik, ll.p, (size_t)nchains * d->N * 8, cudaMemcpyDeviceToHost, st));
  BPC_CUDA(cudaStreamSynchronize(st));
  return BPC_OK;
}

int bpc_moments(bpc_model* m, bpc_data* d, const double* upars, int nchains, double* mu, double* sigma) {
  if (!m || !d || !upars || !mu || !sigma) return fail(BPC_E_INVALID, "null argument");
  if (d->model != m) return fail(BPC_E_INVALID, "data handle belongs to another model");
  if (m->spec.kind == BPC_SIMPLE_BD) return fail(BPC_E_UNSUPPORTED, "bpc_moments needs a multitype model (the one-type template has its closed form)");
  if (nchains <= 0 || d->N == 0) return BPC_OK;
  Loaded* L;
  int rc = model_load(m, d->device, &L);
  if (rc) return rc;
  BPC_CUDA(cudaSetDevice(d->device));
  if ((rc = ensure_workspace(d, nchains))) return rc;
  const ModelSpec& s = m->spec;
  const int dim = s.dim, n = s.ntypes, W = n + n * n;
  if ((rc = d->io_u.alloc((size_t)nchains * dim * 8))) return rc;
  if ((rc = d->io_status.alloc((size_t)nchains * 4))) return rc;
  DevBuf out;
  if ((rc = out.alloc((size_t)nchains * d->N * W * 8))) return rc;
  cudaStream_t st = d->stream;
  if ((rc = order_workspace(d, st))) return rc;
  BPC_CUDA(cudaMemcpyAsync(d->io_u.p, upars, (size_t)nchains * dim * 8, cudaMemcpyHostToDevice, st));
  const double* u_dev = (const double*)d->io_u.p;
  double* theta = (double*)d->theta.p;
  int* status_dev = (int*)d->io_status.p;
  {
    void* args[] = {(void*)&u_dev, (void*)&nchains, (void*)&theta, (void*)&status_dev};
    BPC_CUDA(cudaLaunchKernel((const void*)L->prologue, dim3((nchains + 127) / 128), dim3(128), args, 0, st));
  }
  {
    ObsArgs a{theta, (const double*)d->z0.p, (const double*)d->xo.p, (const double*)d->t.p, (const double*)d->xv.p,
              nullptr, status_dev, (double*)out.p, (const int*)d->perm.p, d->N, d->npad, d->tile, d->ntiles, nchains, nullptr, 0};
    void* args[] = {(void*)&a};
    BPC_CUDA(cudaLaunchKernel((const void*)L->moments, dim3((d->N + s.threads - 1) / s.threads, nchains), dim3(s.threads), args, 0, st));
  }
  g_launches += 2;
  std::vector<double> h((size_t)nchains * d->N * W);
  BPC_CUDA(cudaMemcpyAsync(h.data(), out.p, h.size() * 8, cudaMemcpyDeviceToHost, st));
  BPC_CUDA(cudaStreamSynchronize(st));
  for (size_t i = 0; i < (size_t)nchains * d->N; ++i) {
    for (int j = 0; j < n; ++j) mu[i * n + j] = h[i * W + j];
    for (int j = 0; j < n * n; ++j) sigma[i * n * n + j] = h[i * W + n + j];
  }
  return BPC_OK;
}

int bpc_transformed_parameters(bpc_model* m, bpc_data* d, const double* theta, int ndraws, double* rates) {
  if (!m || !d || !theta || !rates || ndraws < 0) return fail(BPC_E_INVALID, "null argument");
  if (d->model != m) return fail(BPC_E_INVALID, "data handle belongs to another model");
  if (ndraws == 0) return BPC_OK;
  Loaded* L;
  int rc = model_load(m, d->device, &L);
  if (rc) return rc;
  BPC_CUDA(cudaSetDevice(d->device));
  const ModelSpec& s = m->spec;
  const size_t nin = (size_t)ndraws * s.dim, nout = (size_t)ndraws * d->nlev * s.nevents;
  DevBuf in, out;
  if ((rc = in.alloc(nin * 8))) return rc;
  if ((rc = out.alloc(nout * 8))) return rc;
  cudaStream_t st = d->stream;
  BPC_CUDA(cudaMemcpyAsync(in.p, theta, nin * 8, cudaMemcpyHostToDevice, st));
  const double* th_dev = (const double*)in.p;
  const double* fv = (const double*)d->fvar.p;
  double* o = (double*)out.p;
  int nlev = d->nlev, ncols = d->ncols;
  void* args[] = {(void*)&th_dev, (void*)&ndraws, (void*)&fv, (void*)&nlev, (void*)&ncols, (void*)&o};
  const long long total = (long long)ndraws * nlev;
  BPC_CUDA(cudaLaunchKernel((const void*)L->tparams, dim3((unsigned)((total + 127) / 128)), dim3(128), args, 0, st));
  g_launches += 1;
  BPC_CUDA(cudaMemcpyAsync(rates, out.p, nout * 8, cudaMemcpyDeviceToHost, st));
  BPC_CUDA(cudaStreamSynchronize(st));
  return BPC_OK;
}

int bpc_data_nlevels(const bpc_data* d) { return d ? d->nlev : -1; }

int bpc_flop_model(const bpc_model* m, double* per_app_fwd, double* per_app_bwd, double* per_obs) {
  if (!m || !per_app_fwd || !per_app_bwd || !per_obs) return fail(BPC_E_INVALID, "null argument");
  const FlopModel f = flop_model(m->spec);
  *per_app_fwd = f.per_app_fwd; *per_app_bwd = f.per_app_bwd; *per_obs = f.per_obs;
  return BPC_OK;
}

int bpc_flop_model_centred(const bpc_model* m, double* per_term, double* per_obs_centred, double* per_obs_octet) {
  if (!m || !per_term || !per_obs_centred || !per_obs_octet) return fail(BPC_E_INVALID, "null argument");
  const FlopModel f = flop_model(m->spec);
  *per_term = f.per_term_c; *per_obs_centred = f.per_obs_c; *per_obs_octet = f.per_obs_o;
  return BPC_OK;
}

int bpc_kernel_timing(bpc_data* d, int enable) {
  if (!d) return fail(BPC_E_INVALID, "null argument");
  d->timing = enable != 0;
  return BPC_OK;
}

int bpc_kernel_time_ms(bpc_data* d, int reset, double* total_ms, long long* launches) {
  if (!d || !total_ms) return fail(BPC_E_INVALID, "null argument");
  BPC_CUDA(cudaSetDevice(d->device));
  double tot = 0.0;
  for (auto& ev : d->timing_events) {
    BPC_CUDA(cudaEventSynchronize(ev.second));
    float ms = 0.f;
    BPC_CUDA(cudaEventElapsedTime(&ms, ev.first, ev.second));
    tot += ms;
  }
  *total_ms = tot;
  if (launches) *launches = (long long)d->timing_events.size();
  if (reset) {
    for (auto& ev : d->timing_events) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    d->timing_events.clear();
  }
  return BPC_OK;
}

// Taylor applications of L and L^T that the last bpc_log_prob* call performed, per chain; with the
// per-application and per-observation constants of DESIGN.md §4 this is the exact algorithmic flop count.
int bpc_count_flops(bpc_model* m, bpc_data* d, const double* upars, int nchains, double* flops_per_chain) {
  if (!m || !d || !upars || !flops_per_chain) return fail(BPC_E_INVALID, "null argument");
  std::vector<double> lp(nchains);
  int rc = bpc_log_prob(m, d, upars, nchains, 1, lp.data(), nullptr, nullptr);
  if (rc) return rc;
  std::vector<double> apps(nchains);
  BPC_CUDA(cudaMemcpy(apps.data(), d->apps.p, (size_t)nchains * 8, cudaMemcpyDeviceToHost));
  const ModelSpec& s = m->spec;
  const FlopModel f = flop_model(s);
  for (int c = 0; c < nchains; ++c)
    // (a W-path chain that fell back to the ring kernel is counted with the W-path constants: a lower bound)
    flops_per_chain[c] = (s.kind == BPC_SIMPLE_BD) ? f.per_obs * d->N
                         : (d->cmode && d->omode) ? apps[c] + f.per_obs_o * d->N   // (bp_otable counts the flops of the terms and groups)
                         : d->cmode              ? apps[c] + f.per_obs_c * d->N   // (bp_fusedc counts the flops of its terms and groups itself)
                         : d->wmode              ? apps[c] * (f.per_app_fwd + f.per_app_w) + f.per_obs_w * d->N
                                                 : apps[c] * (f.per_app_fwd + f.per_app_bwd) + (d->grouped ? f.per_obs_grouped : f.per_obs) * d->N;
  return BPC_OK;
}

}  // extern "C"
