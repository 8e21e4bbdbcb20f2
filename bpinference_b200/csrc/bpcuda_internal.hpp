// bpcuda_internal.hpp — handle layouts shared by the translation units of libbpcuda (not part of the ABI).
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "codegen.hpp"

namespace bpc {

extern thread_local std::string g_last_error;
extern thread_local long long g_launches;

// Launch of a kernel of the evaluation / sampler chains with programmatic dependent launch: the kernel's CTAs may become resident while
// the previous kernel of the stream still runs; every such kernel starts with bp_pdl_enter() (griddepcontrol.wait: the previous kernel
// has completed and flushed), so the order of the stream is kept.  Used for eager launches only (log_prob_dev has the measurement).
// BPC_PDL=0: ordinary launches.
inline bool pdl_enabled() {
  static const bool on = [] { const char* e = std::getenv("BPC_PDL"); return !(e && std::atoi(e) == 0); }();
  return on;
}
inline cudaError_t pdl_launch(const void* f, dim3 grid, dim3 block, void** args, size_t smem, cudaStream_t st) {
  if (!pdl_enabled()) return cudaLaunchKernel(f, grid, block, args, smem, st);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelExC(&cfg, f, args);
}
int fail(int code, const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);

// device allocation that only grows
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int alloc(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    cap = bytes;
    return 0;
  }
  ~DevBuf() { if (p) cudaFree(p); }
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
};

struct HostBuf {   // pinned host memory (staging of the host-buffer entry points: one asynchronous copy each way)
  void* p = nullptr;
  size_t cap = 0;
  int alloc(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) return cuda_fail(e, "cudaHostAlloc");
    cap = bytes;
    return 0;
  }
  ~HostBuf() { if (p) cudaFreeHost(p); }
  HostBuf() = default;
  HostBuf(const HostBuf&) = delete;
  HostBuf& operator=(const HostBuf&) = delete;
};

struct Loaded {   // one NVRTC module loaded on one device
  cudaLibrary_t lib = nullptr;
  cudaKernel_t prologue = nullptr, obs = nullptr, epilogue = nullptr, loglik = nullptr, grouped = nullptr, moments = nullptr, fusedw = nullptr, wpost = nullptr, fusedc = nullptr, ctable = nullptr, otable = nullptr, fusedo = nullptr, toep = nullptr, tparams = nullptr, fusedo9 = nullptr, fusedo2 = nullptr, groupb = nullptr, groupc = nullptr;
  size_t obs_smem = 0, grouped_smem = 0, o_smem = 0, o9_smem = 0, o2_smem = 0;
};

// algorithmic flop constants of the shipped kernels (SURVEY.md §8(d) convention: FMA = 2, div/sqrt/exp/log = 1)
struct FlopModel { double per_app_fwd, per_app_bwd, per_obs, per_obs_grouped, per_app_w, per_obs_w, per_term_c, per_obs_c, per_obs_o; };
FlopModel flop_model(const ModelSpec& s);

}  // namespace bpc

struct bpc_model {
  bpc::ModelSpec spec;
  std::string source, warnings, compile_log;
  std::vector<char> cubin;
  std::map<int, bpc::Loaded> loaded;
  std::mutex load_mu;   // bpc_sample_multi loads the module on several devices from several threads
};

struct CGrouping {   // groups of consecutive chunks of the sorted table that share a series centre
  int n = 0, maxlen = 0, gmax = 0;   // groups, chunks of the longest one, the cap they were built with
  bpc::DevBuf chunk0, idx, tau, hmax, coef;
};

struct bpc_data;
namespace bpc { int build_grouping(bpc_data* d, int gmax, CGrouping& G, bool store); }

struct bpc_data {
  const bpc_model* model = nullptr;
  int device = 0;
  int N = 0, npad = 0, ngroups = 0;
  bool grouped = false;
  bpc::DevBuf z0, xo, t, xv, perm;
  bpc::DevBuf fvar;                 // function_var as handed over (ndep_levels x ndep, column-major): bpc_transformed_parameters
  int nlev = 1, ncols = 1;
  bpc::DevBuf wpart;   // W path: per-CTA partial sums of the W_j matrices
  bool wmode = false;
  // centred W path (bp_ctable + bp_fusedc): groups of consecutive 32-observation chunks sharing a series centre
  bool cmode = false;
  int cgroups = 0, cgpc = 4, cgsel = 0;   // groups of the grouping in use: cgs[cgsel] (0: bp_fusedc, 1: octet path)
  double cdelta = 0.0, cg_hcap = 0.0;
  int cg_fixed = 0;                 // BPC_CGROUP_CHUNKS: the same cap for both groupings, no search
  std::vector<double> h_t;          // sorted durations (host copy while bpc_data_create builds the groupings)
  CGrouping cgs[2];
  bpc::DevBuf ctab_m1, ctab_e2, csg;
  // octet path (bp_otable + bp_fusedo + bp_toep): eight chains per tensor tile; chosen per chain count (choose_tiling)
  bool omode = false;
  int noct = 0;
  bpc::DevBuf cg_jm, cg_flops, osg, wplain, fb, ctab_l, n0tab;
  bpc::DevBuf gstart, gt, gx, glev, ring;
  int nlevels = 1;   // grouped path: first observation, duration and dependent variables of each group; term ring
  // phase B of the grouped path on its own grid (many observations per group): tiles of <= 4096 observations of one group
  int ntb = 0;
  bpc::DevBuf tgroup, tstart, gtile0, gmom, ypart;
  // ... and with one thread per chain from 128 chains (bp_groupc): tiles of <= 256 observations
  int ntc = 0;
  bpc::DevBuf tgroup_c, tstart_c, gtile0_c;
  int group_threads = 128;
  // workspace sized for ws_chains chains.  ws_gen counts every re-tiling / reallocation: a sampler that replays a captured CUDA
  // graph (raw workspace pointers and tiling baked in) compares it with the generation it captured under and re-captures.
  int ws_chains = -1, tile = 0, ntiles = 0;
  unsigned long long ws_gen = 0;
  cudaEvent_t ws_event = nullptr;   // recorded at the end of every use of the workspace, waited for at the start of the next
  bpc::DevBuf theta, part, apps;
  bpc::DevBuf io_u, io_status, io_out;
  bpc::HostBuf h_in, h_out;
  cudaStream_t stream = nullptr;
  bool timing = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timing_events;
};

namespace bpc {
int model_load(bpc_model* m, int device, Loaded** out);
int order_workspace(bpc_data* d, cudaStream_t st);
int release_workspace(bpc_data* d, cudaStream_t st);
int log_prob_dev(bpc_model* m, bpc_data* d, const double* u_dev, int nchains, int jacobian, double* lp_dev,
                 double* grad_dev, int* status_dev, cudaStream_t st, const int* skip_dev);
}  // namespace bpc
