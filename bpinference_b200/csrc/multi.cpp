// multi.cpp — several GPUs inside ONE process behind the C ABI (SURVEY.md §8(b) "Threading", §8(e)).
//
// The reference runs one chain per R worker process (options(mc.cores), examples/two_type.R:20); an R caller of libbpcuda is a
// single process and R's API is not thread-safe, so the sharding lives here: bpc_sample_multi takes the list of devices, replicates
// the data list and the compiled module on each of them, gives every device its own host thread and its share of the chains
// (global chain ids device_index * chains + c: the Philox streams do not depend on the sharding), and owns ONE single-process NCCL
// communicator (ncclCommInitAll) for the only exchanges of the path:
//   * the pooled warm-up statistics at the window boundaries: 2 + 2 dim doubles (count, sum q, sum q^2, "finished" flag), all-reduce
//   * the half-chain moments behind split R-hat at the end: 4 dim doubles, all-reduce
// Both are latency-bound and off the critical path, which is why plain ncclAllReduce on a side stream is the right tool: there
// is no compute step fused with a collective anywhere in an evaluation.  Nothing of torch is involved.
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): the library has no link-time dependency on it, single-GPU users never load
// it, and inside a Python process that already carries torch's NCCL the same copy is reused.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <atomic>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "bpcuda.h"
#include "bpcuda_internal.hpp"

using namespace bpc;

namespace {

struct Nccl {
  void* h = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  std::string err;
};

Nccl& nccl() {
  static Nccl n;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      n.h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (n.h) break;
    }
    if (!n.h) { n.err = std::string("NCCL not found (dlopen libnccl.so.2): ") + (dlerror() ? dlerror() : ""); return; }
    auto sym = [&](const char* s) { void* p = dlsym(n.h, s); if (!p && n.err.empty()) n.err = std::string("NCCL symbol missing: ") + s; return p; };
    n.CommInitAll = (decltype(n.CommInitAll))sym("ncclCommInitAll");
    n.CommDestroy = (decltype(n.CommDestroy))sym("ncclCommDestroy");
    n.AllReduce = (decltype(n.AllReduce))sym("ncclAllReduce");
    n.GetErrorString = (decltype(n.GetErrorString))sym("ncclGetErrorString");
    n.GetVersion = (decltype(n.GetVersion))sym("ncclGetVersion");
  });
  return n;
}

struct Worker {
  int index = 0, device = 0, rc = BPC_OK;
  std::string error;
  bpc_data* data = nullptr;
  bpc_sampler* sampler = nullptr;
  cudaStream_t stream = nullptr;
  double* pool_dev = nullptr;      // 2 + 2 dim
  double* mom_dev = nullptr;       // 4 dim
  long long exchanges = 0, launches = 0;
};

}  // namespace

extern "C" {

int bpc_nccl_version(void) {
  Nccl& n = nccl();
  int v = 0;
  if (!n.h || !n.GetVersion || n.GetVersion(&v) != ncclSuccess) return 0;
  return v;
}

int bpc_sample_multi(bpc_model* m, const bpc_stan_data* data, const int* devices, int ndev, const bpc_sample_args* a, double* draws,
                     double* lp, int* divergent, int* treedepth, int* n_leapfrog, double* accept_stat, double* stepsize, double* rhat,
                     long long* info) {
  if (!m || !data || !devices || !a || ndev < 1) return fail(BPC_E_INVALID, "null argument");
  const int nvis = bpc_device_count();
  if (nvis == 0) return fail(BPC_E_CUDA, "no CUDA device available: libbpcuda has no CPU fallback");
  for (int i = 0; i < ndev; ++i) {
    if (devices[i] < 0 || devices[i] >= nvis) return fail(BPC_E_INVALID, "device index out of range");
    for (int j = 0; j < i; ++j)
      if (devices[j] == devices[i]) return fail(BPC_E_INVALID, "a device may appear once in the list (one NCCL rank per GPU)");
  }
  Nccl& nc = nccl();
  if (!nc.h || !nc.err.empty()) return fail(BPC_E_UNSUPPORTED, nc.err.empty() ? "NCCL unavailable" : nc.err);
  std::vector<ncclComm_t> comms(ndev);
  {
    const ncclResult_t r = nc.CommInitAll(comms.data(), ndev, devices);
    if (r != ncclSuccess) return fail(BPC_E_CUDA, std::string("ncclCommInitAll: ") + nc.GetErrorString(r));
  }
  const int dim = bpc_model_dim(m), C = a->chains, ns = a->iter - a->warmup;
  const int npool = 2 + (a->metric == 1 ? dim + dim * (dim + 1) / 2 : 2 * dim);   // bpc_sampler_pool_len
  std::vector<Worker> W(ndev);
  // A worker that fails keeps taking part in the exchanges with zeros and a raised "finished" flag, so that nobody waits for it in
  // a collective; the first error is reported when all threads have joined.
  auto body = [&](Worker& w) {
    auto failw = [&](int rc) { w.rc = rc; w.error = bpc_last_error(); };
    cudaError_t e;
    if ((e = cudaSetDevice(w.device)) != cudaSuccess) { w.rc = BPC_E_CUDA; w.error = cudaGetErrorString(e); }
    if (!w.rc && ((e = cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking)) != cudaSuccess ||
                  (e = cudaMalloc(&w.pool_dev, (size_t)npool * 8)) != cudaSuccess || (e = cudaMalloc(&w.mom_dev, (size_t)4 * dim * 8)) != cudaSuccess)) {
      w.rc = BPC_E_CUDA; w.error = cudaGetErrorString(e);
    }
    int rc;
    if (!w.rc && (rc = bpc_data_create(m, data, w.device, 0, &w.data))) failw(rc);
    bpc_sample_args la = *a;
    la.chain_id_offset = a->chain_id_offset + w.index * C;
    la.init_opt_data = nullptr;   // (a subsample handle belongs to one device; the multi-device run does the one-stage ascent)
    if (a->inits) la.inits = a->inits + (size_t)w.index * C * m->spec.nparams;
    if (!w.rc && (rc = bpc_sampler_create(m, w.data, &la, &w.sampler))) failw(rc);
    int state = 0;
    bool all_done = false;
    while (!all_done) {
      if (!w.rc && (rc = bpc_sampler_run(w.sampler, 1LL << 40, &state))) failw(rc);
      if (!a->pooled_adaptation) {
        if (w.rc || state == 1) break;
        continue;   // (state 2 cannot occur without pooling)
      }
      if (w.pool_dev) {
        if (!w.rc && (rc = bpc_sampler_pool_stats(w.sampler, w.pool_dev, npool))) failw(rc);
        if (w.rc) {   // zeros and "finished"
          std::vector<double> z(npool, 0.0);
          z[npool - 1] = 1.0;
          cudaMemcpy(w.pool_dev, z.data(), (size_t)npool * 8, cudaMemcpyHostToDevice);
        }
      }
      const ncclResult_t r = nc.AllReduce(w.pool_dev, w.pool_dev, (size_t)npool, ncclDouble, ncclSum, comms[w.index], w.stream);
      if (r != ncclSuccess && !w.rc) { w.rc = BPC_E_CUDA; w.error = std::string("ncclAllReduce: ") + nc.GetErrorString(r); }
      cudaStreamSynchronize(w.stream);
      ++w.exchanges;
      double flag = (double)ndev;
      if (r == ncclSuccess) cudaMemcpy(&flag, w.pool_dev + (npool - 1), 8, cudaMemcpyDeviceToHost);
      all_done = flag >= (double)ndev;
      if (!all_done && !w.rc && state == 2 && (rc = bpc_sampler_pool_apply(w.sampler, w.pool_dev))) failw(rc);
    }
    // results of this device's chains, and the R-hat moments
    const size_t off = (size_t)w.index * C * ns;
    if (!w.rc && (rc = bpc_sampler_draws(w.sampler, draws ? draws + off * dim : nullptr, lp ? lp + off : nullptr, divergent ? divergent + off : nullptr,
                                         treedepth ? treedepth + off : nullptr, n_leapfrog ? n_leapfrog + off : nullptr,
                                         accept_stat ? accept_stat + off : nullptr, stepsize ? stepsize + off : nullptr))) failw(rc);
    if (w.mom_dev) {
      if (!w.rc && (rc = bpc_sampler_chain_moments(w.sampler, w.mom_dev, 4 * dim, nullptr))) failw(rc);
      if (w.rc) cudaMemset(w.mom_dev, 0, (size_t)4 * dim * 8);
    }
    const ncclResult_t r2 = nc.AllReduce(w.mom_dev, w.mom_dev, (size_t)4 * dim, ncclDouble, ncclSum, comms[w.index], w.stream);
    if (r2 != ncclSuccess && !w.rc) { w.rc = BPC_E_CUDA; w.error = std::string("ncclAllReduce: ") + nc.GetErrorString(r2); }
    cudaStreamSynchronize(w.stream);
    ++w.exchanges;
    w.launches = bpc_launch_count(0);
  };
  std::vector<std::thread> th;
  for (int i = 0; i < ndev; ++i) { W[i].index = i; W[i].device = devices[i]; }
  for (int i = 0; i < ndev; ++i) th.emplace_back(body, std::ref(W[i]));
  for (auto& t : th) t.join();
  int rc = BPC_OK;
  std::string msg;
  long long passes = 0, leap = 0;
  for (auto& w : W) {
    if (w.rc && !rc) { rc = w.rc; msg = "device " + std::to_string(w.device) + ": " + w.error; }
    if (w.sampler) { passes = std::max(passes, bpc_sampler_passes(w.sampler)); if (!w.rc) leap += bpc_sampler_total_leapfrogs(w.sampler); }
  }
  if (!rc && rhat) {
    std::vector<double> red((size_t)4 * dim);
    cudaSetDevice(W[0].device);
    if (cudaMemcpy(red.data(), W[0].mom_dev, red.size() * 8, cudaMemcpyDeviceToHost) != cudaSuccess) { rc = BPC_E_CUDA; msg = "cudaMemcpy (moments)"; }
    else if (ns >= 2) rc = bpc_rhat(red.data(), dim, ns, rhat);
    if (rc && msg.empty()) msg = bpc_last_error();
  }
  if (info) { info[0] = passes; info[1] = leap; info[2] = W[0].exchanges; info[3] = bpc_nccl_version(); }
  long long launches = 0;
  for (auto& w : W) {
    cudaSetDevice(w.device);
    if (w.sampler) bpc_sampler_free(w.sampler);
    if (w.data) bpc_data_free(w.data);
    if (w.pool_dev) cudaFree(w.pool_dev);
    if (w.mom_dev) cudaFree(w.mom_dev);
    if (w.stream) cudaStreamDestroy(w.stream);
    launches += w.launches;
  }
  for (auto c : comms) nc.CommDestroy(c);
  g_launches += launches;
  if (rc) return fail(rc, msg);
  return BPC_OK;
}

}  // extern "C"
