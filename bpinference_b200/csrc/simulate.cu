// simulate.cu — device Gillespie simulator: the ensemble version of bp() / bpsims() (R/simulation.R:9-56, 69-131).
// One thread per replicate, Philox4x32-10 keyed by (seed, replicate).  Same recording rule as the reference's loop
// (R/simulation.R:49-51): a time point crossed by the waiting time of an event records the state AFTER that event;
// the first row is the initial population; after extinction the remaining rows keep the last state.
// Data generator for parity datasets and simulation-based checks — not on the log-density path.
#include <cuda_runtime.h>

#include <cmath>
#include <vector>

#include "bpcuda.h"
#include "bpcuda_internal.hpp"

using namespace bpc;

namespace {

constexpr int SIM_MAXE = 64, SIM_MAXN = 8;

struct SimParams {
  int n, E, ntimes, nrep;
  unsigned long long seed;
  double e_mat[SIM_MAXE][SIM_MAXN];
  double r[SIM_MAXE];
  int parent[SIM_MAXE];
  double z0[SIM_MAXN];
};

__device__ __forceinline__ void philox_round(unsigned& c0, unsigned& c1, unsigned& c2, unsigned& c3, unsigned k0, unsigned k1) {
  const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
  const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
  const unsigned n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
  c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}

__device__ void philox2(unsigned long long seed, unsigned rep, unsigned long long ctr, double& u0, double& u1) {
  unsigned c0 = (unsigned)ctr, c1 = (unsigned)(ctr >> 32), c2 = rep, c3 = 0x5eed5eedu;
  unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) { philox_round(c0, c1, c2, c3, k0, k1); k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
  u0 = (((unsigned long long)c0 << 21) ^ (c1 >> 11)) * (1.0 / 9007199254740992.0) + (0.5 / 9007199254740992.0);
  u1 = (((unsigned long long)c2 << 21) ^ (c3 >> 11)) * (1.0 / 9007199254740992.0) + (0.5 / 9007199254740992.0);
}

__global__ void bp_gillespie(SimParams P, const double* __restrict__ times, double* __restrict__ pops, int* __restrict__ overflow) {
  const int rep = blockIdx.x * blockDim.x + threadIdx.x;
  if (rep >= P.nrep) return;
  double z[SIM_MAXN];
  for (int i = 0; i < P.n; ++i) z[i] = P.z0[i];
  double* out = pops + (size_t)rep * P.ntimes * P.n;
  for (int i = 0; i < P.n; ++i) out[i] = z[i];
  const double tf = times[P.ntimes - 1];
  double t = 0.0;
  int next = 0;                       // first time point with times[next] > t (times are increasing)
  while (next < P.ntimes && !(times[next] > t)) ++next;
  unsigned long long ctr = 0;
  while (t < tf) {
    double lam = 0.0;
    for (int e = 0; e < P.E; ++e) lam += P.r[e] * z[P.parent[e]];
    if (!(lam > 0.0)) {               // extinction (or no event can fire): the state stays
      for (int j = next; j < P.ntimes; ++j) for (int i = 0; i < P.n; ++i) out[(size_t)j * P.n + i] = z[i];
      return;
    }
    double u0, u1;
    philox2(P.seed, (unsigned)rep, ctr++, u0, u1);
    const double dt = -log(u0) / lam;
    const double pick = u1 * lam;
    double cum = 0.0;
    int ev = P.E - 1;
    for (int e = 0; e < P.E; ++e) { cum += P.r[e] * z[P.parent[e]]; if (pick < cum) { ev = e; break; } }
    for (int i = 0; i < P.n; ++i) z[i] += P.e_mat[ev][i];
    z[P.parent[ev]] -= 1.0;
    const double tn = t + dt;
    while (next < P.ntimes && tn >= times[next]) {
      for (int i = 0; i < P.n; ++i) out[(size_t)next * P.n + i] = z[i];
      ++next;
    }
    t = tn;
    if (ctr > (1ull << 40)) { atomicExch(overflow, 1); return; }
  }
}

}  // namespace

extern "C" int bpc_simulate(int ntypes, int nevents, const double* e_mat, const int* p_vec, const double* r_vec, const double* z0,
                            const double* times, int ntimes, int nrep, unsigned long long seed, int device, double* pops) {
  if (!e_mat || !p_vec || !r_vec || !z0 || !times || !pops) return fail(BPC_E_INVALID, "null argument");
  if (ntypes < 1 || ntypes > SIM_MAXN || nevents < 1 || nevents > SIM_MAXE) return fail(BPC_E_UNSUPPORTED, "bpc_simulate supports up to 8 types and 64 events");
  if (ntimes < 1 || nrep < 1) return fail(BPC_E_INVALID, "need at least one time point and one replicate");
  for (int j = 1; j < ntimes; ++j) if (!(times[j] > times[j - 1])) return fail(BPC_E_INVALID, "times must be increasing");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); return fail(BPC_E_CUDA, "no CUDA device available: libbpcuda has no CPU fallback"); }
  cudaError_t e;
  if ((e = cudaSetDevice(device)) != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
  SimParams P{};
  P.n = ntypes; P.E = nevents; P.ntimes = ntimes; P.nrep = nrep; P.seed = seed;
  for (int ev = 0; ev < nevents; ++ev) {
    if (p_vec[ev] < 1 || p_vec[ev] > ntypes) return fail(BPC_E_INVALID, "Parent vector must have integer entries between 1 and the number of types!");
    if (!(r_vec[ev] >= 0.0)) return fail(BPC_E_INVALID, "rates must be non-negative");
    P.parent[ev] = p_vec[ev] - 1;
    P.r[ev] = r_vec[ev];
    for (int i = 0; i < ntypes; ++i) P.e_mat[ev][i] = e_mat[ev + (size_t)nevents * i];
  }
  for (int i = 0; i < ntypes; ++i) P.z0[i] = z0[i];
  DevBuf dt, dp, dov;
  int rc;
  if ((rc = dt.alloc((size_t)ntimes * 8))) return rc;
  if ((rc = dp.alloc((size_t)nrep * ntimes * ntypes * 8))) return rc;
  if ((rc = dov.alloc(4))) return rc;
  cudaMemcpy(dt.p, times, (size_t)ntimes * 8, cudaMemcpyHostToDevice);
  cudaMemset(dov.p, 0, 4);
  bp_gillespie<<<(nrep + 127) / 128, 128>>>(P, (const double*)dt.p, (double*)dp.p, (int*)dov.p);
  g_launches += 1;
  if ((e = cudaDeviceSynchronize()) != cudaSuccess) return cuda_fail(e, "bp_gillespie");
  if ((e = cudaMemcpy(pops, dp.p, (size_t)nrep * ntimes * ntypes * 8, cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
  int ov = 0;
  cudaMemcpy(&ov, dov.p, 4, cudaMemcpyDeviceToHost);
  if (ov) return fail(BPC_E_INVALID, "simulation exceeded the event budget (population explosion)");
  return BPC_OK;
}
