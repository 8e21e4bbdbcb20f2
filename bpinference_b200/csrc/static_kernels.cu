// static_kernels.cu — kernels of libbpcuda that do not depend on the model (compiled by nvcc for sm_100a):
//   * bp_dfma_peak    register-resident DFMA loop: measures the FP64 roofline denominator on this device
//   * default_model   bp_kernels.cuh instantiated for a reference 3-type model, so that build() compiles the
//                     hand-written kernels with nvcc (-Xptxas -v / cuobjdump evidence) even before NVRTC runs
#include <cuda_runtime.h>

#include "bpcuda.h"
#include "bpcuda_internal.hpp"

namespace {

__global__ void __launch_bounds__(256) bp_dfma_peak(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1.0, x2 = x0 + 2.0, x3 = x0 + 3.0, x4 = x0 + 4.0, x5 = x0 + 5.0, x6 = x0 + 6.0,
         x7 = x0 + 7.0;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// the same for the fp64 tensor path (mma.sync m8n8k4, DMMA in SASS): four independent accumulator chains per warp
__global__ void __launch_bounds__(256) bp_dmma_peak(double* out, int iters, double a, double b) {
  double c0 = 0, c1 = 0, d0 = 0, d1 = 0, e0 = 0, e1 = 0, f0 = 0, f1 = 0;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(e0), "+d"(e1) : "d"(a), "d"(b));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(f0), "+d"(f1) : "d"(a), "d"(b));
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1 + d0 + d1 + e0 + e1 + f0 + f1;
}

}  // namespace

extern "C" int bpc_fp64_tensor_peak(int device, double ms, double* tflops) {
  using namespace bpc;
  if (!tflops) return fail(BPC_E_INVALID, "null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); return fail(BPC_E_CUDA, "no CUDA device available"); }
  cudaError_t e;
  if ((e = cudaSetDevice(device)) != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  const int blocks = prop.multiProcessorCount * 4, threads = 256;
  double* out = nullptr;
  if ((e = cudaMalloc(&out, (size_t)blocks * threads * 8)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  int iters = 2000;
  double best = 0.0, spent = 0.0;
  bp_dmma_peak<<<blocks, threads>>>(out, 100, 1.0000001, 1e-9);
  cudaDeviceSynchronize();
  while (spent < ms) {
    cudaEventRecord(e0);
    bp_dmma_peak<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    if ((e = cudaEventSynchronize(e1)) != cudaSuccess) { cudaFree(out); return cuda_fail(e, "bp_dmma_peak"); }
    float t = 0.f;
    cudaEventElapsedTime(&t, e0, e1);
    const double flops = 2.0 * 256 * 32.0 * (double)iters * blocks * (threads / 32);   // 8 x 8 x 4 MACs per mma, 32 mma per iteration per warp
    best = std::max(best, flops / (t * 1e-3) / 1e12);
    spent += t;
    g_launches += 1;
    if (t < 5.0) iters *= 2;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *tflops = best;
  return BPC_OK;
}

extern "C" int bpc_fp64_peak(int device, double ms, double* tflops, double* sm_clock_mhz) {
  using namespace bpc;
  if (!tflops) return fail(BPC_E_INVALID, "null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); return fail(BPC_E_CUDA, "no CUDA device available"); }
  cudaError_t e;
  if ((e = cudaSetDevice(device)) != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  const int blocks = prop.multiProcessorCount * 8, threads = 256;
  double* out = nullptr;
  if ((e = cudaMalloc(&out, (size_t)blocks * threads * 8)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  int iters = 2000;
  double best = 0.0;
  double spent = 0.0;
  bp_dfma_peak<<<blocks, threads>>>(out, 200, 0.999999, 1e-9);   // warm-up
  cudaDeviceSynchronize();
  while (spent < ms) {
    cudaEventRecord(e0);
    bp_dfma_peak<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    if ((e = cudaEventSynchronize(e1)) != cudaSuccess) { cudaFree(out); return cuda_fail(e, "bp_dfma_peak"); }
    float t = 0.f;
    cudaEventElapsedTime(&t, e0, e1);
    const double flops = 2.0 * 8 * 16 * (double)iters * blocks * threads;
    best = std::max(best, flops / (t * 1e-3) / 1e12);
    spent += t;
    g_launches += 1;
    if (t < 5.0) iters *= 2;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *tflops = best;
  if (sm_clock_mhz) { int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device); *sm_clock_mhz = khz / 1000.0; }
  return BPC_OK;
}
