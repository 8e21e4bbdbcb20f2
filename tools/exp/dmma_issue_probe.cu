// How many warps per SM sub-partition, with how many independent accumulator chains each, does the fp64 tensor path (DMMA,
// mma.sync.m8n8k4.f64) need to reach its peak?  Decides how the octet kernel spreads its tensor products over warps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/exp/dmma_issue_probe tools/exp/dmma_issue_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CH>
__global__ void probe(double* out, int iters, double a, double b) {
  double c[CH][2];
#pragma unroll
  for (int k = 0; k < CH; ++k) { c[k][0] = 0; c[k][1] = 0; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 32 / CH; ++j)
#pragma unroll
      for (int k = 0; k < CH; ++k)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < CH; ++k) s += c[k][0] + c[k][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA chains next to DFMA chains in the SAME warp / in OTHER warps of the SM: do the two share issue slots as well as the datapath?
template <int CH>
__global__ void mixed(double* out, int iters, double a, double b, int fma_warps) {
  const int warp = threadIdx.x >> 5;
  double s = 0;
  if (warp < fma_warps) {
    double f[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = k;
    for (int i = 0; i < iters; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] = fma(f[k], a, b);
#pragma unroll
    for (int k = 0; k < 8; ++k) s += f[k];
  } else {
    double c[CH][2];
#pragma unroll
    for (int k = 0; k < CH; ++k) { c[k][0] = 0; c[k][1] = 0; }
    for (int i = 0; i < iters; ++i)
#pragma unroll
      for (int j = 0; j < 32 / CH; ++j)
#pragma unroll
        for (int k = 0; k < CH; ++k)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(a), "d"(b));
#pragma unroll
    for (int k = 0; k < CH; ++k) s += c[k][0] + c[k][1];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH>
static double run(int warps_per_sm, int sms, double* out) {
  const int iters = 4000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<CH><<<sms, warps_per_sm * 32>>>(out, 200, 1.0000001, 1e-9);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    probe<CH><<<sms, warps_per_sm * 32>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float t; cudaEventElapsedTime(&t, e0, e1);
    if (t < best) best = t;
  }
  return 2.0 * 256 * 32.0 * iters * sms * warps_per_sm / (best * 1e-3) / 1e12;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double* out; cudaMalloc(&out, (size_t)sms * 1024 * 8);
  printf("TFLOP/s of DMMA by warps per SM (rows) and independent accumulator chains per warp (columns 1 2 4 8 16)\n");
  for (int w : {4, 8, 12, 16, 24, 32}) {
    printf("warps/SM %2d:", w);
    printf(" %6.2f", run<1>(w, sms, out));
    printf(" %6.2f", run<2>(w, sms, out));
    printf(" %6.2f", run<4>(w, sms, out));
    printf(" %6.2f", run<8>(w, sms, out));
    printf(" %6.2f\n", run<16>(w, sms, out));
  }
  // mixed: 16 warps per SM, k of them DFMA (8 chains), the rest DMMA (8 chains): total flops of both
  printf("16 warps per SM, k DFMA warps + (16 - k) DMMA warps: DMMA TF, DFMA TF\n");
  for (int k : {0, 4, 8, 12, 16}) {
    const int iters = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    mixed<8><<<sms, 512>>>(out, 200, 1.0000001, 1e-9, k);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    mixed<8><<<sms, 512>>>(out, iters, 1.0000001, 1e-9, k);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float t; cudaEventElapsedTime(&t, e0, e1);
    printf("k %2d: whole kernel %.3f ms; DMMA warps alone would be %.2f TF, DFMA warps %.2f TF over that time\n", k, t,
           2.0 * 256 * 32.0 * iters * sms * (16 - k) / (t * 1e-3) / 1e12, 2.0 * 32 * 32.0 * iters * sms * k / (t * 1e-3) / 1e12);
  }
  return 0;
}
