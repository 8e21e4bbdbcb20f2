// experiment: fp64 tensor-core (mma.sync m8n8k4 f64) throughput on this device, next to the DFMA loop
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) dmma_loop(double* out, int iters, double a, double b) {
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
int main() {
  double* out; cudaMalloc(&out, 148 * 8 * 256 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int warps = 1; warps <= 8; warps *= 2) {
    int blocks = 148 * 4, threads = 32 * warps * 1; 
    dmma_loop<<<blocks, threads>>>(out, 100, 1.0000001, 1e-9); cudaDeviceSynchronize();
    int iters = 20000;
    cudaEventRecord(e0); dmma_loop<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 256 * 32.0 * iters * blocks * (threads / 32);
    printf("threads/CTA %d (CTAs %d): %.2f TFLOP/s fp64 tensor (m8n8k4), %.3f ms, err %s\n", threads, blocks, flops / (ms * 1e-3) / 1e12, ms, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
