// experiment: do DFMA and DMMA (mma.sync m8n8k4 f64) share one datapath on this part?  Times nd DFMAs + nm DMMAs per
// loop iteration per warp for several (nd, nm); "sum" behaviour = shared pipe, "max" behaviour = separate pipes.
#include <cstdio>
#include <cuda_runtime.h>
template <int ND, int NM>
__global__ void __launch_bounds__(128) mix(double* out, int iters, double a, double b) {
  double f[8] = {0, 1, 2, 3, 4, 5, 6, 7};
  double c[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int j = 0; j < ND; ++j) f[j & 7] = fma(f[j & 7], a, b);
#pragma unroll
      for (int j = 0; j < NM; ++j)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[j & 3][0]), "+d"(c[j & 3][1]) : "d"(a), "d"(b));
    }
  }
  double s = 0;
  for (int j = 0; j < 8; ++j) s += f[j];
  for (int j = 0; j < 4; ++j) s += c[j][0] + c[j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ND, int NM>
void run(double* out, int ctas_per_sm) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int blocks = 148 * ctas_per_sm, iters = 4000;
  mix<ND, NM><<<blocks, 128>>>(out, 10, 1.0000001, 1e-9); cudaDeviceSynchronize();
  cudaEventRecord(e0); mix<ND, NM><<<blocks, 128>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double warps = blocks * 4.0, per = 4.0 * iters;
  double fd = 2.0 * 32 * ND * per * warps, fm = 512.0 * NM * per * warps;
  // cycles per warp-iteration per scheduler: 148*4 schedulers
  double cyc = ms * 1e-3 * 1.965e9 / (per * warps / (148.0 * 4));
  printf("ctas/SM %d  ND %2d NM %2d: %.3f ms  DFMA %.2f TF + DMMA %.2f TF = %.2f TF   %.1f scheduler-cycles per (ND DFMA + NM DMMA)\n", ctas_per_sm, ND, NM, ms, fd / ms / 1e9, fm / ms / 1e9, (fd + fm) / ms / 1e9, cyc);
}
int main() {
  double* out; cudaMalloc(&out, 148 * 16 * 128 * 8);
  for (int c : {1, 2, 4}) {
    run<16, 0>(out, c); run<0, 2>(out, c); run<16, 2>(out, c); run<8, 2>(out, c); run<32, 2>(out, c); run<16, 4>(out, c); run<0, 4>(out, c); run<32, 0>(out, c);
  }
  return 0;
}
