// issue throughput of fp64 vector instructions on one SM (B200): clocks per warp-instruction per SM sub-partition
#include <cstdio>
#include <cuda_runtime.h>
template <int KIND>
__global__ void thr(double *out, long long *clk, double x, double y) {
  double a0 = x + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 128; i++) {
    if (KIND == 0) { a0 = fma(a0, y, x); a1 = fma(a1, y, x); a2 = fma(a2, y, x); a3 = fma(a3, y, x); a4 = fma(a4, y, x); a5 = fma(a5, y, x); a6 = fma(a6, y, x); a7 = fma(a7, y, x); }
    if (KIND == 1) { a0 += x; a1 += x; a2 += x; a3 += x; a4 += x; a5 += x; a6 += x; a7 += x; }
    if (KIND == 2) { a0 *= y; a1 *= y; a2 *= y; a3 *= y; a4 *= y; a5 *= y; a6 *= y; a7 *= y; }
  }
  __syncthreads();
  long long t1 = clock64();
  out[threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
int main() {
  double *o; long long *c, h;
  cudaMalloc(&o, 1024 * 8); cudaMalloc(&c, 8);
  const char *names[3] = {"DFMA", "DADD", "DMUL"};
  for (int kind = 0; kind < 3; kind++)
    for (int threads = 128; threads <= 1024; threads *= 2) {
      for (int r = 0; r < 2; r++) {
        if (kind == 0) thr<0><<<1, threads>>>(o, c, 1.0000001, 0.9999999);
        if (kind == 1) thr<1><<<1, threads>>>(o, c, 1.0000001, 0.9999999);
        if (kind == 2) thr<2><<<1, threads>>>(o, c, 1.0000001, 0.9999999);
      }
      cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
      const double per_smsp = (threads / 32 / 4.0) * 128 * 8;   // warp-instructions per sub-partition
      printf("%s %4d threads: %lld clk, %.2f clk per warp-instruction per SMSP\n", names[kind], threads, h, h / per_smsp);
    }
  return 0;
}
