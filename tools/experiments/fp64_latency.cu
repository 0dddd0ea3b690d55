// dependent-chain latencies of the fp64 operations the M2DP power iteration is made of (B200): clocks per operation
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double *out, long long *clk, double x, double y) {
  double a = x + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; i++) { a = fma(a, y, x); a = fma(a, y, x); a = fma(a, y, x); a = fma(a, y, x); }
  long long t1 = clock64();
  double b = a;
#pragma unroll 1
  for (int i = 0; i < 256; i++) { b += x; b += y; b += x; b += y; }
  long long t2 = clock64();
  double c = b;
#pragma unroll 1
  for (int i = 0; i < 256; i++) { c += __shfl_xor_sync(0xffffffffu, c, 1); c += __shfl_xor_sync(0xffffffffu, c, 2); }
  long long t3 = clock64();
  double d = c;
#pragma unroll 1
  for (int i = 0; i < 256; i++) { asm volatile("bar.sync 1, 128;" ::: "memory"); }
  long long t4 = clock64();
  __shared__ double sm[128];
  sm[threadIdx.x] = d;
#pragma unroll 1
  for (int i = 0; i < 256; i++) { d = fma(sm[(threadIdx.x + i) & 127], y, d); }
  long long t5 = clock64();
  float f = (float)d;
#pragma unroll 1
  for (int i = 0; i < 256; i++) { f = rsqrtf(f) + 1.0f; }
  long long t6 = clock64();
  out[threadIdx.x] = d + f;
  if (threadIdx.x == 0) {
    clk[0] = (t1 - t0) / 1024; clk[1] = (t2 - t1) / 1024; clk[2] = (t3 - t2) / 512; clk[3] = (t4 - t3) / 256;
    clk[4] = (t5 - t4) / 256; clk[5] = (t6 - t5) / 256;
  }
}
int main() {
  double *o; long long *c, h[6];
  cudaMalloc(&o, 128 * 8); cudaMalloc(&c, 6 * 8);
  for (int r = 0; r < 2; r++) lat<<<1, 128>>>(o, c, 1.0000001, 0.9999999);
  cudaMemcpy(h, c, sizeof(h), cudaMemcpyDeviceToHost);
  printf("DFMA %lld  DADD %lld  shfl64+DADD %lld  bar.sync(128) %lld  LDS+DFMA %lld  rsqrtf+FADD %lld  clk\n", h[0], h[1], h[2], h[3], h[4], h[5]);
  return 0;
}
