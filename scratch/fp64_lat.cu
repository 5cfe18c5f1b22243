// micro-benchmark: dependent-issue latency and per-SM throughput of FP64 DFMA / sqrt / divide on one SM (one block)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void chain(double *out, long long *cyc, int n, double x) {
  double a = threadIdx.x * 1e-3, b = 1.0000001;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) a = fma(a, b, x);
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void chain4(double *out, long long *cyc, int n, double x) {
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b = 1.0000001;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { a0 = fma(a0, b, x); a1 = fma(a1, b, x); a2 = fma(a2, b, x); a3 = fma(a3, b, x); }
  long long t1 = clock64();
  out[threadIdx.x] = a0 + a1 + a2 + a3;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void chain_sqrt(double *out, long long *cyc, int n, double x) {
  double a = 2.0 + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) a = sqrt(a) + x;
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void chain_div(double *out, long long *cyc, int n, double x) {
  double a = 2.0 + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) a = x / a + 1.5;
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void chain_shfl(double *out, long long *cyc, int n) {
  double a = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) a += __shfl_xor_sync(0xffffffffu, a, 1 + (i & 15));
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void chain_f32(float *out, long long *cyc, int n, float x) {
  float a = threadIdx.x * 1e-3f, b = 1.0000001f;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) a = fmaf(a, b, x);
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double *out; long long *cyc, h;
  cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
  const int n = 4096;
  for (int threads : {32, 64, 128, 256, 512, 1024}) {
    chain<<<1, threads>>>(out, cyc, n, 1e-9); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("threads %4d  DFMA dependent: %.1f cyc/op", threads, (double)h / n);
    chain4<<<1, threads>>>(out, cyc, n, 1e-9); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("   4 chains: %.1f cyc/4 ops", (double)h / n);
    chain_sqrt<<<1, threads>>>(out, cyc, n, 1e-9); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("   sqrt+add: %.1f", (double)h / n);
    chain_div<<<1, threads>>>(out, cyc, n, 3.0); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("   div+add: %.1f", (double)h / n);
    chain_shfl<<<1, threads>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("   shfl64+add: %.1f", (double)h / n);
    chain_f32<<<1, threads>>>((float *)out, cyc, n, 1e-9f); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("   FFMA: %.1f\n", (double)h / n);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
