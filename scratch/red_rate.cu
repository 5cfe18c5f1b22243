// micro-benchmark: what the L2 does with FP64 REDs.  Random SEG-byte segments (256 B = one warp-wide
// RED of 32 doubles, the FM row kernel's pattern; 64 B = four 8-lane segments per warp instruction, the
// FFM pair kernel's) in a buffer that fits L2 (48 MB) or does not (1 GB); the same address stream as
// RED.ADD.F64, as plain stores, as loads and as RED.ADD.F32 pairs, U = 8 independent operations in
// flight per lane.  Prints GB/s of operand bytes and bytes per SM clock.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

template <int OP, int SEG>
__global__ void __launch_bounds__(256) k(double *buf, uint32_t nSeg, int iters, double *sink) {
  constexpr int LPS = SEG / 8;               // lanes per segment
  const int lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t sub = lane / LPS;
  double acc = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const uint32_t h = mix((warp * 977u + it) * 64u + u * 4u + sub);
      const uint32_t s = (uint32_t)(((uint64_t)h * nSeg) >> 32);
      double *p = buf + (size_t)s * LPS + (lane % LPS);
      if (OP == 0) atomicAdd(p, 1.0);
      else if (OP == 1) *p = 1.0 + it;
      else if (OP == 2) acc += __ldg(p);
      else if (OP == 3) { float *f = reinterpret_cast<float *>(p); atomicAdd(f, 1.0f); atomicAdd(f + 1, 1.0f); }
      else if (OP == 4) acc += atomicAdd(p, 1.0);   // ATOM (returns): round trip
    }
  }
  if (OP == 2 || OP == 4) if (acc == 123.456) sink[0] = acc;
}

template <int OP, int SEG>
static void run(const char *name, double *buf, size_t bytes, double *sink, int clkMHz) {
  const int blocks = 148 * 8, threads = 256, iters = 256;
  const uint32_t nSeg = (uint32_t)(bytes / SEG);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<OP, SEG><<<blocks, threads>>>(buf, nSeg, 16, sink);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<OP, SEG><<<blocks, threads>>>(buf, nSeg, iters, sink);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double ops = (double)blocks * threads * iters * 8;          // lane operations
  const double GBs = ops * 8 / (ms * 1e-3) / 1e9;
  printf("%-28s seg %3d B  buf %5zu MB  %8.3f ms  %8.1f GB/s  %7.1f B/clk (at %d MHz)\n", name, SEG,
         bytes >> 20, ms, GBs, GBs * 1e9 / (clkMHz * 1e6), clkMHz);
}

int main() {
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0); clk /= 1000;
  double *buf, *sink; const size_t big = (size_t)1 << 30;
  cudaMalloc(&buf, big); cudaMalloc(&sink, 8); cudaMemset(buf, 0, big);
  for (size_t bytes : {(size_t)48 << 20, big}) {
    run<0, 256>("RED.ADD.F64", buf, bytes, sink, clk);
    run<0, 64>("RED.ADD.F64", buf, bytes, sink, clk);
    run<4, 256>("ATOM.ADD.F64 (returning)", buf, bytes, sink, clk);
    run<3, 256>("RED.ADD.F32 x2", buf, bytes, sink, clk);
    run<1, 256>("store", buf, bytes, sink, clk);
    run<1, 64>("store", buf, bytes, sink, clk);
    run<2, 256>("load (nc)", buf, bytes, sink, clk);
    run<2, 64>("load (nc)", buf, bytes, sink, clk);
  }
  return 0;
}
