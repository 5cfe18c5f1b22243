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

// TMA bulk reduction: the warp writes SEG bytes into one of NBUF shared-memory buffers, one lane issues
// cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 of the whole segment
template <int SEG, int NBUF>
__global__ void __launch_bounds__(256) kbulk(double *buf, uint32_t nSeg, int iters) {
  __shared__ __align__(128) double sb[8][NBUF][SEG / 8];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  for (int it = 0; it < iters * 8; ++it) {
    const int b = it % NBUF;
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NBUF - 1) : "memory");
    __syncwarp();
    for (int e = lane; e < SEG / 8; e += 32) sb[wib][b][e] = 1.0;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      const uint32_t h = mix((warp * 977u + (it >> 3)) * 64u + (it & 7) * 4u);
      const uint32_t s = (uint32_t)(((uint64_t)h * nSeg) >> 32);
      double *dst = buf + (size_t)s * (SEG / 8);
      const uint32_t src = (uint32_t)__cvta_generic_to_shared(&sb[wib][b][0]);
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(dst), "r"(src), "n"(SEG) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int SEG, int NBUF>
static void runbulk(double *buf, size_t bytes, int clkMHz) {
  const int blocks = 148 * 8, threads = 256, iters = 256;
  const uint32_t nSeg = (uint32_t)(bytes / SEG);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  kbulk<SEG, NBUF><<<blocks, threads>>>(buf, nSeg, 16);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  kbulk<SEG, NBUF><<<blocks, threads>>>(buf, nSeg, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double bytesMoved = (double)blocks * (threads / 32) * iters * 8 * SEG;
  const double GBs = bytesMoved / (ms * 1e-3) / 1e9;
  printf("%-28s seg %3d B  buf %5zu MB  %8.3f ms  %8.1f GB/s  %7.1f B/clk (at %d MHz)  [%d buffers/warp] %s\n",
         "TMA bulk reduce .add.f64", SEG, bytes >> 20, ms, GBs, GBs * 1e9 / (clkMHz * 1e6), clkMHz, NBUF,
         cudaGetErrorString(cudaGetLastError()));
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
    runbulk<256, 4>(buf, bytes, clk);
    runbulk<512, 4>(buf, bytes, clk);
    runbulk<512, 2>(buf, bytes, clk);
    runbulk<2048, 2>(buf, bytes, clk);
  }
  return 0;
}
