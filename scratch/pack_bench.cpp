// host-only: the staging team's per-row work (pack 39 values + narrow 39 ids per row) with T threads, AVX2 vs AVX-512
#include "../nimfm_b200/csrc/host_stage.cpp"
#include <chrono>
#include <cstdio>
#include <thread>
#include <vector>
int main(int argc, char **argv) {
  const size_t rows = 4u << 20, z = 39, n = rows * z;
  std::vector<double> val(n);
  std::vector<int64_t> idx(n);
  for (size_t i = 0; i < n; i++) { val[i] = (i % z) < 13 ? 0.25 + (i & 1023) * 1e-3 : 1.0; idx[i] = (int64_t)(i * 2654435761u % 1000000); }
  double *out = (double *)aligned_alloc(64, n * 8);
  int32_t *idst = (int32_t *)aligned_alloc(64, n * 4);
  std::vector<uint64_t> mask(n / 64 + 8);
  std::vector<uint32_t> blk(n / 256 + 8);
  for (size_t i = 0; i < n; i++) { out[i] = 0; idst[i] = 0; }
  for (int T : {4, 8, 12, 16}) {
    for (int rep = 0; rep < 3; rep++) {
      auto t0 = std::chrono::steady_clock::now();
      std::vector<std::thread> th;
      std::vector<int64_t> cnt(T);
      for (int t = 0; t < T; t++)
        th.emplace_back([&, t] {
          size_t a = (n / T * t) & ~(size_t)255, b = t == T - 1 ? n : (n / T * (t + 1)) & ~(size_t)255;
          cnt[t] = pack(val.data() + a, b - a, out + a, mask.data() + a / 64, blk.data() + a / 256, 0);
          narrow(idx.data() + a, idst + a, b - a, 1000000);
        });
      for (auto &x : th) x.join();
      double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (rep == 2) printf("avx512=%d T=%2d: %.1f M rows/s (%.1f GB/s read)  packed %lld\n", (int)use_avx512(), T, rows / s / 1e6, n * 16 / s / 1e9, (long long)cnt[0]);
    }
  }
}
