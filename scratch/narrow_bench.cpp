#include <immintrin.h>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
static int narrow(const int64_t *src, int32_t *dst, size_t n, int64_t d) {
  size_t i = 0; int64_t bad = 0;
#ifdef __AVX2__
  const __m256i perm = _mm256_setr_epi32(0, 2, 4, 6, 0, 0, 0, 0);
  __m256i acc = _mm256_setzero_si256();
  for (; i + 8 <= n; i += 8) {
    __m256i a = _mm256_loadu_si256((const __m256i *)(src + i)), b = _mm256_loadu_si256((const __m256i *)(src + i + 4));
    acc = _mm256_or_si256(acc, _mm256_or_si256(a, b));
    __m128i lo = _mm256_castsi256_si128(_mm256_permutevar8x32_epi32(a, perm));
    __m128i hi = _mm256_castsi256_si128(_mm256_permutevar8x32_epi32(b, perm));
    _mm256_stream_si256((__m256i *)(dst + i), _mm256_set_m128i(hi, lo));
  }
  alignas(32) int64_t t[4]; _mm256_store_si256((__m256i *)t, acc); bad = t[0] | t[1] | t[2] | t[3];
#endif
  for (; i < n; i++) { bad |= src[i]; dst[i] = (int32_t)src[i]; }
  return (bad < 0 || bad >= d);
}
int main(int argc, char **argv) {
  int T = argc > 1 ? atoi(argv[1]) : 4;
  size_t n = 200u << 20;
  std::vector<int64_t> src(n);
  int32_t *dst = (int32_t *)aligned_alloc(64, n * 4);
  for (size_t i = 0; i < n; i++) src[i] = i & 0xfffff;
  for (size_t i = 0; i < n; i++) dst[i] = 0;
  for (int rep = 0; rep < 3; rep++) {
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++) th.emplace_back([&, t] { size_t a = n / T * t, b = t == T - 1 ? n : n / T * (t + 1); narrow(src.data() + a, dst + a, b - a, 1 << 20); });
    for (auto &x : th) x.join();
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("T=%d %.2f Gelem/s (%.1f GB/s read)\n", T, n / s / 1e9, n * 8 / s / 1e9);
  }
}
