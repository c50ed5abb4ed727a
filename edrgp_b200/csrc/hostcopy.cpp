// Host-side copy used by the row streamer (h2d.cu) to fill its pinned blocks: non-temporal stores, so that the
// destination lines are not read before they are overwritten.  Staging 2 GB of pageable rows moves 2 GB in, 2 GB out
// and 2 GB again when the copy engine reads the pinned block; with ordinary stores the write-allocate adds a fourth
// pass, and the 16-core host of a B200 box runs out of memory bandwidth before it runs out of cores.
#include <cstddef>
#include <cstdint>
#include <cstring>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace edrgp {

#if defined(__x86_64__)
__attribute__((target("avx2"))) static void copy_nt_avx2(char* dst, const char* src, size_t n) {
  size_t i = 0;
  // destination stripes start on 64-byte boundaries (page-aligned slots, 64-byte stripe granularity)
  for (; i + 128 <= n; i += 128) {
    const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i));
    const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 32));
    const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 64));
    const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 96));
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), a);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 32), b);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 64), c);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 96), d);
  }
  if (i < n) std::memcpy(dst + i, src + i, n - i);
  _mm_sfence();
}
#endif

void host_copy_streaming(void* dst, const void* src, size_t n) {
#if defined(__x86_64__)
  static const bool avx2 = __builtin_cpu_supports("avx2");
  if (avx2 && n >= 4096 && (reinterpret_cast<uintptr_t>(dst) & 31u) == 0) {
    copy_nt_avx2(static_cast<char*>(dst), static_cast<const char*>(src), n);
    return;
  }
#endif
  std::memcpy(dst, src, n);
}

}  // namespace edrgp
