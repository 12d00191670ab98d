// Filling a pinned staging slot from the caller's pageable array with non-temporal stores.
//
// std::memcpy of a 4 MiB piece uses ordinary stores: every destination line is first read into the cache (read for
// ownership), so a copied byte costs three bytes of host-memory traffic, and on the B200 boxes' hosts the staging
// threads of host_staging.h saturate host memory (tools/c1_breakdown_probe.py: 18 GB/s for one thread, ~40 GB/s for
// four to six).  The slot is written once and then read by the GPU's copy engine, never by a CPU: streaming stores skip
// the ownership read.  Plain host code (no CUDA); compiled by the host compiler, AVX2 selected at run time.
#include <cstddef>
#include <cstdint>
#include <cstring>

#if defined(__x86_64__)
#include <immintrin.h>

namespace {
__attribute__((target("avx2"))) void fill_avx2(char* dst, const char* src, size_t bytes) {
    // dst: 32-byte aligned on entry (the caller peels the head)
    size_t i = 0;
    for (; i + 128 <= bytes; i += 128) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 32));
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 64));
        const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 96));
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), a);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 32), b);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 64), c);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 96), d);
    }
    if (i < bytes) std::memcpy(dst + i, src + i, bytes - i);
    _mm_sfence();        // the streamed lines must be globally visible before the slot is published to the copy engine
}
}  // namespace

extern "C" int aig_host_copy_streaming_supported() { return __builtin_cpu_supports("avx2") ? 1 : 0; }

extern "C" void aig_host_copy_streaming(void* dst, const void* src, size_t bytes) {
    char* d = static_cast<char*>(dst);
    const char* s = static_cast<const char*>(src);
    if (bytes < 4096 || !__builtin_cpu_supports("avx2")) {
        std::memcpy(d, s, bytes);
        return;
    }
    const size_t head = (32 - (reinterpret_cast<uintptr_t>(d) & 31)) & 31;
    if (head) { std::memcpy(d, s, head); d += head; s += head; bytes -= head; }
    fill_avx2(d, s, bytes);
}
#else
extern "C" int aig_host_copy_streaming_supported() { return 0; }
extern "C" void aig_host_copy_streaming(void* dst, const void* src, size_t bytes) { std::memcpy(dst, src, bytes); }
#endif
