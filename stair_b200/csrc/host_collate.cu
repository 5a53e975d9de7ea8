// Host-side collate staging (no GPU work): the row copies of `collate_fn` / `to_device` (video_nmn/dataset.py:463-476) for a whole
// batch in one call.  The reference moves one question at a time (batch_size = 1); a 4096-question batch is 4096 dtype-converting
// copies of 131 KB (fp32 dataset tensors -> bf16 pinned staging), which through torch is one dispatcher round trip each
// (0.35-0.45 s per batch, mostly interpreter).  Here a small thread pool converts every row block straight into the pinned buffer.
#include "stair_common.cuh"

#include <atomic>
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include <cstring>
#include <thread>
#include <vector>

namespace stair {
namespace {

// fp32 -> bf16, round to nearest even: bit-identical to torch's conversion for every non-NaN input (incl. denormals, infinities, the
// carry into the exponent); NaN -> the canonical quiet NaN 0x7FC0 (torch's own result for a NaN depends on whether its vectorised or
// its scalar path converted the element: 0xFFFF or 0x7FC0)
inline uint16_t f32_to_bf16(uint32_t u) {
    if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7FC0;
    return static_cast<uint16_t>((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
}

// Vector forms of the same conversion (same bits as f32_to_bf16 for every input), chosen at run time: the library is built on one machine
// and runs on another.  The scalar loop converts ~0.3 G elements/s per thread (the NaN branch keeps it scalar); these are memory-bound.
#if defined(__x86_64__) && defined(__GNUC__)
#define STAIR_X86_DISPATCH 1
__attribute__((target("avx512f,avx512bw"))) void f32_to_bf16_avx512(const uint32_t* s, uint16_t* d, long long n) {
    const __m512i absmask = _mm512_set1_epi32(0x7fffffff), inf = _mm512_set1_epi32(0x7f800000), bias = _mm512_set1_epi32(0x7fff),
                  one = _mm512_set1_epi32(1), qnan = _mm512_set1_epi32(0x7FC0);
    long long i = 0;
    for (; i + 16 <= n; i += 16) {
        const __m512i v = _mm512_loadu_si512(s + i);
        const __mmask16 nan = _mm512_cmpgt_epu32_mask(_mm512_and_si512(v, absmask), inf);
        __m512i r = _mm512_srli_epi32(_mm512_add_epi32(_mm512_add_epi32(v, bias), _mm512_and_si512(_mm512_srli_epi32(v, 16), one)), 16);
        r = _mm512_mask_mov_epi32(r, nan, qnan);
        _mm256_storeu_si256(reinterpret_cast<__m256i*>(d + i), _mm512_cvtepi32_epi16(r));
    }
    for (; i < n; ++i) d[i] = f32_to_bf16(s[i]);
}
__attribute__((target("avx2"))) void f32_to_bf16_avx2(const uint32_t* s, uint16_t* d, long long n) {
    const __m256i absmask = _mm256_set1_epi32(0x7fffffff), inf = _mm256_set1_epi32(0x7f800000), bias = _mm256_set1_epi32(0x7fff),
                  one = _mm256_set1_epi32(1), qnan = _mm256_set1_epi32(0x7FC0);
    long long i = 0;
    for (; i + 16 <= n; i += 16) {
        __m256i r[2];
        for (int k = 0; k < 2; ++k) {
            const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i + 8 * k));
            const __m256i nan = _mm256_cmpgt_epi32(_mm256_and_si256(v, absmask), inf);      // both sides non-negative: the signed compare is exact
            const __m256i x = _mm256_srli_epi32(_mm256_add_epi32(_mm256_add_epi32(v, bias), _mm256_and_si256(_mm256_srli_epi32(v, 16), one)), 16);
            r[k] = _mm256_blendv_epi8(x, qnan, nan);
        }
        // packus works per 128-bit lane: [r0.lo r1.lo | r0.hi r1.hi] -> permute the 64-bit quarters back into element order
        const __m256i pk = _mm256_permute4x64_epi64(_mm256_packus_epi32(r[0], r[1]), 0xD8);
        _mm256_storeu_si256(reinterpret_cast<__m256i*>(d + i), pk);
    }
    for (; i < n; ++i) d[i] = f32_to_bf16(s[i]);
}
#endif

void f32_to_bf16_n(const uint32_t* s, uint16_t* d, long long n) {
#ifdef STAIR_X86_DISPATCH
    static const int level = __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512f") ? 2 : (__builtin_cpu_supports("avx2") ? 1 : 0);
    if (level == 2) return f32_to_bf16_avx512(s, d, n);
    if (level == 1) return f32_to_bf16_avx2(s, d, n);
#endif
    for (long long i = 0; i < n; ++i) d[i] = f32_to_bf16(s[i]);
}

void convert_block(const void* src, int src_dtype, void* dst, int dst_dtype, long long n) {
    if (src_dtype == dst_dtype) {
        std::memcpy(dst, src, static_cast<size_t>(n) * (src_dtype == STAIR_F32 ? 4 : 2));
    } else if (src_dtype == STAIR_F32) {                     // fp32 -> bf16
        f32_to_bf16_n(reinterpret_cast<const uint32_t*>(src), reinterpret_cast<uint16_t*>(dst), n);
    } else {                                                 // bf16 -> fp32
        const uint16_t* s = reinterpret_cast<const uint16_t*>(src);
        uint32_t* d = reinterpret_cast<uint32_t*>(dst);
        for (long long i = 0; i < n; ++i) d[i] = static_cast<uint32_t>(s[i]) << 16;
    }
}

}  // namespace
}  // namespace stair

// dst rows [dst_row[i], dst_row[i] + rows[i]) <- src[i] (rows[i] x cols elements, contiguous), converting src_dtype -> dst_dtype
// (STAIR_F32 / STAIR_BF16); dst is row-major with pitch `cols`.  `threads` host threads (<= 0: hardware concurrency, at most 32).
extern "C" int stair_host_collate_rows(const void* const* src, const long long* rows, const long long* dst_row, int n, long long cols,
                                       int src_dtype, void* dst, int dst_dtype, int threads) {
    using namespace stair;
    if (n <= 0) return STAIR_OK;
    if (!src || !rows || !dst_row || !dst || cols <= 0) return STAIR_ERR_ARG;
    if ((src_dtype != STAIR_F32 && src_dtype != STAIR_BF16) || (dst_dtype != STAIR_F32 && dst_dtype != STAIR_BF16)) return STAIR_ERR_ARG;
    int nt = threads > 0 ? threads : static_cast<int>(std::thread::hardware_concurrency());
    if (nt < 1) nt = 1;
    if (nt > 32) nt = 32;
    if (nt > n) nt = n;
    const size_t dsz = dst_dtype == STAIR_F32 ? 4 : 2;
    std::atomic<int> next(0);
    auto work = [&]() {
        for (;;) {
            const int i0 = next.fetch_add(8);                // 8 questions per grab: ~1 MB of source per grab at RX
            if (i0 >= n) break;
            const int i1 = i0 + 8 < n ? i0 + 8 : n;
            for (int i = i0; i < i1; ++i)
                convert_block(src[i], src_dtype, reinterpret_cast<char*>(dst) + static_cast<size_t>(dst_row[i]) * cols * dsz, dst_dtype, rows[i] * cols);
        }
    };
    if (nt == 1) { work(); return STAIR_OK; }
    std::vector<std::thread> pool;
    pool.reserve(nt - 1);
    try {
        for (int t = 0; t < nt - 1; ++t) pool.emplace_back(work);
    } catch (...) {
        // could not start (all) helpers: the calling thread finishes whatever is left
    }
    work();
    for (auto& th : pool) th.join();
    return STAIR_OK;
}
