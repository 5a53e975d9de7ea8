// Host-side collate staging (no GPU work): the row copies of `collate_fn` / `to_device` (video_nmn/dataset.py:463-476) for a whole
// batch in one call.  The reference moves one question at a time (batch_size = 1); a 4096-question batch is 4096 dtype-converting
// copies of 131 KB (fp32 dataset tensors -> bf16 pinned staging), which through torch is one dispatcher round trip each
// (0.35-0.45 s per batch, mostly interpreter).  Here a small thread pool converts every row block straight into the pinned buffer.
#include "stair_common.cuh"

#include <atomic>
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

void convert_block(const void* src, int src_dtype, void* dst, int dst_dtype, long long n) {
    if (src_dtype == dst_dtype) {
        std::memcpy(dst, src, static_cast<size_t>(n) * (src_dtype == STAIR_F32 ? 4 : 2));
    } else if (src_dtype == STAIR_F32) {                     // fp32 -> bf16
        const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
        uint16_t* d = reinterpret_cast<uint16_t*>(dst);
        for (long long i = 0; i < n; ++i) d[i] = f32_to_bf16(s[i]);
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
