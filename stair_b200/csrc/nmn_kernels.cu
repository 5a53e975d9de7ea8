// Memory-bound kernels of the NMN hot path: temporal-attention maps, relate/scan, gating, reductions, answer heads.
// Each kernel states the reference lines it implements (video_nmn/modules.py, video_nmn/module_net.py) and reads its
// [.., H] rows with 16-byte vector loads (8 x bf16 / 2 x float4), one warp (or a few lanes) per row.
#include "nmn_kernels.cuh"
#include <mutex>

namespace stair {

#define DISPATCH_DT(dt, AT, ...)                                   \
    do {                                                           \
        if ((dt) == STAIR_BF16) { typedef bf16 AT; __VA_ARGS__; }  \
        else { typedef float AT; __VA_ARGS__; }                    \
    } while (0)

static inline int blocks_for(long long work, int per_block) {
    long long b = (work + per_block - 1) / per_block;
    return static_cast<int>(b < 1 ? 1 : b);
}

// three-way bf16 split of an fp32 value: x ~= p0 + p1 + p2 (each exactly representable in bf16)
__device__ __forceinline__ void split3(float x, bf16& p0, bf16& p1, bf16& p2) {
    p0 = __float2bfloat16_rn(x);
    float r = x - __bfloat162float(p0);
    p1 = __float2bfloat16_rn(r);
    r -= __bfloat162float(p1);
    p2 = __float2bfloat16_rn(r);
}

// write 8 fp32 values as 1 or 3 bf16 planes (16-byte stores)
__device__ __forceinline__ void store_planes8(const float (&v)[8], bf16* dst, long long plane_stride, int nplanes) {
    if (nplanes == 1) {
        Vec8<bf16> o;
#pragma unroll
        for (int i = 0; i < 8; ++i) o.v[i] = v[i];
        o.store(dst);
    } else {
        uint4 r0, r1, r2;
        bf16* h0 = reinterpret_cast<bf16*>(&r0); bf16* h1 = reinterpret_cast<bf16*>(&r1); bf16* h2 = reinterpret_cast<bf16*>(&r2);
#pragma unroll
        for (int i = 0; i < 8; ++i) split3(v[i], h0[i], h1[i], h2[i]);
        *reinterpret_cast<uint4*>(dst) = r0;
        *reinterpret_cast<uint4*>(dst + plane_stride) = r1;
        *reinterpret_cast<uint4*>(dst + 2 * plane_stride) = r2;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// GEMM operand staging
// ------------------------------------------------------------------------------------------------------------------
template <typename ST>
__global__ void stage_rows_kernel(const ST* __restrict__ src, long long ld_src, const int* __restrict__ slots, int rps, int unit,
                                  bf16* __restrict__ dst, long long ld_dst, long long plane_rows, int nplanes, long long rows, int cols) {
    const int chunks = static_cast<int>(ld_dst / 8);
    const long long total = rows * chunks;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / chunks;
        const int c = static_cast<int>(i % chunks) * 8;
        const long long sr = slots ? static_cast<long long>(__ldg(slots + r / rps)) * unit + r % rps : r;
        float v[8];
        if (c + 8 <= cols && (ld_src % 8) == 0) {
            Vec8<ST> x; x.load(src + sr * ld_src + c);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = x.v[j];
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = (c + j < cols) ? ld1<ST>(src + sr * ld_src + c + j) : 0.f;
        }
        store_planes8(v, dst + r * ld_dst + c, plane_rows * ld_dst, nplanes);
    }
}

int launch_stage_rows(int sdt, const void* src, long long ld_src, const int* slots, int rps, int unit, bf16* dst, long long ld_dst,
                      long long plane_rows, int nplanes, long long rows, int cols, cudaStream_t st) {
    if (rows <= 0) return STAIR_OK;
    if (ld_dst % 8) return STAIR_ERR_ARG;
    const int grid = min(blocks_for(rows * (ld_dst / 8), 256), 148 * 16);
    DISPATCH_DT(sdt, ST, (stage_rows_kernel<ST><<<grid, 256, 0, st>>>(reinterpret_cast<const ST*>(src), ld_src, slots, rps < 1 ? 1 : rps, unit,
                                                                     dst, ld_dst, plane_rows, nplanes, rows, cols)));
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

template <typename AT>
__global__ void concat_vec_kernel(const AT* __restrict__ vec, const int* __restrict__ a_idx, const int* __restrict__ b_idx, int mode,
                                  bf16* __restrict__ dst, long long plane_rows, int nplanes, int n, int H) {
    const int hc = H / 8;
    const int width = (mode == STAIR_CAT_PAIR ? 2 : 3) * H;
    const long long total = static_cast<long long>(n) * hc;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int r = static_cast<int>(i / hc), c = static_cast<int>(i % hc) * 8;
        Vec8<AT> a, b;
        a.load(vec + static_cast<long long>(__ldg(a_idx + r)) * H + c);
        b.load(vec + static_cast<long long>(__ldg(b_idx + r)) * H + c);
        bf16* d = dst + static_cast<long long>(r) * width + c;
        const long long ps = plane_rows * width;
        float t[8];
        if (mode == STAIR_CAT_EXISTS) {            // [feat | keyword | feat*keyword], keyword = arg0, feat = arg1
            store_planes8(b.v, d, ps, nplanes);
            store_planes8(a.v, d + H, ps, nplanes);
#pragma unroll
            for (int j = 0; j < 8; ++j) t[j] = b.v[j] * a.v[j];
            store_planes8(t, d + 2 * H, ps, nplanes);
        } else if (mode == STAIR_CAT_XOR) {        // [|f1-f2| | f1 | f2]
#pragma unroll
            for (int j = 0; j < 8; ++j) t[j] = fabsf(a.v[j] - b.v[j]);
            store_planes8(t, d, ps, nplanes);
            store_planes8(a.v, d + H, ps, nplanes);
            store_planes8(b.v, d + 2 * H, ps, nplanes);
        } else {                                   // [a | b]
            store_planes8(a.v, d, ps, nplanes);
            store_planes8(b.v, d + H, ps, nplanes);
        }
    }
}

int launch_concat_vec(int dt, const void* vec, const int* a_idx, const int* b_idx, int mode, bf16* dst, long long plane_rows,
                      int nplanes, int n, int H, cudaStream_t st) {
    if (n <= 0) return STAIR_OK;
    const int grid = min(blocks_for(static_cast<long long>(n) * (H / 8), 256), 148 * 16);
    DISPATCH_DT(dt, AT, (concat_vec_kernel<AT><<<grid, 256, 0, st>>>(reinterpret_cast<const AT*>(vec), a_idx, b_idx, mode, dst, plane_rows, nplanes, n, H)));
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

template <typename AT>
__global__ void decoder_concat_kernel(const AT* __restrict__ vec, const int* __restrict__ root_node, const int* __restrict__ out_slot, const AT* __restrict__ qfeat,
                                      bf16* __restrict__ dst, long long plane_rows, int nplanes, int B, int H) {
    const int hc = H / 8;
    const long long total = static_cast<long long>(B) * hc * 2;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int b = static_cast<int>(i / (2 * hc));
        const int c = static_cast<int>(i % (2 * hc)) * 8;
        Vec8<AT> x;
        if (c < H) x.load(vec + static_cast<long long>(out_slot[__ldg(root_node + b)]) * H + c);
        else x.load(qfeat + static_cast<long long>(b) * H + (c - H));
        store_planes8(x.v, dst + static_cast<long long>(b) * 2 * H + c, plane_rows * 2 * H, nplanes);
    }
}

int launch_decoder_concat(int dt, const void* vec, const int* root_node, const int* out_slot, const void* qfeat, bf16* dst, long long plane_rows,
                          int nplanes, int B, int H, cudaStream_t st) {
    if (B <= 0) return STAIR_OK;
    const int grid = min(blocks_for(static_cast<long long>(B) * (H / 4), 256), 148 * 16);
    DISPATCH_DT(dt, AT, (decoder_concat_kernel<AT><<<grid, 256, 0, st>>>(reinterpret_cast<const AT*>(vec), root_node, out_slot,
                                                                         reinterpret_cast<const AT*>(qfeat), dst, plane_rows, nplanes, B, H)));
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// phrase embedding: mean(token_feature[s:e], 0)            video_nmn/module_net.py:126-131
// ------------------------------------------------------------------------------------------------------------------
template <typename AT>
__global__ void word_embed_kernel(const AT* __restrict__ tokfeat, const int* __restrict__ q_off, const int* __restrict__ pos_q,
                                  const int* __restrict__ span_s, const int* __restrict__ span_e, AT* __restrict__ vec, int out_base,
                                  int n, int H) {
    const int hc = H / 8;
    const long long total = static_cast<long long>(n) * hc;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int r = static_cast<int>(i / hc), c = static_cast<int>(i % hc) * 8;
        const int q = __ldg(pos_q + r);
        const int base = __ldg(q_off + q), L = __ldg(q_off + q + 1) - base;
        int s = __ldg(span_s + r), e = __ldg(span_e + r);
        if (s < 0) { s = 0; e = L; }                 // (None, None) span: token_feature[None:None] = whole question
        s = min(s, L); e = min(e, L);                // python slice clamping
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int t = s; t < e; ++t) {
            Vec8<AT> x; x.load(tokfeat + static_cast<long long>(base + t) * H + c);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += x.v[j];
        }
        const float cnt = static_cast<float>(e - s);   // empty slice -> 0/0 = NaN exactly like torch.mean of an empty tensor
        Vec8<AT> o;
#pragma unroll
        for (int j = 0; j < 8; ++j) o.v[j] = acc[j] / cnt;
        o.store(vec + static_cast<long long>(out_base + r) * H + c);
    }
}

int launch_word_embed(int dt, const void* tokfeat, const int* q_off, const int* pos_q, const int* span_s, const int* span_e,
                      void* vec, int out_base, int n, int H, cudaStream_t st) {
    if (n <= 0) return STAIR_OK;
    const int grid = min(blocks_for(static_cast<long long>(n) * (H / 8), 256), 148 * 16);
    DISPATCH_DT(dt, AT, (word_embed_kernel<AT><<<grid, 256, 0, st>>>(reinterpret_cast<const AT*>(tokfeat), q_off, pos_q, span_s, span_e,
                                                                     reinterpret_cast<AT*>(vec), out_base, n, H)));
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// cosine attention maps                        Localize modules.py:205-216, ExistsFrame modules.py:170-177
// nn.CosineSimilarity: each side divided by max(norm, 1e-8), then dotted.  One warp per frame row.
// ------------------------------------------------------------------------------------------------------------------
// Row kernels below keep RPW rows of CH 16-byte chunks per lane in flight per warp, in storage format (Raw8; H = 512 -> 8 rows
// x 2 chunks): all loads of an iteration are issued before the first warp reduction, so a warp has 8-16 independent 16-byte
// requests outstanding instead of 2 (one row at a time was latency-bound at 11-28 % of the HBM roofline,
// profiles/r1_module_kernel_roofline.txt).
constexpr int MAXC = 8;     // vec8 chunks per lane: rows up to H = 32*8*8 = 2048
#define DISPATCH_CH(H, ...)                                                                   \
    do {                                                                                      \
        const int hc__ = (H) / 8;                                                             \
        if (hc__ <= 32) { constexpr int CH = 1, RPW = 8; __VA_ARGS__; }                       \
        else if (hc__ <= 64) { constexpr int CH = 2, RPW = 4; __VA_ARGS__; }                  \
        else if (hc__ <= 128) { constexpr int CH = 4, RPW = 2; __VA_ARGS__; }                 \
        else { constexpr int CH = 8, RPW = 1; __VA_ARGS__; }                                  \
    } while (0)
// the single-pass HasItem tail: measured at H = 512 with the grid sized to the resident blocks, 8 / 4 / 2 rows per warp and pass reach
// 59.6 / 70.9 / 78.7 % of the HBM roofline (fewer registers -> more resident warps, finer interleaving of loads and reductions);
// the two-pass kernels (cosine maps, layernorm) measured the other way round (R = 4 -> 2: 68 -> 57 %, 82 -> 72 %)
#define DISPATCH_CH16(H, ...)                                                                 \
    do {                                                                                      \
        const int hc__ = (H) / 8;                                                             \
        if (hc__ <= 32) { constexpr int CH = 1, RPW = 4; __VA_ARGS__; }                       \
        else if (hc__ <= 64) { constexpr int CH = 2, RPW = 2; __VA_ARGS__; }                  \
        else if (hc__ <= 128) { constexpr int CH = 4, RPW = 4; __VA_ARGS__; }                 \
        else { constexpr int CH = 8, RPW = 2; __VA_ARGS__; }                                  \
    } while (0)

// Row kernels are persistent: blocks of 8 warps, each warp RPW rows per grid-stride iteration, and never more blocks than are RESIDENT
// at once (148 SMs x the kernel's true occupancy, queried once per instantiation).  Sizing the grid for 8 blocks per SM when the
// registers allow 2-3 ran 1024 blocks as 2.3-3.5 waves of 296-444 — the last wave with half the machine idle (rowdot 58 %, cosine
// maps 67-69 % of the HBM roofline at streaming sizes).  With one resident sweep the warps differ by at most one iteration.
static int kernel_occ(const void* fn) {
    struct Entry { const void* fn; int occ; };
    static Entry cache[64];
    static int n = 0;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    for (int i = 0; i < n; ++i) if (cache[i].fn == fn) return cache[i].occ;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, 256, 0) != cudaSuccess || occ < 1) { cudaGetLastError(); occ = 1; }
    if (occ > 8) occ = 8;
    if (n < 64) { cache[n].fn = fn; cache[n].occ = occ; ++n; }
    return occ;
}
template <typename KF>
static inline int row_grid(KF fn, long long rows, int rpw) {
    long long b = (rows + 8LL * rpw - 1) / (8LL * rpw);
    if (b < 1) b = 1;
    const long long cap = 148LL * kernel_occ(reinterpret_cast<const void*>(fn));
    return static_cast<int>(b > cap ? cap : b);
}

// att[(out_base + inst*K + k)*T + t] = (cos(f_row, kw_row) + 1) * 0.49 for RPW frame rows at a time.
// f rows: contiguous [rows, H] (feat_idx == null) or gathered VID slots; keyword rows: contiguous [n*K, H] (kw_idx == null) or VEC rows.
template <typename AT, int CH, int RPW>
__global__ void __launch_bounds__(256, 2) cos_rows_kernel(const AT* __restrict__ f, const int* __restrict__ feat_idx, const AT* __restrict__ kmat,
                                const int* __restrict__ kw_idx, int K, int T, int H, float* __restrict__ att, long long out_base, int rows) {
    const int lane = threadIdx.x & 31, hc = H / 8;
    const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    for (int row0 = warp * RPW; row0 < rows; row0 += nwarps * RPW) {
        Raw8<AT> x[RPW][CH];
        int inst[RPW];
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const int row = row0 + r < rows ? row0 + r : rows - 1;
            inst[r] = row / T;
            const int t = row - inst[r] * T;
            const AT* fr = feat_idx ? f + (static_cast<long long>(__ldg(feat_idx + inst[r])) * T + t) * H : f + static_cast<long long>(row) * H;
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                const int c = lane + 32 * i;
                if (c < hc) x[r][i].load(fr + c * 8); else x[r][i].zero();
            }
        }
        for (int k = 0; k < K; ++k) {
            Raw8<AT> y[RPW][CH];
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                const AT* kr = kw_idx ? kmat + static_cast<long long>(__ldg(kw_idx + inst[r])) * H : kmat + (static_cast<long long>(inst[r]) * K + k) * H;
#pragma unroll
                for (int i = 0; i < CH; ++i) {
                    const int c = lane + 32 * i;
                    if (c < hc) y[r][i].load(kr + c * 8); else y[r][i].zero();
                }
            }
            float dot[RPW], kk[RPW], ff[RPW];
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                float d = 0.f, q = 0.f, s2 = 0.f;
#pragma unroll
                for (int i = 0; i < CH; ++i) {
                    float xv[8], yv[8];
                    x[r][i].unpack(xv); y[r][i].unpack(yv);
#pragma unroll
                    for (int j = 0; j < 8; ++j) { d += xv[j] * yv[j]; q += yv[j] * yv[j]; s2 += xv[j] * xv[j]; }
                }
                dot[r] = d; kk[r] = q; ff[r] = s2;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int r = 0; r < RPW; ++r) {
                    dot[r] += __shfl_xor_sync(0xffffffffu, dot[r], o);
                    kk[r] += __shfl_xor_sync(0xffffffffu, kk[r], o);
                    ff[r] += __shfl_xor_sync(0xffffffffu, ff[r], o);
                }
            }
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                if (lane == r && row0 + r < rows) {                       // lane r writes row r: RPW scattered 4-byte stores in one instruction
                    const float nf = fmaxf(sqrtf(ff[r]), 1e-8f), nk = fmaxf(sqrtf(kk[r]), 1e-8f);
                    const int t = (row0 + r) - inst[r] * T;
                    att[(out_base + static_cast<long long>(inst[r]) * K + k) * T + t] = (dot[r] / (nf * nk) + 1.0f) * 0.49f;
                }
            }
        }
    }
}

// Instance-major variant for T % 8 == 0: one warp owns a whole instance, R (8 / 4 / 2 by row width) frame rows in flight per pass.  The keyword row and its
// norm are loaded / reduced once per (instance, keyword) instead of once per frame row, |f_t|^2 once per frame instead of once per
// (frame, keyword), and the 8 results of a pass leave as one 32-byte store: ~1.75x fewer instructions per row than cos_rows_kernel,
// which was issue-bound at 44 % of the HBM roofline (profiles/r1_module_kernel_roofline.txt).
// Reduce V per-lane partial sums over the warp with V/2 + V/4 + ... + 1 + (5 - log2 V) shuffles instead of 5 V: each exchange step
// halves the number of live values per lane.  Returns, in every lane, the warp-wide sum of v[lane >> (5 - log2 V)].
template <int V>
__device__ __forceinline__ float warp_sum_multi(float (&v)[V], int lane) {
    static_assert(V == 2 || V == 4 || V == 8 || V == 16, "power of two");
    int bit = 16;
#pragma unroll
    for (int half = V / 2; half >= 1; half >>= 1, bit >>= 1) {
        const bool up = (lane & bit) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = up ? v[i] : v[i + half], keep = up ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
        }
    }
    for (; bit >= 1; bit >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], bit);
    return v[0];
}

template <typename AT, int CH, int R>
__global__ void __launch_bounds__(256, 2) cos_inst_kernel(const AT* __restrict__ f, const int* __restrict__ feat_idx, const AT* __restrict__ kmat,
                                                          const int* __restrict__ kw_idx, int K, int T, int H, float* __restrict__ att,
                                                          long long out_base, int n) {
    constexpr int V = 2 * R, SH = R == 8 ? 1 : (R == 4 ? 2 : 3);       // values per reduction; lane >> SH = index of the value a lane ends up with
    const int lane = threadIdx.x & 31, hc = H / 8;
    const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    const int my_r = (lane >> SH) & (R - 1);                            // frame row (within a pass) whose results this lane receives
    const bool writer = (lane & 16) == 0 && (lane & ((1 << SH) - 1)) == 0;
    for (int inst = warp; inst < n; inst += nwarps) {
        const AT* fbase = feat_idx ? f + static_cast<long long>(__ldg(feat_idx + inst)) * T * H : f + static_cast<long long>(inst) * T * H;
        float yv[CH][8];
        float nk = 1.f;
        for (int t0 = 0; t0 < T; t0 += R) {
            Raw8<AT> x[R][CH];
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int i = 0; i < CH; ++i) {
                    const int c = lane + 32 * i;
                    if (c < hc) x[r][i].load_stream(fbase + static_cast<long long>(t0 + r) * H + c * 8); else x[r][i].zero();
                }
            float ff_mine = 0.f;
            for (int k = 0; k < K; ++k) {
                if (K > 1 || t0 == 0) {                                // one keyword: its row and norm are computed once per instance
                    const AT* kr = kw_idx ? kmat + static_cast<long long>(__ldg(kw_idx + inst)) * H : kmat + (static_cast<long long>(inst) * K + k) * H;
                    float kk = 0.f;
#pragma unroll
                    for (int i = 0; i < CH; ++i) {
                        const int c = lane + 32 * i;
                        Raw8<AT> y;
                        if (c < hc) y.load(kr + c * 8); else y.zero();
                        y.unpack(yv[i]);
#pragma unroll
                        for (int j = 0; j < 8; ++j) kk = fmaf(yv[i][j], yv[i][j], kk);
                    }
                    nk = fmaxf(sqrtf(warp_sum(kk)), 1e-8f);
                }
                float v[V];                                             // v[r] = dot(f_r, k), v[R + r] = |f_r|^2 (first keyword only)
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    float d = 0.f, s2 = 0.f;
#pragma unroll
                    for (int i = 0; i < CH; ++i) {
                        float xv[8];
                        x[r][i].unpack(xv);
#pragma unroll
                        for (int j = 0; j < 8; ++j) { d = fmaf(xv[j], yv[i][j], d); if (k == 0) s2 = fmaf(xv[j], xv[j], s2); }
                    }
                    v[r] = d; v[R + r] = s2;
                }
                const float mine = warp_sum_multi<V>(v, lane);           // lanes < 16: dot of row my_r ; lanes >= 16: |f|^2 of row my_r
                const float other = __shfl_xor_sync(0xffffffffu, mine, 16);
                if (k == 0) ff_mine = other;
                if (writer) att[(out_base + static_cast<long long>(inst) * K + k) * T + t0 + my_r] = (mine / (fmaxf(sqrtf(ff_mine), 1e-8f) * nk) + 1.0f) * 0.49f;
            }
        }
    }
}

#define DISPATCH_CH_ONLY(H, ...)                                                              \
    do {                                                                                      \
        const int hc__ = (H) / 8;                                                             \
        if (hc__ <= 32) { constexpr int CH = 1, R = 8; __VA_ARGS__; }                         \
        else if (hc__ <= 64) { constexpr int CH = 2, R = 4; __VA_ARGS__; }                    \
        else { constexpr int CH = 4, R = 2; __VA_ARGS__; }                                    \
    } while (0)

int g_row_stream = 0;      // HasItem tail: 0 = register-staged kernel (76-79 % of the HBM roofline at streaming sizes), 1 = TMA-staged (54 %)
int g_cos_impl = 0;      // 0 = instance-major cosine maps when T % 8 == 0 (product); 1 = row-major cos_rows_kernel (comparison)

int launch_cos_att(int dt, const void* f, const void* kmat, int K, int T, int H, float* att, long long out_base, int n, cudaStream_t st) {
    if (n <= 0) return STAIR_OK;
    if (H % 8 || H > 256 * MAXC) return STAIR_ERR_UNSUPPORTED;
    if (g_row_stream >= 2 && row_stream_ok(dt, K, T, H)) return launch_cos_stream(dt, f, nullptr, kmat, nullptr, K, T, H, att, out_base, n, st);
    const long long rows = static_cast<long long>(n) * T;
    if (T % 8 == 0 && H <= 1024 && g_cos_impl == 0) {
        DISPATCH_DT(dt, AT, DISPATCH_CH_ONLY(H, (cos_inst_kernel<AT, CH, R><<<row_grid(cos_inst_kernel<AT, CH, R>, n, 1), 256, 0, st>>>(
                                reinterpret_cast<const AT*>(f), nullptr, reinterpret_cast<const AT*>(kmat), nullptr, K, T, H, att, out_base, n))));
        STAIR_CHECK_LAUNCH();
        return STAIR_OK;
    }
    DISPATCH_DT(dt, AT, DISPATCH_CH(H, (cos_rows_kernel<AT, CH, RPW><<<row_grid(cos_rows_kernel<AT, CH, RPW>, rows, RPW), 256, 0, st>>>(
                            reinterpret_cast<const AT*>(f), nullptr, reinterpret_cast<const AT*>(kmat), nullptr, K, T, H, att, out_base, static_cast<int>(rows)))));
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

int launch_existsframe(int dt, const void* vid, const int* feat_idx, const void* vec, const int* kw_idx, float* att, int out_base,
                       int n, int T, int H, cudaStream_t st) {
    if (n <= 0) return STAIR_OK;
    if (H % 8 || H > 256 * MAXC) return STAIR_ERR_UNSUPPORTED;
    if (g_row_stream >= 2 && row_stream_ok(dt, 1, T, H)) return launch_cos_stream(dt, vid, feat_idx, vec, kw_idx, 1, T, H, att, out_base, n, st);
    const long long rows = static_cast<long long>(n) * T;
    if (T % 8 == 0 && H <= 1024 && g_cos_impl == 0) {
        DISPATCH_DT(dt, AT, DISPATCH_CH_ONLY(H, (cos_inst_kernel<AT, CH, R><<<row_grid(cos_inst_kernel<AT, CH, R>, n, 1), 256, 0, st>>>(
                                reinterpret_cast<const AT*>(vid), feat_idx, reinterpret_cast<const AT*>(vec), kw_idx, 1, T, H, att, out_base, n))));
        STAIR_CHECK_LAUNCH();
        return STAIR_OK;
    }
    DISPATCH_DT(dt, AT, DISPATCH_CH(H, (cos_rows_kernel<AT, CH, RPW><<<row_grid(cos_rows_kernel<AT, CH, RPW>, rows, RPW), 256, 0, st>>>(
                            reinterpret_cast<const AT*>(vid), feat_idx, reinterpret_cast<const AT*>(vec), kw_idx, 1, T, H, att, out_base, static_cast<int>(rows)))));
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Temporal.relate[mode]                                                  video_nmn/modules.py:255-278, 316-323
// a = mean_K(att); while -> a; T_max <= 32: sigma(L3 relu(L2 relu(L1 a))) with Linear(T,T);
// else three Conv1d(1,1,k,'same',zeros): even k pads (k-1)/2 left, k-1-(k-1)/2 right.
// One block per instance, one thread per frame, two shared buffers.
// ------------------------------------------------------------------------------------------------------------------
struct RelateParams { const float* p[6]; };

__global__ void temporal_relate_kernel(const float* __restrict__ att, const int* __restrict__ att_idx, int K, int mode, int conv_k,
                                       RelateParams rp, float* __restrict__ att_out, int aux_base, int n, int T) {
    extern __shared__ float sm[];
    float* a = sm;
    float* b = sm + T;
    const int i = blockIdx.x;
    const int t = threadIdx.x;
    if (t < T) {
        const long long r0 = __ldg(att_idx + i);
        float s = 0.f;
        for (int k = 0; k < K; ++k) s += att[(r0 + k) * T + t];
        a[t] = s / static_cast<float>(K);
    }
    __syncthreads();
    if (mode != 0) {
        for (int layer = 0; layer < 3; ++layer) {
            const float* w = rp.p[2 * layer];
            const float* bias = rp.p[2 * layer + 1];
            float y = 0.f;
            if (t < T) {
                if (conv_k == 0) {                       // Linear(T, T): y[t] = sum_u W[t][u] a[u] + b[t]
                    y = __ldg(bias + t);
                    for (int u = 0; u < T; ++u) y += __ldg(w + t * T + u) * a[u];
                } else {                                 // Conv1d 'same' (cross-correlation): y[t] = b + sum_j w[j] a[t + j - left]
                    const int k = layer < 2 ? conv_k : 2 * conv_k + 1;
                    const int left = (k - 1) / 2;
                    y = __ldg(bias);
                    for (int j = 0; j < k; ++j) {
                        const int u = t + j - left;
                        if (u >= 0 && u < T) y += __ldg(w + j) * a[u];
                    }
                }
                y = layer < 2 ? fmaxf(y, 0.f) : sigmoidf_(y);
                b[t] = y;
            }
            __syncthreads();
            float* tmp = a; a = b; b = tmp;
        }
    }
    if (t < T) att_out[static_cast<long long>(aux_base + i) * T + t] = a[t];
}

int launch_temporal_relate(const float* att, const int* att_idx, int K, int mode, int conv_k, const float* const* params,
                           float* att_out, int aux_base, int n, int T, cudaStream_t st) {
    if (n <= 0) return STAIR_OK;
    if (T > 1024) return STAIR_ERR_UNSUPPORTED;
    RelateParams rp;
    for (int j = 0; j < 6; ++j) rp.p[j] = params ? params[j] : nullptr;
    const int threads = ((T + 31) / 32) * 32;
    temporal_relate_kernel<<<n, threads, 2 * T * sizeof(float), st>>>(att, att_idx, K, mode, conv_k, rp, att_out, aux_base, n, T);
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// LayerNorm(H), eps 1e-5, biased variance (Temporal.layer_norm, modules.py:283,327).  One warp per row.
// ------------------------------------------------------------------------------------------------------------------
template <typename AT, int CH, int RPW>
__global__ void __launch_bounds__(256, 2) layernorm_kernel(const AT* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                 AT* __restrict__ out, long long rows, int H) {
    const int lane = threadIdx.x & 31, hc = H / 8;
    const long long warp = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
    const float fh = static_cast<float>(H);
    for (long long row0 = warp * RPW; row0 < rows; row0 += nwarps * RPW) {
        Raw8<AT> v[RPW][CH];
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const long long row = row0 + r < rows ? row0 + r : rows - 1;
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                const int c = lane + 32 * i;
                if (c < hc) v[r][i].load_stream(x + row * H + c * 8); else v[r][i].zero();
            }
        }
        float mean[RPW], rstd[RPW];
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                float f[8]; v[r][i].unpack(f);
#pragma unroll
                for (int j = 0; j < 8; ++j) s += f[j];
            }
            mean[r] = s;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int r = 0; r < RPW; ++r) mean[r] += __shfl_xor_sync(0xffffffffu, mean[r], o);
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            mean[r] /= fh;
            float q = 0.f;
#pragma unroll
            for (int i = 0; i < CH; ++i)
                if (lane + 32 * i < hc) {
                    float f[8]; v[r][i].unpack(f);
#pragma unroll
                    for (int j = 0; j < 8; ++j) { const float d = f[j] - mean[r]; q += d * d; }
                }
            rstd[r] = q;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int r = 0; r < RPW; ++r) rstd[r] += __shfl_xor_sync(0xffffffffu, rstd[r], o);
#pragma unroll
        for (int r = 0; r < RPW; ++r) rstd[r] = rsqrtf(rstd[r] / fh + 1e-5f);
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            const int c = lane + 32 * i;
            if (c < hc) {
                Vec8<float> g, b; g.load(gamma + c * 8); b.load(beta + c * 8);
#pragma unroll
                for (int r = 0; r < RPW; ++r) {
                    if (row0 + r < rows) {
                        float f[8]; v[r][i].unpack(f);
                        Vec8<AT> o;
#pragma unroll
                        for (int j = 0; j < 8; ++j) o.v[j] = (f[j] - mean[r]) * rstd[r] * g.v[j] + b.v[j];
                        o.store(out + (row0 + r) * H + c * 8);
                    }
                }
            }
        }
    }
}

int launch_layernorm(int dt, const void* x, const float* gamma, const float* beta, void* out, long long rows, int H, cudaStream_t st) {
    if (rows <= 0) return STAIR_OK;
    if (H % 8 || H > 256 * MAXC) return STAIR_ERR_UNSUPPORTED;
    DISPATCH_DT(dt, AT, DISPATCH_CH(H, (layernorm_kernel<AT, CH, RPW><<<row_grid(layernorm_kernel<AT, CH, RPW>, rows, RPW), 256, 0, st>>>(
                            reinterpret_cast<const AT*>(x), gamma, beta, reinterpret_cast<AT*>(out), rows, H))));
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

// sum over frames: out[i] = sum_t x[i*T + t]                                  (Filter aggregation, modules.py:374-376)
template <typename AT>
__global__ void sum_T_kernel(const AT* __restrict__ x, AT* __restrict__ out, int n, int T, int H) {
    const int hc = H / 8;
    const long long total = static_cast<long long>(n) * hc;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / hc;
        const int c = static_cast<int>(i % hc) * 8;
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        int t = 0;
        for (; t + 8 <= T; t += 8) {                       // 8 independent 16-byte loads in flight, summed in frame order
            Vec8<AT> v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u].load_stream(x + (r * T + t + u) * H + c);
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] += v[u].v[j];
        }
        for (; t < T; ++t) {
            Vec8<AT> v; v.load(x + (r * T + t) * H + c);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += v.v[j];
        }
        Vec8<AT> o;
#pragma unroll
        for (int j = 0; j < 8; ++j) o.v[j] = acc[j];
        o.store(out + r * H + c);
    }
}

int launch_sum_T(int dt, const void* x, void* out, int n, int T, int H, cudaStream_t st) {
    if (n <= 0) return STAIR_OK;
    const int grid = min(blocks_for(static_cast<long long>(n) * (H / 8), 128), 148 * 16);
    DISPATCH_DT(dt, AT, (sum_T_kernel<AT><<<grid, 128, 0, st>>>(reinterpret_cast<const AT*>(x), reinterpret_cast<AT*>(out), n, T, H)));
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

// FilterFrame.attention: sigmoid(Linear(2H,1)([x_t | keyword]))                         modules.py:405-409
template <typename AT>
__global__ void ff_attn_kernel(const AT* __restrict__ x, const AT* __restrict__ vec, const int* __restrict__ kw_idx,
                               const float* __restrict__ w, const float* __restrict__ b, float* __restrict__ a, long long rows, int T, int H) {
    const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
    const int hc = H / 8;
    for (long long row = blockIdx.x * static_cast<long long>(warps) + (threadIdx.x >> 5); row < rows; row += static_cast<long long>(gridDim.x) * warps) {
        const int i = static_cast<int>(row / T);
        const AT* kr = vec + static_cast<long long>(__ldg(kw_idx + i)) * H;
        float s = 0.f;
        for (int c = lane; c < hc; c += 32) {
            Vec8<AT> xv, kv; xv.load(x + row * H + c * 8); kv.load(kr + c * 8);
#pragma unroll
            for (int j = 0; j < 8; ++j) s += xv.v[j] * __ldg(w + c * 8 + j) + kv.v[j] * __ldg(w + H + c * 8 + j);
        }
        s = warp_sum(s);
        if (lane == 0) a[row] = sigmoidf_(s + __ldg(b));
    }
}

int launch_ff_attn(int dt, const void* x, const void* vec, const int* kw_idx, const float* w, const float* b, float* a,
                   int n, int T, int H, cudaStream_t st) {
    if (n <= 0) return STAIR_OK;
    const long long rows = static_cast<long long>(n) * T;
    const int grid = min(blocks_for(rows, 8), 148 * 8);
    DISPATCH_DT(dt, AT, (ff_attn_kernel<AT><<<grid, 256, 0, st>>>(reinterpret_cast<const AT*>(x), reinterpret_cast<const AT*>(vec), kw_idx, w, b, a, rows, T, H)));
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

// AttnVideo: out[t,:] = attn[t] * feat[t,:]                                               modules.py:330-340
template <typename AT>
__global__ void attnvideo_kernel(AT* __restrict__ vid, const int* __restrict__ feat_idx, const float* __restrict__ att,
                                 const int* __restrict__ att_idx, int out_base, int n, int T, int H) {
    const int hc = H / 8;
    const long long per = static_cast<long long>(T) * hc;
    const long long total = per * n;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int r = static_cast<int>(i / per);
        const int rem = static_cast<int>(i % per);
        const int t = rem / hc, c = (rem % hc) * 8;
        const float s = att[static_cast<long long>(__ldg(att_idx + r)) * T + t];
        Vec8<AT> v; v.load(vid + (static_cast<long long>(__ldg(feat_idx + r)) * T + t) * H + c);
#pragma unroll
        for (int j = 0; j < 8; ++j) v.v[j] *= s;
        v.store(vid + (static_cast<long long>(out_base + r) * T + t) * H + c);
    }
}

int launch_attnvideo(int dt, void* vid, const int* feat_idx, const float* att, const int* att_idx, int out_base,
                     int n, int T, int H, cudaStream_t st) {
    if (n <= 0) return STAIR_OK;
    const int grid = min(blocks_for(static_cast<long long>(n) * T * (H / 8), 256), 148 * 16);
    DISPATCH_DT(dt, AT, (attnvideo_kernel<AT><<<grid, 256, 0, st>>>(reinterpret_cast<AT*>(vid), feat_idx, att, att_idx, out_base, n, T, H)));
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

// Relate: softmax_T(attn +/- beta[:T])  (nn.Softmax() on a 1-D tensor -> dim 0)           modules.py:417-435
__global__ void relate_kernel(const float* __restrict__ att, const int* __restrict__ att_idx, const float* __restrict__ beta, float sign,
                              float* __restrict__ att_out, int out_base, int n, int T) {
    const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
    for (int i = blockIdx.x * warps + (threadIdx.x >> 5); i < n; i += gridDim.x * warps) {
        const float* src = att + static_cast<long long>(att_idx ? __ldg(att_idx + i) : i) * T;
        float m = -INFINITY;
        for (int t = lane; t < T; t += 32) m = fmaxf(m, src[t] + sign * __ldg(beta + t));
        m = warp_max(m);
        float s = 0.f;
        for (int t = lane; t < T; t += 32) s += expf(src[t] + sign * __ldg(beta + t) - m);
        s = warp_sum(s);
        float* dst = att_out + static_cast<long long>(out_base + i) * T;
        for (int t = lane; t < T; t += 32) dst[t] = expf(src[t] + sign * __ldg(beta + t) - m) / s;
    }
}

int launch_relate(const float* att, const int* att_idx, const float* beta, int sign, float* att_out, int out_base, int n, int T, cudaStream_t st) {
    if (n <= 0) return STAIR_OK;
    relate_kernel<<<min(blocks_for(n, 8), 148 * 8), 256, 0, st>>>(att, att_idx, beta, sign > 0 ? 1.f : (sign < 0 ? -1.f : 0.f), att_out, out_base, n, T);
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

// HasItem tail: sigmoid(Linear(H,1)(x_t))                                                  modules.py:128-129
template <typename AT, int CH, int RPW>
__global__ void rowdot_sigmoid_kernel(const AT* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                      float* __restrict__ att, long long out_off, long long rows, int H) {
    const int lane = threadIdx.x & 31, hc = H / 8;
    const long long warp = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
    Vec8<float> wv[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        const int c = lane + 32 * i;
        if (c < hc) wv[i].load(w + c * 8);
        else {
#pragma unroll
            for (int j = 0; j < 8; ++j) wv[i].v[j] = 0.f;
        }
    }
    const float bias = __ldg(b);
    for (long long row0 = warp * RPW; row0 < rows; row0 += nwarps * RPW) {
        Raw8<AT> v[RPW][CH];
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const long long row = row0 + r < rows ? row0 + r : rows - 1;
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                const int c = lane + 32 * i;
                if (c < hc) v[r][i].load_stream(x + row * H + c * 8); else v[r][i].zero();
            }
        }
        float s[RPW];
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            float a = 0.f;
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                float f[8]; v[r][i].unpack(f);
#pragma unroll
                for (int j = 0; j < 8; ++j) a += f[j] * wv[i].v[j];
            }
            s[r] = a;
        }
        // RPW sums with RPW/2 + ... + 1 + (5 - log2 RPW) shuffles; lane l ends up with the sum of row l >> SH
        constexpr int SH = RPW == 8 ? 2 : (RPW == 4 ? 3 : 4);
        const float mine = warp_sum_multi<RPW>(s, lane);
        const int r = lane >> SH;
        if ((lane & ((1 << SH) - 1)) == 0 && row0 + r < rows) att[out_off + row0 + r] = sigmoidf_(mine + bias);       // RPW consecutive floats per warp
    }
}

int launch_rowdot_sigmoid(int dt, const void* x, const float* w, const float* b, float* att, int out_base, int n, int T, int H, cudaStream_t st) {
    if (n <= 0) return STAIR_OK;
    if (H % 8 || H > 256 * MAXC) return STAIR_ERR_UNSUPPORTED;
    if (g_row_stream && row_stream_ok(dt, 1, T, H)) return launch_rowdot_stream(dt, x, w, b, att, out_base, n, T, H, st);
    const long long rows = static_cast<long long>(n) * T;
    DISPATCH_DT(dt, AT, DISPATCH_CH16(H, (rowdot_sigmoid_kernel<AT, CH, RPW><<<row_grid(rowdot_sigmoid_kernel<AT, CH, RPW>, rows, RPW), 256, 0, st>>>(
                            reinterpret_cast<const AT*>(x), w, b, att, static_cast<long long>(out_base) * T, rows, H))));
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

__global__ void drop_rows_kernel(float* __restrict__ x, long long rows, DropSpec d) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < rows; i += static_cast<long long>(gridDim.x) * blockDim.x)
        x[i] = drop_keep(drop_row_hash(d.key_lo, d.key_hi, d.row0 + i), 0u, d.thresh) ? x[i] * d.scale : 0.0f;
}

int launch_drop_rows(float* x, long long rows, DropSpec d, cudaStream_t st) {
    if (rows <= 0 || !d.thresh) return STAIR_OK;
    long long blocks = (rows + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    drop_rows_kernel<<<static_cast<int>(blocks), 256, 0, st>>>(x, rows, d);
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

// Choose: cos(k1,q) > cos(k2,q) ? k1 : k2 (strict '>', tie -> k2); device-side select, no host sync  modules.py:40-56
template <typename AT>
__global__ void choose_kernel(AT* __restrict__ vec, const int* __restrict__ k1, const int* __restrict__ k2, const int* __restrict__ q,
                              int out_base, int n, int H) {
    const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
    const int hc = H / 8;
    for (int i = blockIdx.x * warps + (threadIdx.x >> 5); i < n; i += gridDim.x * warps) {
        const AT* a = vec + static_cast<long long>(__ldg(k1 + i)) * H;
        const AT* b = vec + static_cast<long long>(__ldg(k2 + i)) * H;
        const AT* c = vec + static_cast<long long>(__ldg(q + i)) * H;
        float aa = 0, bb = 0, cc = 0, ac = 0, bc = 0;
        for (int ch = lane; ch < hc; ch += 32) {
            Vec8<AT> x, y, z; x.load(a + ch * 8); y.load(b + ch * 8); z.load(c + ch * 8);
#pragma unroll
            for (int j = 0; j < 8; ++j) { aa += x.v[j] * x.v[j]; bb += y.v[j] * y.v[j]; cc += z.v[j] * z.v[j]; ac += x.v[j] * z.v[j]; bc += y.v[j] * z.v[j]; }
        }
        aa = warp_sum(aa); bb = warp_sum(bb); cc = warp_sum(cc); ac = warp_sum(ac); bc = warp_sum(bc);
        const float nq = fmaxf(sqrtf(cc), 1e-8f);
        const float c1 = ac / (fmaxf(sqrtf(aa), 1e-8f) * nq), c2 = bc / (fmaxf(sqrtf(bb), 1e-8f) * nq);
        const AT* src = c1 > c2 ? a : b;
        AT* dst = vec + static_cast<long long>(out_base + i) * H;
        for (int ch = lane; ch < hc; ch += 32) { Vec8<AT> v; v.load(src + ch * 8); v.store(dst + ch * 8); }
    }
}

int launch_choose(int dt, void* vec, const int* k1, const int* k2, const int* q, int out_base, int n, int H, cudaStream_t st) {
    if (n <= 0) return STAIR_OK;
    DISPATCH_DT(dt, AT, (choose_kernel<AT><<<min(blocks_for(n, 8), 148 * 8), 256, 0, st>>>(reinterpret_cast<AT*>(vec), k1, k2, q, out_base, n, H)));
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

// And (min) / XorFrame (|a-b|) on VEC rows or attention rows                             modules.py:7-12, 75-80
template <typename AT>
__global__ void binary_kernel(AT* __restrict__ base, const int* __restrict__ a_idx, const int* __restrict__ b_idx, int out_base, int unit,
                              int len, int op, int n) {
    const long long total = static_cast<long long>(n) * len;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int r = static_cast<int>(i / len), e = static_cast<int>(i % len);
        const float x = ld1<AT>(base + static_cast<long long>(__ldg(a_idx + r)) * unit + e);
        const float y = ld1<AT>(base + static_cast<long long>(__ldg(b_idx + r)) * unit + e);
        // torch.min propagates NaN; fminf does not
        const float z = op == STAIR_BIN_MIN ? ((x != x || y != y) ? NAN : fminf(x, y)) : fabsf(x - y);
        st1<AT>(base + static_cast<long long>(out_base) * unit + static_cast<long long>(r) * len + e, z);
    }
}

int launch_binary(int dt, void* base, const int* a_idx, const int* b_idx, int out_base, int unit, int len, int op, int n, cudaStream_t st) {
    if (n <= 0) return STAIR_OK;
    const int grid = min(blocks_for(static_cast<long long>(n) * len, 256), 148 * 16);
    DISPATCH_DT(dt, AT, (binary_kernel<AT><<<grid, 256, 0, st>>>(reinterpret_cast<AT*>(base), a_idx, b_idx, out_base, unit, len, op, n)));
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

// Array2: stack([f1, f2]) -> two adjacent VEC rows                                         modules.py:438-443
template <typename AT>
__global__ void array2_kernel(AT* __restrict__ vec, const int* __restrict__ a_idx, const int* __restrict__ b_idx, int out_base, int n, int H) {
    const int hc = H / 8;
    const long long total = static_cast<long long>(n) * 2 * hc;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int r = static_cast<int>(i / (2 * hc));
        const int rem = static_cast<int>(i % (2 * hc));
        const int which = rem / hc, c = (rem % hc) * 8;
        Vec8<AT> v; v.load(vec + static_cast<long long>(__ldg((which ? b_idx : a_idx) + r)) * H + c);
        v.store(vec + (static_cast<long long>(out_base) + 2LL * r + which) * H + c);
    }
}

int launch_array2(int dt, void* vec, const int* a_idx, const int* b_idx, int out_base, int n, int H, cudaStream_t st) {
    if (n <= 0) return STAIR_OK;
    const int grid = min(blocks_for(static_cast<long long>(n) * (H / 4), 256), 148 * 16);
    DISPATCH_DT(dt, AT, (array2_kernel<AT><<<grid, 256, 0, st>>>(reinterpret_cast<AT*>(vec), a_idx, b_idx, out_base, n, H)));
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

// Superlative tail                                                                         modules.py:243-247
template <typename AT>
__global__ void super_mix_kernel(const float* __restrict__ att, int K, int T, int H, int is_min, const AT* __restrict__ act_base,
                                 const int* __restrict__ act_idx, int act_unit, AT* __restrict__ dst, int n) {
    extern __shared__ float w[];           // [K]
    const int i = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
    for (int k = warp; k < K; k += warps) {
        float s = 0.f;
        for (int t = lane; t < T; t += 32) s += att[(static_cast<long long>(i) * K + k) * T + t];
        s = warp_sum(s);
        if (lane == 0) w[k] = s;
    }
    __syncthreads();
    if (warp == 0) {
        float m = -INFINITY;
        for (int k = lane; k < K; k += 32) m = fmaxf(m, w[k]);
        m = warp_max(m);
        float s = 0.f;
        for (int k = lane; k < K; k += 32) s += expf(w[k] - m);
        s = warp_sum(s);
        for (int k = lane; k < K; k += 32) { const float p = expf(w[k] - m) / s; w[k] = is_min ? 1.f - p : p; }
    }
    __syncthreads();
    const AT* rows = act_base + static_cast<long long>(__ldg(act_idx + i)) * act_unit * H;
    for (int c = threadIdx.x; c < H / 8; c += blockDim.x) {
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int k = 0; k < K; ++k) {
            Vec8<AT> v; v.load(rows + static_cast<long long>(k) * H + c * 8);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += w[k] * v.v[j];
        }
        Vec8<AT> o;
#pragma unroll
        for (int j = 0; j < 8; ++j) o.v[j] = acc[j];
        o.store(dst + static_cast<long long>(i) * H + c * 8);
    }
}

int launch_super_mix(int dt, const float* att, int K, int T, int H, int is_min, const void* act_base, const int* act_idx, int act_unit,
                     void* dst, int n, cudaStream_t st) {
    if (n <= 0) return STAIR_OK;
    DISPATCH_DT(dt, AT, (super_mix_kernel<AT><<<n, 128, K * sizeof(float), st>>>(att, K, T, H, is_min, reinterpret_cast<const AT*>(act_base), act_idx,
                                                                                act_unit, reinterpret_cast<AT*>(dst), n)));
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

// pretrain heads: Linear(H, nout<=2) (Equals :29, Xor :63, Exists :150) and L2Normalize (module_net.py:211-216)
template <typename AT>
__global__ void small_head_kernel(const AT* __restrict__ vec, int row_base, const float* __restrict__ w, const float* __restrict__ b,
                                  int nout, float* __restrict__ out, int out_base, int n, int H) {
    const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
    const int hc = H / 8;
    for (int idx = blockIdx.x * warps + (threadIdx.x >> 5); idx < n * nout; idx += gridDim.x * warps) {
        const int i = idx / nout, o = idx % nout;
        const AT* x = vec + static_cast<long long>(row_base + i) * H;
        float s = 0.f;
        for (int c = lane; c < hc; c += 32) {
            Vec8<AT> v; v.load(x + c * 8);
#pragma unroll
            for (int j = 0; j < 8; ++j) s += v.v[j] * __ldg(w + o * H + c * 8 + j);
        }
        s = warp_sum(s);
        if (lane == 0) out[static_cast<long long>(out_base + i) * 2 + o] = s + __ldg(b + o);
    }
}

int launch_small_head(int dt, const void* vec, int row_base, const float* w, const float* b, int nout, float* out, int out_base, int n, int H, cudaStream_t st) {
    if (n <= 0) return STAIR_OK;
    if (nout > 2) return STAIR_ERR_ARG;
    DISPATCH_DT(dt, AT, (small_head_kernel<AT><<<min(blocks_for(static_cast<long long>(n) * nout, 8), 148 * 8), 256, 0, st>>>(
                            reinterpret_cast<const AT*>(vec), row_base, w, b, nout, out, out_base, n, H)));
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

template <typename AT>
__global__ void l2norm_kernel(const AT* __restrict__ vec, int row_base, float* __restrict__ out, int out_base, int n, int H) {
    const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
    const int hc = H / 8;
    for (int i = blockIdx.x * warps + (threadIdx.x >> 5); i < n; i += gridDim.x * warps) {
        const AT* x = vec + static_cast<long long>(row_base + i) * H;
        float ss = 0.f;
        for (int c = lane; c < hc; c += 32) {
            Vec8<AT> v; v.load(x + c * 8);
#pragma unroll
            for (int j = 0; j < 8; ++j) ss += v.v[j] * v.v[j];
        }
        const float inv = 1.0f / fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
        float* o = out + static_cast<long long>(out_base + i) * H;
        for (int c = lane; c < hc; c += 32) {
            Vec8<AT> v; v.load(x + c * 8);
            Vec8<float> r;
#pragma unroll
            for (int j = 0; j < 8; ++j) r.v[j] = v.v[j] * inv;
            r.store(o + c * 8);
        }
    }
}

// same, R rows per warp and pass held in registers (each row read once, past L1); H <= 256 * CH
template <typename AT, int CH, int R>
__global__ void __launch_bounds__(256) l2norm_rows_kernel(const AT* __restrict__ vec, int row_base, float* __restrict__ out, int out_base, int n, int H) {
    const int lane = threadIdx.x & 31, hc = H / 8;
    const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nwarps = gridDim.x * (blockDim.x >> 5);
    for (int i0 = warp * R; i0 < n; i0 += nwarps * R) {
        Raw8<AT> x[R][CH];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int i = i0 + r < n ? i0 + r : n - 1;
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                const int c = lane + 32 * k;
                if (c < hc) x[r][k].load_stream(vec + static_cast<long long>(row_base + i) * H + c * 8); else x[r][k].zero();
            }
        }
        float ss[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                float f[8]; x[r][k].unpack(f);
#pragma unroll
                for (int j = 0; j < 8; ++j) a = fmaf(f[j], f[j], a);
            }
            ss[r] = a;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int r = 0; r < R; ++r) ss[r] += __shfl_xor_sync(0xffffffffu, ss[r], o);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (i0 + r >= n) break;
            const float inv = 1.0f / fmaxf(sqrtf(ss[r]), 1e-12f);
            float* o = out + static_cast<long long>(out_base + i0 + r) * H;
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                const int c = lane + 32 * k;
                if (c < hc) {
                    float f[8]; x[r][k].unpack(f);
                    Vec8<float> y;
#pragma unroll
                    for (int j = 0; j < 8; ++j) y.v[j] = f[j] * inv;
                    y.store(o + c * 8);
                }
            }
        }
    }
}

int launch_l2norm(int dt, const void* vec, int row_base, float* out, int out_base, int n, int H, cudaStream_t st) {
    if (n <= 0) return STAIR_OK;
    if (H % 8 == 0 && H <= 1024) {
        DISPATCH_DT(dt, AT, DISPATCH_CH_ONLY(H, (l2norm_rows_kernel<AT, CH, 4><<<row_grid(l2norm_rows_kernel<AT, CH, 4>, n, 4), 256, 0, st>>>(
                                reinterpret_cast<const AT*>(vec), row_base, out, out_base, n, H))));
        STAIR_CHECK_LAUNCH();
        return STAIR_OK;
    }
    DISPATCH_DT(dt, AT, (l2norm_kernel<AT><<<min(blocks_for(n, 8), 148 * 8), 256, 0, st>>>(reinterpret_cast<const AT*>(vec), row_base, out, out_base, n, H)));
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

// torch.argmax(logits): index of the first maximal element (NaN counts as maximal, like torch)   train_module.py:252
__global__ void argmax_kernel(const float* __restrict__ x, int* __restrict__ out, int rows, int cols) {
    const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
    for (int r = blockIdx.x * warps + (threadIdx.x >> 5); r < rows; r += gridDim.x * warps) {
        float best = -INFINITY; int bi = 0x7fffffff; bool bnan = false;
        for (int c = lane; c < cols; c += 32) {
            const float v = x[static_cast<long long>(r) * cols + c];
            const bool vn = v != v;
            if (bi == 0x7fffffff || (vn && !bnan) || (!bnan && v > best)) { best = v; bi = c; bnan = vn; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            const bool on = ov != ov;
            bool take;
            if (oi == 0x7fffffff) take = false;
            else if (bi == 0x7fffffff) take = true;
            else if (on != bnan) take = on;
            else if (on) take = oi < bi;
            else take = ov > best || (ov == best && oi < bi);
            if (take) { best = ov; bi = oi; bnan = on; }
        }
        if (lane == 0) out[r] = bi;
    }
}

int launch_argmax(const float* logits, int* out, int rows, int cols, cudaStream_t st) {
    if (rows <= 0) return STAIR_OK;
    argmax_kernel<<<min(blocks_for(rows, 8), 148 * 8), 256, 0, st>>>(logits, out, rows, cols);
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// TemporalModule.relate_: cumsum-based before/after/between masks                          modules.py:290-308
// (dead code in the reference forward; exported and parity-tested because the north star names the scans.)
// One warp per instance; inclusive warp prefix scans over T in chunks of 32 with a running carry.
// ------------------------------------------------------------------------------------------------------------------
// The partial sums are carried in double and rounded to fp32 once per output: torch's CPU cumsum (the reference's relate_,
// modules.py:302-304) accumulates in double too, so a run of zeros after the ReLU leaves the scan EXACTLY flat — an fp32 Kogge-Stone scan
// associates every prefix differently and breaks such plateaus by an ulp, which moves the argmax of the mask.
__device__ __forceinline__ double warp_incl_scan(double v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
    }
    return v;
}

// before[t] = sum_{u<=t} relu(a[u]) ; after[t] = sum_{u>=t} relu(a[u]) ; written to shared memory
__device__ void scan_before_after(const float* a, int T, int lane, float* before, float* after) {
    double carry = 0.0;
    for (int t0 = 0; t0 < T; t0 += 32) {
        const int t = t0 + lane;
        const double v = t < T ? static_cast<double>(fmaxf(a[t], 0.f)) : 0.0;
        const double s = warp_incl_scan(v, lane) + carry;
        if (t < T) before[t] = static_cast<float>(s);
        carry = __shfl_sync(0xffffffffu, s, 31);
    }
    carry = 0.0;
    for (int t0 = 0; t0 < T; t0 += 32) {
        const int t = T - 1 - (t0 + lane);
        const double v = t >= 0 ? static_cast<double>(fmaxf(a[t], 0.f)) : 0.0;
        const double s = warp_incl_scan(v, lane) + carry;
        if (t >= 0) after[t] = static_cast<float>(s);
        carry = __shfl_sync(0xffffffffu, s, 31);
    }
}

__global__ void relate_scan_kernel(const float* __restrict__ att, int mode, float* __restrict__ out, int n, int T) {
    extern __shared__ float sm[];
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* before = sm + static_cast<size_t>(warp) * 2 * T;
    float* after = before + T;
    const int K = mode == 3 ? 2 : 1;
    for (int i = blockIdx.x * warps + warp; i < n; i += gridDim.x * warps) {
        const float* a = att + static_cast<long long>(i) * K * T;
        float* o = out + static_cast<long long>(i) * T;
        if (mode == 0) {
            for (int t = lane; t < T; t += 32) o[t] = a[t];
            continue;
        }
        scan_before_after(a, T, lane, before, after);
        __syncwarp();
        if (mode == 1) { for (int t = lane; t < T; t += 32) o[t] = before[t]; }
        else if (mode == 2) { for (int t = lane; t < T; t += 32) o[t] = after[t]; }
        else {
            float ma[32];      // min(before_a, after_a) for this lane's frames (T <= 1024)
            int cnt = 0;
            for (int t = lane; t < T; t += 32) ma[cnt++] = fminf(before[t], after[t]);
            __syncwarp();
            scan_before_after(a + T, T, lane, before, after);
            __syncwarp();
            cnt = 0;
            for (int t = lane; t < T; t += 32) o[t] = fmaxf(ma[cnt++], fminf(before[t], after[t]));
        }
        __syncwarp();
    }
}

int launch_relate_scan(const float* att, int mode, float* out, int n, int T, cudaStream_t st) {
    if (n <= 0) return STAIR_OK;
    if (T > 1024 || mode < 0 || mode > 3) return STAIR_ERR_ARG;
    const int warps = 4;
    relate_scan_kernel<<<min(blocks_for(n, warps), 148 * 8), warps * 32, warps * 2 * T * sizeof(float), st>>>(att, mode, out, n, T);
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

}  // namespace stair

// ---- exported single-operator entry points -----------------------------------------------------------------------
using namespace stair;

extern "C" int stair_relate_scan(const float* att, int mode, float* out, int n, int T, void* stream) {
    return launch_relate_scan(att, mode, out, n, T, reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int stair_cos_attention(int dtype, const void* f, const void* k, int K, int T, int H, float* att, int n, void* stream) {
    return launch_cos_att(dtype, f, k, K, T, H, att, 0, n, reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int stair_relate(const float* att, const float* beta, int sign, float* out, int n, int T, void* stream) {
    return launch_relate(att, nullptr, beta, sign, out, 0, n, T, reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int stair_argmax(const float* x, int32_t* out, int rows, int cols, void* stream) {
    return launch_argmax(x, out, rows, cols, reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int stair_l2normalize(int dtype, const void* x, float* out, int n, int H, void* stream) {
    return launch_l2norm(dtype, x, 0, out, 0, n, H, reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int stair_layernorm(int dtype, const void* x, const float* gamma, const float* beta, void* out, long long rows, int H, void* stream) {
    return launch_layernorm(dtype, x, gamma, beta, out, rows, H, reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int stair_cast_bf16(const float* src, long long ld_src, void* dst, long long ld_dst, long long rows, int cols, void* stream) {
    return launch_stage_rows(STAIR_F32, src, ld_src, nullptr, 1, 1, reinterpret_cast<bf16*>(dst), ld_dst, rows, 1, rows, cols,
                             reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int stair_split3(const float* src, long long ld_src, void* dst, long long ld_dst, long long plane_rows, long long rows, int cols, void* stream) {
    return launch_stage_rows(STAIR_F32, src, ld_src, nullptr, 1, 1, reinterpret_cast<bf16*>(dst), ld_dst, plane_rows, 3, rows, cols,
                             reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int stair_sum_frames(int dtype, const void* x, void* out, int n, int T, int H, void* stream) {
    return launch_sum_T(dtype, x, out, n, T, H, reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int stair_attn_video(int dtype, void* vid, const int32_t* feat_idx, const float* att, const int32_t* att_idx, int out_base, int n, int T,
                                int H, void* stream) {
    return launch_attnvideo(dtype, vid, feat_idx, att, att_idx, out_base, n, T, H, reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int stair_exists_frame(int dtype, const void* vid, const int32_t* feat_idx, const void* vec, const int32_t* kw_idx, float* att, int out_base,
                                  int n, int T, int H, void* stream) {
    return launch_existsframe(dtype, vid, feat_idx, vec, kw_idx, att, out_base, n, T, H, reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int stair_hasitem_tail(int dtype, const void* x, const float* w, const float* b, float* att, int out_base, int n, int T, int H, void* stream) {
    return launch_rowdot_sigmoid(dtype, x, w, b, att, out_base, n, T, H, reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int stair_set_cos_impl(int impl) { g_cos_impl = impl ? 1 : 0; return STAIR_OK; }
extern "C" int stair_set_row_stream(int on) { g_row_stream = on < 0 ? 0 : (on > 2 ? 2 : on); return STAIR_OK; }
