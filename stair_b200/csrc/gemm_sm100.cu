// Dense contraction for the NMN hot path:  C[M,N] = act(row_scale[m] * (A[M,K] · W[N,K]^T) + bias[n])
//
// This is the only GEMM-shaped work on the path (nn.Linear in every module, the LSTM input/recurrent
// projections, the decoder: video_nmn/modules.py passim, video_nmn/module_net.py:39-53).  Both operands are
// K-major bf16 (activations row-major [M,K]; nn.Linear weights are already [N,K]), accumulation is fp32.
//
// sm_100a design: persistent warp-specialised kernel, one CTA per SM.
//   warp 0 lane 0 : TMA producer  (cp.async.bulk.tensor.2d, 128B swizzle, mbarrier complete_tx)
//   warp 1 lane 0 : tcgen05.mma issuer (UMMA 128 x BN x 16, kind::f16, fp32 accumulators in TMEM)
//   warp 2        : TMEM allocator / deallocator
//   warps 4..7    : epilogue (tcgen05.ld 32x32b -> registers -> row_scale/bias/ReLU -> global)
// smem ring of STAGES {A 128x64, W BNx64} tiles; TMEM double-buffered (2 x BN columns) so the epilogue of
// tile i overlaps the MMAs of tile i+1.
//
// "Split" mode (nplanes = 3) evaluates an fp32-grade product from bf16 planes x = x0 + x1 + x2 (each bf16):
// the six significant plane products are accumulated into the same TMEM tile by looping the K range six
// times with different plane row-offsets.  It is the strict-parity mode.
#include "tc_ptx.cuh"
#include <cstdlib>

namespace stair {

constexpr int BM = 128;
constexpr int BK = 64;                         // 64 bf16 = 128 B = one SWIZZLE_128B atom row
constexpr int A_STAGE_BYTES = BM * BK * 2;     // 16 KiB
constexpr int GEMM_THREADS = 256;
constexpr bool g_tail_wait_full = true;         // true = wait for the TMA stores to complete before exit; false = only for their smem reads
                                                // (what CUTLASS epilogues do; tests pass, but it measured no gain: 1.507 ms per forward either way)

struct GemmParams {
    int M, N, K;
    int num_kb;            // ceil(K / 64)
    int ksplit;            // split-K factor (>1 only with accumulate: partial tiles are added to C with vector atomics)
    int kb_per_split;      // k-blocks per split
    int nseg;              // 1, or 6 in split mode
    int a_plane_rows, w_plane_rows;
    const float* bias;     // [N] or null
    const float* row_scale;// [M] or null
    void* C;
    long long ldc;
    int out_dtype, act, accumulate, vec_ok;
    int tma_store;         // epilogue stages tiles in shared memory and writes them with TMA (C aligned, no accumulate)
    const int* a_slots;    // gather mode (else null)
    int slot_rows;         // rows per slot (divides BM)
    int num_slots;         // ceil(M / slot_rows)
    int* err_flag;
    unsigned long long* dbg;   // debug timeline of block 0 (globaltimer ns), null in production
    DropSpec drop;             // dropout after the activation (thresh 0 = off)
    int atomic_acc;            // fp32 accumulate with atomics even when ksplit == 1
    int mn_major;              // operands stored [k][m] / [k][n] (C = A^T . W): MN-major UMMA tiles, 3-D tensor maps {mn, k, plane}
    bf16* sum_out;             // frame-sum epilogue (null = off): [M / sum_T, ld_sum] bf16, written INSTEAD of C
    long long ld_sum;
    int sum_T;
};

// plane pairs (a,b) of the 6-term bf16x3 product, smallest contributions first
__device__ __constant__ int c_seg_a[6] = {2, 0, 1, 1, 0, 0};
__device__ __constant__ int c_seg_b[6] = {0, 2, 1, 0, 1, 0};

__device__ __forceinline__ void dbg_stamp(const GemmParams& p, int slot) {
    if (p.dbg && blockIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        p.dbg[slot] = t;
    }
}

// act(rs * acc + bias) for one 32-column TMEM chunk of this thread's row
__device__ __forceinline__ void epilogue_drop(const GemmParams& p, int row, int n, float (&v)[32]) {
    const uint32_t rh = drop_row_hash(p.drop.key_lo, p.drop.key_hi, p.drop.row0 + row);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = drop_keep(rh, static_cast<uint32_t>(n + j), p.drop.thresh) ? v[j] * p.drop.scale : 0.0f;
}

__device__ __forceinline__ void epilogue_math(const GemmParams& p, int row, int n, float rs, const uint32_t (&raw)[32], float (&v)[32]) {
    if (p.bias && n + 32 <= p.N) {
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + n);     // n is a multiple of 32: 16-byte aligned
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 b = __ldg(b4 + j);
            v[4 * j] = fmaf(__uint_as_float(raw[4 * j]), rs, b.x);
            v[4 * j + 1] = fmaf(__uint_as_float(raw[4 * j + 1]), rs, b.y);
            v[4 * j + 2] = fmaf(__uint_as_float(raw[4 * j + 2]), rs, b.z);
            v[4 * j + 3] = fmaf(__uint_as_float(raw[4 * j + 3]), rs, b.w);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            float x = __uint_as_float(raw[j]) * rs;
            if (p.bias && n + j < p.N) x += __ldg(p.bias + n + j);
            v[j] = x;
        }
    }
    if (p.act == STAIR_ACT_RELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
    }
    if (p.drop.thresh) epilogue_drop(p, row, n, v);
}

// write this thread's 32 values into the warp's staging buffer (row = lane, 128-byte rows, SWIZZLE_128B chunk order)
__device__ __forceinline__ void stage_chunk(uint32_t buf, int lane, int half, int out_dtype, const float (&v)[32]) {
    const uint32_t rowaddr = buf + static_cast<uint32_t>(lane) * 128u;
    const uint32_t sw = static_cast<uint32_t>(lane & 7);
    if (out_dtype == STAIR_BF16) {                  // 32 columns = 64 bytes = 16-byte chunks 4*half .. 4*half+3
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint32_t ch = static_cast<uint32_t>(4 * half + c);
            st_shared_v4(rowaddr + ((ch ^ sw) << 4), pack_bf16(v[8 * c], v[8 * c + 1]), pack_bf16(v[8 * c + 2], v[8 * c + 3]),
                         pack_bf16(v[8 * c + 4], v[8 * c + 5]), pack_bf16(v[8 * c + 6], v[8 * c + 7]));
        }
    } else {                                        // 32 columns = 128 bytes = the whole row
#pragma unroll
        for (int c = 0; c < 8; ++c)
            st_shared_v4(rowaddr + ((static_cast<uint32_t>(c) ^ sw) << 4), __float_as_uint(v[4 * c]), __float_as_uint(v[4 * c + 1]),
                         __float_as_uint(v[4 * c + 2]), __float_as_uint(v[4 * c + 3]));
    }
}

// EPI = epilogue warp sets (4 warps each, one per TMEM lane quarter).  With EPI = 2 the second set (warps 8-11) takes the upper half of
// the tile's columns: short-K GEMMs (module Linears K = 512, text projection K = 300) are bound by the epilogue — TMEM -> registers ->
// bias / ReLU / dropout -> bf16 -> swizzled smem -> TMA store takes longer than the tile's MMAs — so two sets halve the tile time; the
// extra staging costs one pipeline stage of shared memory, which a K loop of <= 8 k-blocks does not miss.
template <int BN, int STAGES, int EPI = 1>
struct GemmSmem {
    static constexpr int B_STAGE_BYTES = BN * BK * 2;
    static constexpr int TILE_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES);
    static constexpr int STAGING_BYTES = EPI * 4 * 2 * 4096;      // 4 EPI epilogue warps x 2 buffers x (32 rows x 128 B)
    static constexpr int BAR_BYTES = (2 * STAGES + 4) * 8 + 16;
    static constexpr int TOTAL = TILE_BYTES + STAGING_BYTES + BAR_BYTES + 1024;   // +1024: manual alignment slack
};

// ------------------------------------------------------------------------------------------------
// epilogue store of one 32-column chunk of one row
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void epilogue_store(const GemmParams& p, int row, int n, const uint32_t (&raw)[32]) {
    const float rs = p.row_scale ? __ldg(p.row_scale + row) : 1.0f;
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        float x = __uint_as_float(raw[j]) * rs;
        if (p.bias && n + j < p.N) x += __ldg(p.bias + n + j);
        if (p.act == STAIR_ACT_RELU) x = fmaxf(x, 0.0f);
        v[j] = x;
    }
    if (p.drop.thresh) epilogue_drop(p, row, n, v);
    const bool full = p.vec_ok && (n + 32 <= p.N);
    if (p.out_dtype == STAIR_BF16) {
        bf16* out = reinterpret_cast<bf16*>(p.C) + static_cast<long long>(row) * p.ldc + n;
        if (full) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                uint4 r;
                __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
                for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[j + 2 * i], v[j + 2 * i + 1]);
                *reinterpret_cast<uint4*>(out + j) = r;
            }
        } else {
            for (int j = 0; j < 32; ++j) if (n + j < p.N) out[j] = __float2bfloat16_rn(v[j]);
        }
    } else {
        float* out = reinterpret_cast<float*>(p.C) + static_cast<long long>(row) * p.ldc + n;
        if (full) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                float4 r = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                if (p.ksplit > 1 || p.atomic_acc) { atomicAdd(reinterpret_cast<float4*>(out + j), r); continue; }     // red.global.add.v4.f32
                if (p.accumulate) {
                    float4 o = *reinterpret_cast<const float4*>(out + j);
                    r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w;
                }
                *reinterpret_cast<float4*>(out + j) = r;
            }
        } else {
            for (int j = 0; j < 32; ++j)
                if (n + j < p.N) {
                    if (p.ksplit > 1 || p.atomic_acc) atomicAdd(out + j, v[j]);
                    else out[j] = p.accumulate ? out[j] + v[j] : v[j];
                }
        }
    }
}


// TMA-store epilogue of one accumulator tile for one epilogue warp: TMEM -> registers (next chunk's tcgen05.ld in flight while this
// one is processed) -> row_scale / bias / ReLU / dropout -> swizzled smem -> TMA store.  t_base = TMEM address of the warp's lane
// quarter and the accumulator stage, row = this thread's global row, row0 = the warp's first global row, [c_begin, c_end) = the
// 32-column chunks this warp owns, stage0 = shared address of the warp's two 4 KiB staging buffers, sbuf = which one is next.
__device__ __forceinline__ void epilogue_tile_tma(const GemmParams& p, const CUtensorMap& tmC, uint32_t t_base, int row, int row0, int n0,
                                                  int c_begin, int c_end, uint32_t stage0, int& sbuf, int lane) {
    const float rs = (p.row_scale && row < p.M) ? __ldg(p.row_scale + row) : 1.0f;
    const int per_buf = p.out_dtype == STAIR_BF16 ? 2 : 1;          // TMEM chunks per 128-byte staging row
    uint32_t ra[32], rb[32];
    float v[32];
    tmem_ld32(t_base + static_cast<uint32_t>(c_begin * 32), ra);
#pragma unroll 1
    for (int c = c_begin; c < c_end; c += 2) {
        if (n0 + c * 32 >= p.N) break;                               // warp-uniform
        tmem_ld_wait();
        tmem_ld32(t_base + static_cast<uint32_t>((c + 1) * 32), rb);
#pragma unroll
        for (int hsel = 0; hsel < 2; ++hsel) {
            const int cc = c + hsel;
            if (hsel == 1) {
                tmem_ld_wait();
                if (cc + 1 < c_end) tmem_ld32(t_base + static_cast<uint32_t>((cc + 1) * 32), ra);
            }
            const int n = n0 + cc * 32;
            const int half = per_buf == 2 ? hsel : 0;
            if (half == 0) {                                         // about to refill a staging buffer: it must be drained
                if (lane == 0) bulk_wait_read1();
                __syncwarp();
            }
            const uint32_t buf = stage0 + static_cast<uint32_t>(sbuf) * 4096u;
            if (n < p.N) {
                epilogue_math(p, row, n, rs, hsel == 0 ? ra : rb, v);
                stage_chunk(buf, lane, half, p.out_dtype, v);
            }
            if (half == per_buf - 1) {
                fence_async_smem();
                __syncwarp();
                const int nbox = n - (per_buf == 2 ? 32 : 0);
                if (lane == 0 && nbox < p.N) { tma_store_2d(&tmC, buf, nbox, row0); bulk_commit(); }
                sbuf ^= 1;
            }
        }
    }
    tmem_ld_wait();
}

// Frame-sum epilogue of one accumulator tile for one epilogue warp (GemmArgs::sum_out): the activated values are rounded to bf16 exactly as the
// stored tile would have been, transposed through the warp's staging buffer (row = lane -> column = lane) and summed over the T consecutive
// rows of an instance in frame order (the order of sum_T_kernel).  T <= 32: an instance lies inside the warp's 32 rows.  T = 64 / 128: the
// instance spans 2 / 4 lane quarters; their 32-row partial sums meet in the epilogue set's shared scratch behind a named barrier (every warp
// of the set runs this function for every tile, also for rows past M, which contribute zeros).
__device__ __forceinline__ void epilogue_tile_sum(const GemmParams& p, uint32_t t_base, int row, int row0, int n0, int c_begin, int c_end,
                                                  float* scratch, float* part, int quarter, int bar_id, int lane) {
    const float rs = (p.row_scale && row < p.M) ? __ldg(p.row_scale + row) : 1.0f;
    const int T = p.sum_T;
    const int len = T < 32 ? T : 32, groups = 32 / len, span = T > 32 ? T / 32 : 1;      // span = lane quarters per instance
    int parity = 0;
#pragma unroll 1
    for (int c = c_begin; c < c_end; ++c) {
        const int n = n0 + c * 32;
        if (n >= p.N) break;                                          // warp-uniform, identical in the four warps of a set
        uint32_t raw[32];
        float v[32];
        tmem_ld32(t_base + static_cast<uint32_t>(c * 32), raw);
        tmem_ld_wait();
        epilogue_math(p, row, n, rs, raw, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) scratch[lane * 33 + j] = row < p.M ? __bfloat162float(__float2bfloat16_rn(v[j])) : 0.0f;
        __syncwarp();
        for (int g = 0; g < groups; ++g) {
            float s = 0.0f;
            for (int t = 0; t < len; ++t) s += scratch[(g * len + t) * 33 + lane];
            if (span == 1) {
                const long long inst = (row0 + g * len) / T;
                if (row0 + g * len < p.M && n + lane < p.N) p.sum_out[inst * p.ld_sum + n + lane] = __float2bfloat16_rn(s);
            } else {
                part[(parity * 4 + quarter) * 32 + lane] = s;
            }
        }
        if (span > 1) {
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
            if (quarter % span == 0) {
                float s = 0.0f;
                for (int q = 0; q < span; ++q) s += part[(parity * 4 + quarter + q) * 32 + lane];
                const long long inst = row0 / T;
                if (row0 < p.M && n + lane < p.N) p.sum_out[inst * p.ld_sum + n + lane] = __float2bfloat16_rn(s);
            }
            parity ^= 1;
        }
        __syncwarp();
    }
}

// OCC = CTAs per SM the kernel is built for.  OCC = 2 (BN = 128, two pipeline stages, 99 KB of shared memory, 256 TMEM columns) was written
// for the module phase, whose GEMMs are a few microseconds of tensor work behind ~8 us of fixed latency (launch, barrier / TMEM set-up,
// first operand fetch, epilogue drain): two resident CTAs per SM let the GEMMs of two lanes run side by side.  Measured: no gain (the phase
// is bound by the length of its dependency chains, not by SM slots), so it is a comparison form (stair_set_gemm_small).
template <int BN, int STAGES, int EPI, int OCC>
__global__ void __launch_bounds__(GEMM_THREADS + 128 * (EPI - 1), OCC)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
    using S = GemmSmem<BN, STAGES, EPI>;
    static_assert(EPI == 1 || (EPI == 2 && BN >= 128), "two epilogue sets split the tile's 64-column boxes");
    constexpr int B_STAGE_BYTES = S::B_STAGE_BYTES;
    constexpr uint32_t TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;      // power of two by construction (BN in {64,128,256})
    constexpr uint32_t IDESC = make_idesc_bf16(BM, BN);

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
    uint8_t* sStage = smem + S::TILE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::TILE_BYTES + S::STAGING_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) dbg_stamp(p, 0);
    const int tiles_n = (p.N + BN - 1) / BN;
    const int tiles_m = (p.M + BM - 1) / BM;
    const int mn_tiles = tiles_m * tiles_n;
    const int total_tiles = mn_tiles * p.ksplit;           // work item = (output tile, K split); the K splits of a tile are adjacent in time

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
        if (p.tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmC)) : "memory");
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 128 * EPI); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    if (threadIdx.x == 0) dbg_stamp(p, 1);

    if (warp == 0) {
        // ===================== TMA producer (whole warp: in gather mode lane j loads slot j of the tile) ==========
        int stage = 0; uint32_t phase = 0;
        for (int item = blockIdx.x; item < total_tiles; item += gridDim.x) {
            const int tile = item % mn_tiles, ks = item / mn_tiles;
            const int kb0 = ks * p.kb_per_split, kb1 = min(p.num_kb, kb0 + p.kb_per_split);
            const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
            int my_slot = -1, nvalid = 0;
            if (p.a_slots) {
                const int s0 = m0 / p.slot_rows;
                nvalid = min(BM / p.slot_rows, p.num_slots - s0);
                if (lane < nvalid) my_slot = __ldg(p.a_slots + s0 + lane);
            }
            const uint32_t a_bytes = p.a_slots ? static_cast<uint32_t>(nvalid * p.slot_rows * BK * 2) : A_STAGE_BYTES;
            for (int seg = 0; seg < p.nseg; ++seg) {
                const int a_row = m0 + (p.nseg > 1 ? c_seg_a[seg] * p.a_plane_rows : 0);
                const int b_row = n0 + (p.nseg > 1 ? c_seg_b[seg] * p.w_plane_rows : 0);
                for (int kb = kb0; kb < kb1; ++kb) {
                    if (lane == 0) {
                        mbar_wait(&empty_bar[stage], phase ^ 1, p.err_flag, 101);
                        mbar_arrive_expect_tx(&full_bar[stage], a_bytes + B_STAGE_BYTES);
                        if (p.mn_major) {
                            // [64 k-rows][64 mn] boxes (128-byte rows): one box per 64 output rows / columns of the tile; rows
                            // past K and columns past M / N are zero-filled by TMA (per plane: the plane is the third coordinate)
                            const int pa = p.nseg > 1 ? c_seg_a[seg] : 0, pb = p.nseg > 1 ? c_seg_b[seg] : 0;
#pragma unroll
                            for (int j = 0; j < BN / 64; ++j)
                                tma_load_3d(sB + stage * B_STAGE_BYTES + j * 8192, &tmB, &full_bar[stage], n0 + 64 * j, kb * BK, pb);
#pragma unroll
                            for (int j = 0; j < BM / 64; ++j)
                                tma_load_3d(sA + stage * A_STAGE_BYTES + j * 8192, &tmA, &full_bar[stage], m0 + 64 * j, kb * BK, pa);
                        } else {
                        tma_load_2d(sB + stage * B_STAGE_BYTES, &tmB, &full_bar[stage], kb * BK, b_row);
                        if (!p.a_slots) tma_load_2d(sA + stage * A_STAGE_BYTES, &tmA, &full_bar[stage], kb * BK, a_row);
                        }
                    }
                    if (p.a_slots) {
                        __syncwarp();
                        if (my_slot >= 0)
                            tma_load_3d(sA + stage * A_STAGE_BYTES + lane * p.slot_rows * (BK * 2), &tmA, &full_bar[stage],
                                        kb * BK, 0, my_slot);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===================== MMA issuer =====================
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int item = blockIdx.x; item < total_tiles; item += gridDim.x) {
                const int ks = item / mn_tiles;
                const int kiters = p.nseg * (min(p.num_kb, (ks + 1) * p.kb_per_split) - ks * p.kb_per_split);
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1, p.err_flag, 102);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
                for (int it = 0; it < kiters; ++it) {
                    mbar_wait(&full_bar[stage], phase, p.err_flag, 103);
                    tcgen05_fence_after();
                    if (it == 0 && item == blockIdx.x) dbg_stamp(p, 2);
                    if (p.mn_major) {
                        const uint64_t adesc = make_umma_desc_mnmajor_sw128(smem_u32(sA + stage * A_STAGE_BYTES));
                        const uint64_t bdesc = make_umma_desc_mnmajor_sw128(smem_u32(sB + stage * B_STAGE_BYTES));
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)      // UMMA_K = 16 = two 8-row groups of 1024 B -> +128 in the (addr>>4) field
                            umma_bf16(d_tmem, adesc + 128 * k, bdesc + 128 * k, IDESC | UMMA_IDESC_A_MN_MAJOR | UMMA_IDESC_B_MN_MAJOR,
                                      (it | k) != 0 ? 1u : 0u);
                    } else {
                    const uint64_t adesc = make_umma_desc_kmajor_sw128(smem_u32(sA + stage * A_STAGE_BYTES));
                    const uint64_t bdesc = make_umma_desc_kmajor_sw128(smem_u32(sB + stage * B_STAGE_BYTES));
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        // +32 B per UMMA_K=16 bf16 inside the swizzle atom -> +2 in the (addr>>4) field
                        umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, IDESC, (it | k) != 0 ? 1u : 0u);
                    }
                    }
                    umma_commit(&empty_bar[stage]);            // frees the smem slot when these MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tmem_full[acc]);                  // accumulator complete -> epilogue
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int quarter = (warp - 4) & 3;                    // TMEM lanes [32*quarter, 32*quarter+32)
        const int eset = (warp - 4) >> 2;                      // epilogue set: 32-column chunks [c_begin, c_end) of the tile
        constexpr int CHUNKS = BN / 32 / EPI;
        const int c_begin = eset * CHUNKS, c_end = c_begin + CHUNKS;
        int acc = 0; uint32_t acc_phase = 0;
        int sbuf = 0;
        for (int item = blockIdx.x; item < total_tiles; item += gridDim.x) {
            const int tile = item % mn_tiles;
            const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
            mbar_wait(&tmem_full[acc], acc_phase, p.err_flag, 104);
            tcgen05_fence_after();
            if (threadIdx.x == 128 && item == blockIdx.x) dbg_stamp(p, 3);
            const int row = m0 + quarter * 32 + lane;
            const uint32_t t_base = tmem_base + static_cast<uint32_t>(acc * BN) + (static_cast<uint32_t>(quarter * 32) << 16);
            if (p.sum_out) {
                epilogue_tile_sum(p, t_base, row, m0 + quarter * 32, n0, c_begin, c_end,
                                  reinterpret_cast<float*>(sStage + (eset * 4 + quarter) * 8192), reinterpret_cast<float*>(sStage + eset * 4 * 8192 + 4608),
                                  quarter, 1 + eset, lane);
            } else if (p.tma_store) {
                epilogue_tile_tma(p, tmC, t_base, row, m0 + quarter * 32, n0, c_begin, c_end,
                                  smem_u32(sStage) + static_cast<uint32_t>(eset * 4 + quarter) * 8192u, sbuf, lane);
            } else {
#pragma unroll 1
            for (int c0 = c_begin * 32; c0 < c_end * 32; c0 += 32) {
                if (n0 + c0 >= p.N) break;                     // warp-uniform
                uint32_t v[32];
                tmem_ld32(t_base + static_cast<uint32_t>(c0), v);
                tmem_ld_wait();
                if (row < p.M) epilogue_store(p, row, n0 + c0, v);
            }
            }
            tcgen05_fence_before();
            mbar_arrive(&tmem_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        // the staging buffers must have been READ by the TMA unit before the CTA exits; the writes themselves complete as part of the
        // grid (waiting for them here, wait_group 0, costs every CTA a store round trip at the tail of every GEMM)
        if (p.tma_store && lane == 0) { if (g_tail_wait_full) bulk_wait_all(); else bulk_wait_read_all(); }
    }
    if (threadIdx.x == 128) dbg_stamp(p, 4);
    tcgen05_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) dbg_stamp(p, 5);
    if (warp == 2) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): one cluster of two CTAs (the two SMs of a TPC) per 256 x 256 output tile.
//
// Why: every GEMM on this path is bound by L2 -> SM operand traffic: a 128 x 256 x 64 k-block of the single-CTA kernel
// stages 16 KB of A + 32 KB of B per SM.  Here CTA r of the pair stages rows [m0 + 128 r, +128) of A (16 KB) and only rows
// [n0 + 128 r, +128) of W (16 KB, half of the B tile); the leader CTA (rank 0) issues ONE tcgen05.mma.cta_group::2 of M = 256,
// N = 256 per UMMA_K that reads both CTAs' shared memory, and each CTA's TMEM receives the 128 x 256 accumulator of its own rows:
// 32 KB instead of 48 KB of operands per SM for the same MACs (-33 %), and half the B shared-memory reads per SM.
//
// Protocol (per CTA: same warp roles as the single-CTA kernel):
//   full[s]   lives in the LEADER: count 1 (its own producer's arrive.expect_tx of BOTH CTAs' bytes); both producers' TMA loads
//             complete_tx on it (cp.async.bulk.tensor...cta_group::2 with the leader's barrier address)
//   empty[s]  per CTA, count 1: the leader's tcgen05.commit.cta_group::2...multicast arrives on both CTAs' copies
//   tmem_full[a]  per CTA, count 1: multicast commit after the tile's last MMA
//   tmem_empty[a] in the LEADER, count 2 x 4 EPI: one arrive per epilogue warp of either CTA (remote arrive for the peer)
// Setup / teardown are cluster-synchronised: barriers are initialised before any remote arrive, TMEM (cta_group::2 alloc, issued by
// warp 2 of both CTAs) is released only after both CTAs are done.
// Operands: K-major (every forward / dX GEMM) or MN-major (mn_major: the weight-gradient contraction dW = dZ^T . X read in place, 3-D TMA
// boxes of 64 k-rows x 64 columns, MN-major UMMA descriptors), optionally split over K (work item = (tile, K range); partial tiles are
// added with vector atomics by the direct-store epilogue); or K-major with the A rows gathered slot by slot from the frame-feature arena
// (each CTA's producer warp issues its own 128 rows as 3-D TMA boxes, one lane per slot).  Restriction: N % 256 == 0 (the launcher falls
// back otherwise).
// ------------------------------------------------------------------------------------------------
template <int STAGES, int EPI>
struct GemmSmem2 {
    static constexpr int BN = 256;
    static constexpr int B_STAGE_BYTES = (BN / 2) * BK * 2;                      // this CTA's half of the B tile
    static constexpr int TILE_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES);
    static constexpr int STAGING_BYTES = EPI * 4 * 2 * 4096;
    static constexpr int BAR_BYTES = (2 * STAGES + 4) * 8 + 16;
    static constexpr int TOTAL = TILE_BYTES + STAGING_BYTES + BAR_BYTES + 1024;
};

template <int STAGES, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS + 128 * (EPI - 1), 1)
gemm_tcgen05_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
    using S = GemmSmem2<STAGES, EPI>;
    constexpr int BN = 256;
    constexpr int B_STAGE_BYTES = S::B_STAGE_BYTES;
    constexpr uint32_t TMEM_COLS = 512;                             // two 256-column accumulator stages
    constexpr uint32_t IDESC = make_idesc_bf16(2 * BM, BN);         // M = 256 across the pair

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);       // same offset in both CTAs (same kernel, same static layout)
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
    uint8_t* sStage = smem + S::TILE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::TILE_BYTES + S::STAGING_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();                        // 0 = leader (issues the MMAs), 1 = peer
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int tiles_n = p.N / BN;
    const int tiles_m = (p.M + 2 * BM - 1) / (2 * BM);
    const int mn_tiles = tiles_m * tiles_n;
    const int total_tiles = mn_tiles * p.ksplit;                    // work item = (output tile, K split)

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
        if (p.tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmC)) : "memory");
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 2 * 4 * EPI); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();                                             // both CTAs' barriers initialised, both TMEM allocations made
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        if (p.a_slots) {
            // ===================== TMA producer, gather mode (whole warp: lane j loads slot j of this CTA's 128 A rows) ===============
            // K-major, one plane.  The leader expects the bytes of BOTH CTAs' valid slots (the last M tile may hold fewer, or none for the peer).
            const int spt = BM / p.slot_rows;                       // slots per CTA tile
            int stage = 0; uint32_t phase = 0;
            for (int item = pair; item < total_tiles; item += npairs) {
                const int tile = item % mn_tiles;
                const int m_pair = (tile / tiles_n) * (2 * BM);
                const int nb = (tile % tiles_n) * BN + static_cast<int>(rank) * (BN / 2);
                const int s_pair = m_pair / p.slot_rows;
                const int nv0 = max(0, min(spt, p.num_slots - s_pair)), nv1 = max(0, min(spt, p.num_slots - s_pair - spt));
                const int s0 = s_pair + static_cast<int>(rank) * spt, nvalid = rank == 0 ? nv0 : nv1;
                const int my_slot = lane < nvalid ? __ldg(p.a_slots + s0 + lane) : -1;
                const uint32_t tx_bytes = static_cast<uint32_t>((nv0 + nv1) * p.slot_rows * BK * 2) + 2 * B_STAGE_BYTES;
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    uint32_t full_leader = 0;
                    if (lane == 0) {
                        mbar_wait(&empty_bar[stage], phase ^ 1, p.err_flag, 201);
                        full_leader = mapa_shared(smem_u32(&full_bar[stage]), 0);
                        if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
                        tma_load_2d_pair(sB + stage * B_STAGE_BYTES, &tmB, full_leader, kb * BK, nb);
                    }
                    full_leader = __shfl_sync(0xffffffffu, full_leader, 0);      // also orders the slot loads after lane 0's empty-barrier wait
                    if (my_slot >= 0)
                        tma_load_3d_pair(sA + stage * A_STAGE_BYTES + lane * p.slot_rows * (BK * 2), &tmA, full_leader, kb * BK, 0, my_slot);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        } else if (lane == 0) {
            // ===================== TMA producer (both CTAs; bytes are signalled on the leader's full barrier) =====================
            int stage = 0; uint32_t phase = 0;
            for (int item = pair; item < total_tiles; item += npairs) {
                const int tile = item % mn_tiles, ks = item / mn_tiles;
                const int kb0 = ks * p.kb_per_split, kb1 = min(p.num_kb, kb0 + p.kb_per_split);
                const int m0 = (tile / tiles_n) * (2 * BM) + static_cast<int>(rank) * BM;
                const int nb = (tile % tiles_n) * BN + static_cast<int>(rank) * (BN / 2);
                for (int seg = 0; seg < p.nseg; ++seg) {
                    const int a_row = m0 + (p.nseg > 1 ? c_seg_a[seg] * p.a_plane_rows : 0);
                    const int b_row = nb + (p.nseg > 1 ? c_seg_b[seg] * p.w_plane_rows : 0);
                    const int pa = p.nseg > 1 ? c_seg_a[seg] : 0, pb = p.nseg > 1 ? c_seg_b[seg] : 0;
                    for (int kb = kb0; kb < kb1; ++kb) {
                        mbar_wait(&empty_bar[stage], phase ^ 1, p.err_flag, 201);
                        const uint32_t full_leader = mapa_shared(smem_u32(&full_bar[stage]), 0);
                        if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * (A_STAGE_BYTES + B_STAGE_BYTES));
                        if (p.mn_major) {                           // [64 k-rows][64 columns] boxes; rows past K / columns past M, N are zero-filled
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                tma_load_3d_pair(sB + stage * B_STAGE_BYTES + j * 8192, &tmB, full_leader, nb + 64 * j, kb * BK, pb);
                                tma_load_3d_pair(sA + stage * A_STAGE_BYTES + j * 8192, &tmA, full_leader, m0 + 64 * j, kb * BK, pa);
                            }
                        } else {
                            tma_load_2d_pair(sB + stage * B_STAGE_BYTES, &tmB, full_leader, kb * BK, b_row);
                            tma_load_2d_pair(sA + stage * A_STAGE_BYTES, &tmA, full_leader, kb * BK, a_row);
                        }
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            // ===================== MMA issuer (leader CTA only) =====================
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int item = pair; item < total_tiles; item += npairs) {
                const int ks = item / mn_tiles;
                const int kiters = p.nseg * (min(p.num_kb, (ks + 1) * p.kb_per_split) - ks * p.kb_per_split);
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1, p.err_flag, 202);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
                for (int it = 0; it < kiters; ++it) {
                    mbar_wait(&full_bar[stage], phase, p.err_flag, 203);
                    tcgen05_fence_after();
                    if (p.mn_major) {
                        const uint64_t adesc = make_umma_desc_mnmajor_sw128(smem_u32(sA + stage * A_STAGE_BYTES));
                        const uint64_t bdesc = make_umma_desc_mnmajor_sw128(smem_u32(sB + stage * B_STAGE_BYTES));
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            umma_bf16_pair(d_tmem, adesc + 128 * k, bdesc + 128 * k, IDESC | UMMA_IDESC_A_MN_MAJOR | UMMA_IDESC_B_MN_MAJOR, (it | k) != 0 ? 1u : 0u);
                    } else {
                    const uint64_t adesc = make_umma_desc_kmajor_sw128(smem_u32(sA + stage * A_STAGE_BYTES));
                    const uint64_t bdesc = make_umma_desc_kmajor_sw128(smem_u32(sB + stage * B_STAGE_BYTES));
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) umma_bf16_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, IDESC, (it | k) != 0 ? 1u : 0u);
                    }
                    umma_commit_pair(&empty_bar[stage], 3);         // frees the stage in BOTH CTAs when these MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit_pair(&tmem_full[acc], 3);               // accumulators complete -> both CTAs' epilogues
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue (each CTA drains its own 128 rows) =====================
        const int quarter = (warp - 4) & 3;
        const int eset = (warp - 4) >> 2;
        constexpr int CHUNKS = BN / 32 / EPI;
        const int c_begin = eset * CHUNKS, c_end = c_begin + CHUNKS;
        int acc = 0; uint32_t acc_phase = 0;
        int sbuf = 0;
        for (int item = pair; item < total_tiles; item += npairs) {
            const int tile = item % mn_tiles;
            const int m0 = (tile / tiles_n) * (2 * BM) + static_cast<int>(rank) * BM, n0 = (tile % tiles_n) * BN;
            mbar_wait(&tmem_full[acc], acc_phase, p.err_flag, 204);
            tcgen05_fence_after();
            const int row = m0 + quarter * 32 + lane;
            const uint32_t t_base = tmem_base + static_cast<uint32_t>(acc * BN) + (static_cast<uint32_t>(quarter * 32) << 16);
            if (p.sum_out) {
                epilogue_tile_sum(p, t_base, row, m0 + quarter * 32, n0, c_begin, c_end,
                                  reinterpret_cast<float*>(sStage + (eset * 4 + quarter) * 8192), reinterpret_cast<float*>(sStage + eset * 4 * 8192 + 4608),
                                  quarter, 1 + eset, lane);
            } else if (p.tma_store) {
                if (m0 + quarter * 32 < p.M)
                    epilogue_tile_tma(p, tmC, t_base, row, m0 + quarter * 32, n0, c_begin, c_end,
                                      smem_u32(sStage) + static_cast<uint32_t>(eset * 4 + quarter) * 8192u, sbuf, lane);
            } else {                                                // accumulating / split-K outputs: direct (atomic) stores per row
#pragma unroll 1
                for (int c0 = c_begin * 32; c0 < c_end * 32; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(t_base + static_cast<uint32_t>(c0), v);
                    tmem_ld_wait();
                    if (row < p.M) epilogue_store(p, row, n0 + c0, v);
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {                                        // one arrive per warp on the LEADER's tmem_empty
                if (rank == 0) mbar_arrive(&tmem_empty[acc]);
                else mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty[acc]), 0));
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (p.tma_store && lane == 0) bulk_wait_all();
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();                                             // the leader's MMAs read the peer's shared memory and write its TMEM
    if (warp == 2) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// Plain CUDA-core kernel with identical semantics.  Debug aid only (selected with stair_set_gemm_impl(1));
// it lets every other kernel be validated on the GPU independently of the tcgen05 path.
// ------------------------------------------------------------------------------------------------
__global__ void gemm_simt_kernel(const bf16* __restrict__ A, long long lda, const bf16* __restrict__ W, long long ldw, GemmParams p) {
    __shared__ float sa[16][17], sb[16][17];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int row = blockIdx.y * 16 + ty, col = blockIdx.x * 16 + tx;
    float acc = 0.f;
    for (int seg = 0; seg < p.nseg; ++seg) {
        const long long a_off = p.nseg > 1 ? static_cast<long long>(c_seg_a[seg]) * p.a_plane_rows : 0;
        const long long b_off = p.nseg > 1 ? static_cast<long long>(c_seg_b[seg]) * p.w_plane_rows : 0;
        for (int k0 = 0; k0 < p.K; k0 += 16) {
            const int ar = blockIdx.y * 16 + ty, br = blockIdx.x * 16 + ty;
            if (p.mn_major) {
                sa[ty][tx] = (ar < p.M && k0 + tx < p.K) ? __bfloat162float(A[(a_off + k0 + tx) * lda + ar]) : 0.f;
                sb[ty][tx] = (br < p.N && k0 + tx < p.K) ? __bfloat162float(W[(b_off + k0 + tx) * ldw + br]) : 0.f;
            } else {
            sa[ty][tx] = (ar < p.M && k0 + tx < p.K) ? __bfloat162float(A[(a_off + ar) * lda + k0 + tx]) : 0.f;
            sb[ty][tx] = (br < p.N && k0 + tx < p.K) ? __bfloat162float(W[(b_off + br) * ldw + k0 + tx]) : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 16; ++k) acc += sa[ty][k] * sb[tx][k];
            __syncthreads();
        }
    }
    if (row < p.M && col < p.N) {
        float x = acc * (p.row_scale ? p.row_scale[row] : 1.f);
        if (p.bias) x += p.bias[col];
        if (p.act == STAIR_ACT_RELU) x = fmaxf(x, 0.f);
        if (p.drop.thresh)
            x = drop_keep(drop_row_hash(p.drop.key_lo, p.drop.key_hi, p.drop.row0 + row), static_cast<uint32_t>(col), p.drop.thresh) ? x * p.drop.scale : 0.f;
        if (p.out_dtype == STAIR_BF16) reinterpret_cast<bf16*>(p.C)[static_cast<long long>(row) * p.ldc + col] = __float2bfloat16_rn(x);
        else {
            float* o = reinterpret_cast<float*>(p.C) + static_cast<long long>(row) * p.ldc + col;
            *o = p.accumulate ? *o + x : x;
        }
    }
}

// slot arena [slots][slot_rows][cols] with row pitch `row_pitch_elems`; one box = one slot x 64 columns
static int make_tmap_bf16_slots(CUtensorMap* tm, const void* base, uint64_t cols, uint64_t slot_rows, uint64_t slots,
                                uint64_t row_pitch_elems) {
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) return STAIR_ERR_CUDA;
    cuuint64_t gdim[3] = {cols, slot_rows, slots};
    cuuint64_t gstride[2] = {row_pitch_elems * 2, row_pitch_elems * 2 * slot_rows};
    cuuint32_t box[3] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(slot_rows), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? STAIR_OK : STAIR_ERR_ARG;
}

// MN-major operand [planes][plane_rows (k), ld] with `mn` valid columns and `k` valid rows per plane; box = 64 k-rows x 64 columns
static int make_tmap_bf16_mn(CUtensorMap* tm, const void* base, uint64_t mn, uint64_t k, uint64_t planes, uint64_t plane_rows, uint64_t ld) {
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) return STAIR_ERR_CUDA;
    cuuint64_t gdim[3] = {mn, k, planes};
    cuuint64_t gstride[2] = {ld * 2, ld * 2 * plane_rows};
    cuuint32_t box[3] = {64, 64, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? STAIR_OK : STAIR_ERR_ARG;
}

// output tile map for the epilogue's TMA stores: box = 32 rows x 128 bytes, SWIZZLE_128B
static int make_tmap_out(CUtensorMap* tm, const void* base, int out_dtype, uint64_t cols, uint64_t rows, uint64_t row_pitch_elems) {
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) return STAIR_ERR_CUDA;
    const uint64_t esz = out_dtype == STAIR_BF16 ? 2 : 4;
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {row_pitch_elems * esz};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / esz), 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, out_dtype == STAIR_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                     const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? STAIR_OK : STAIR_ERR_ARG;
}

static int g_gemm_impl = 0;      // 0 = tcgen05 (product), 1 = SIMT debug kernel
static int g_epilogue_impl = 0;  // 0 = staged TMA-store epilogue, 1 = direct per-row stores (debug / comparison)
static int* g_err_flag = nullptr;
static int g_num_sms = 0;
static int g_dw_wide = 0;        // 1 = 128 x 256 tiles for MN-major (weight-gradient) GEMMs (measured slower: 33.4 vs 30.7 us, profiles/micro_dw.py)
static int g_gemm_epi2 = 1;      // 1 = two epilogue warp sets for short-K GEMMs (num_kb <= 8); 0 = always one set (comparison)
static int g_wide_tiles_min = 1;  // 128 x 256 tiles when there are more than this many half-waves (SMs / 2) of 128 x 128 tiles.  Wide tiles move 25 %
                                  // fewer operand bytes per flop and these GEMMs are L2 -> SM bandwidth bound: measured at B = 4096, forward
                                  // 1.550 / 1.519 / 1.507 ms and training step 6.59 / 6.53 / 6.73 ms for 4 (two waves, the old rule) / 1 / 0
static int g_split_k = 1;        // 1 = split-K for accumulating GEMMs with few output tiles (weight gradients)
static unsigned long long* g_dbg = nullptr;

void err_flag_free() {
    if (g_err_flag) { cudaFreeHost(g_err_flag); g_err_flag = nullptr; }
}

int* err_flag_ptr() {
    if (!g_err_flag) {
        if (cudaMallocHost(reinterpret_cast<void**>(&g_err_flag), sizeof(int)) != cudaSuccess) return nullptr;   // pinned, device-visible
        *g_err_flag = 0;
    }
    return g_err_flag;
}

template <int BN, int STAGES, int EPI = 1, int OCC = 1>
static int launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const GemmParams& p, cudaStream_t st) {
    using S = GemmSmem<BN, STAGES, EPI>;
    static_assert(OCC == 1 || (OCC * (S::TOTAL + 1024) <= 228 * 1024 && OCC * 2 * BN <= 512), "co-resident CTAs must share the SM's shared memory and TMEM");
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, STAGES, EPI, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL) != cudaSuccess)
            return STAIR_ERR_CUDA;
        if (OCC > 1 && cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, STAGES, EPI, OCC>, cudaFuncAttributePreferredSharedMemoryCarveout, 100) != cudaSuccess)
            return STAIR_ERR_CUDA;
        configured = true;
    }
    const int tiles = ceil_div(p.M, BM) * ceil_div(p.N, BN);
    GemmParams q = p;
    // split-K for accumulating GEMMs with few output tiles and a long contraction (weight gradients: [N,K] outputs of a few
    // tiles, contraction over thousands of rows): every SM gets a K range, partial tiles are added with vector atomics
    if (p.accumulate && !p.bias && !p.row_scale && p.act == STAIR_ACT_NONE && g_split_k && tiles * 2 <= g_num_sms && p.num_kb >= 8) {
        int ks = g_num_sms / tiles;
        if (ks > p.num_kb / 4) ks = p.num_kb / 4;
        if (ks > 1) {
            q.kb_per_split = ceil_div(p.num_kb, ks);
            q.ksplit = ceil_div(p.num_kb, q.kb_per_split);
        }
    }
    const int items = tiles * q.ksplit;
    const int grid = items < OCC * g_num_sms ? items : OCC * g_num_sms;
    gemm_tcgen05_kernel<BN, STAGES, EPI, OCC><<<grid, GEMM_THREADS + 128 * (EPI - 1), S::TOTAL, st>>>(ta, tb, tc, q);
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

static int pair_mode_default() {                       // STAIR_GEMM_PAIR=0|1|2 overrides the default at library load (A/B runs)
    const char* e = getenv("STAIR_GEMM_PAIR");
    return (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1;
}
static int g_pair_gather = 1;                       // 1 = the CTA-pair kernel also serves gathered (frame-arena) A operands
static int g_pair_mn = 1;                           // 1 = the CTA-pair kernel also serves MN-major (weight-gradient) GEMMs; 0 = K-major only (comparison)
static int g_pair_mode = pair_mode_default();      // 1 (default) = CTA-pair (cta_group::2) kernel for eligible GEMMs, 0 = never, 2 = whenever legal (tests)

template <int STAGES, int EPI>
static int launch_tc_pair(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const GemmParams& p, cudaStream_t st) {
    using S = GemmSmem2<STAGES, EPI>;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(gemm_tcgen05_pair_kernel<STAGES, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL) != cudaSuccess)
            return STAIR_ERR_CUDA;
        configured = true;
    }
    const int tiles = ceil_div(p.M, 2 * BM) * (p.N / 256);
    const int max_pairs = g_num_sms / 2;
    GemmParams q = p;
    // split-K for accumulating GEMMs with few output tiles and a long contraction (weight gradients), as in launch_tc
    if (p.accumulate && !p.bias && !p.row_scale && p.act == STAIR_ACT_NONE && g_split_k && tiles * 2 <= max_pairs && p.num_kb >= 8) {
        int ks = max_pairs / tiles;
        if (ks > p.num_kb / 4) ks = p.num_kb / 4;
        if (ks > 1) {
            q.kb_per_split = ceil_div(p.num_kb, ks);
            q.ksplit = ceil_div(p.num_kb, q.kb_per_split);
        }
    }
    const int items = tiles * q.ksplit;
    const int pairs = items < max_pairs ? items : max_pairs;
    gemm_tcgen05_pair_kernel<STAGES, EPI><<<2 * pairs, GEMM_THREADS + 128 * (EPI - 1), S::TOTAL, st>>>(ta, tb, tc, q);
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

thread_local long long g_launch_count = 0;
bool gemm_sum_epilogue_ok(int T) { return T == 1 || T == 2 || T == 4 || T == 8 || T == 16 || T == 32 || T == 64 || T == 128; }
static int small_default() { const char* e = getenv("STAIR_GEMM_SMALL"); return (e && (e[0] == '0' || e[0] == '1')) ? e[0] - '0' : 0; }
static int g_gemm_small = small_default();      // 1 = module-sized GEMMs (N <= 1024, at most a few waves of tiles) use the two-CTAs-per-SM form.  Off:
                                                // bit-identical and no faster (1.331 vs 1.337 ms per forward, profiles/r2_module_phase_analysis.txt)

int launch_gemm(const GemmArgs& a, cudaStream_t st) {
    if (a.M <= 0 || a.N <= 0) return STAIR_OK;
    if (a.K <= 0 || (a.nplanes != 1 && a.nplanes != 3) || (a.lda % 8) || (a.ldw % 8)) return STAIR_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(a.A) & 15) || (reinterpret_cast<uintptr_t>(a.W) & 15)) return STAIR_ERR_ARG;
    if (a.accumulate && a.out_dtype != STAIR_F32) return STAIR_ERR_ARG;
    const bool gather = a.a_slots != nullptr;
    if (gather && (a.nplanes != 1 || !gemm_gather_ok(a.slot_rows) || a.M % a.slot_rows)) return STAIR_ERR_ARG;
    if (a.mn_major && (gather || (a.nplanes > 1 && (a.a_plane_rows < a.K || a.w_plane_rows < a.K)))) return STAIR_ERR_ARG;
    GemmParams p;
    p.M = a.M; p.N = a.N; p.K = a.K; p.num_kb = ceil_div(a.K, BK); p.nseg = a.nplanes == 3 ? 6 : 1;
    p.ksplit = 1; p.kb_per_split = p.num_kb;
    p.a_plane_rows = a.a_plane_rows; p.w_plane_rows = a.w_plane_rows;
    p.bias = a.bias; p.row_scale = a.row_scale; p.C = a.C; p.ldc = a.ldc; p.out_dtype = a.out_dtype; p.act = a.act;
    p.accumulate = a.accumulate;
    p.drop = a.drop;
    p.mn_major = a.mn_major;
    p.atomic_acc = (a.atomic_acc && a.accumulate) ? 1 : 0;
    const int esz = a.out_dtype == STAIR_BF16 ? 2 : 4;
    p.vec_ok = ((reinterpret_cast<uintptr_t>(a.C) & 15) == 0 && (a.ldc * esz) % 16 == 0) ? 1 : 0;
    p.tma_store = (!a.accumulate && p.vec_ok && g_epilogue_impl == 0) ? 1 : 0;
    p.a_slots = a.a_slots; p.slot_rows = gather ? a.slot_rows : BM; p.num_slots = gather ? a.M / a.slot_rows : 0;
    p.sum_out = reinterpret_cast<bf16*>(a.sum_out); p.ld_sum = a.ld_sum; p.sum_T = a.sum_T;
    if (a.sum_out) {
        if (!gemm_sum_epilogue_ok(a.sum_T) || a.M % a.sum_T || a.accumulate || a.mn_major || g_gemm_impl == 1) return STAIR_ERR_ARG;
        p.tma_store = 0;                                           // C is never written in this mode
    }
    p.err_flag = err_flag_ptr();
    p.dbg = g_dbg;

    if (g_gemm_impl == 1 && !gather) {
        dim3 grid(ceil_div(a.N, 16), ceil_div(a.M, 16)), block(16, 16);
        gemm_simt_kernel<<<grid, block, 0, st>>>(reinterpret_cast<const bf16*>(a.A), a.lda, reinterpret_cast<const bf16*>(a.W), a.ldw, p);
        STAIR_CHECK_LAUNCH();
        return STAIR_OK;
    }
    if (!g_num_sms) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return STAIR_ERR_CUDA;
        if (cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return STAIR_ERR_CUDA;
    }
    const uint64_t a_rows = static_cast<uint64_t>(a.nplanes - 1) * a.a_plane_rows + a.M;
    const uint64_t w_rows = static_cast<uint64_t>(a.nplanes - 1) * a.w_plane_rows + a.N;
    const int tiles128 = ceil_div(a.M, BM) * ceil_div(a.N, 128);
    // wide tiles when there are plenty of them.  The weight-gradient contraction (16 output tiles, K ~ 20 000 split over all SMs) is bound
    // by L2 -> SM operand traffic (every 128-column operand slice is re-read by 4 tiles: 160 MB for 40 MB of operands, ~6 TB/s); 128 x 256
    // tiles move 25 % fewer bytes but double the same-address atomics of the split-K epilogue and measured slower (g_dw_wide).
    // CTA-pair kernel: no gather, N a multiple of 256, and enough 256 x 256 tiles (a quarter wave) — or, for the accumulating weight-gradient
    // contractions (few output tiles, K in the tens of thousands), enough K to split over the pairs
    const bool small = g_gemm_small && a.nplanes == 1 && !a.mn_major && !a.accumulate && a.N > 64 && a.N <= 1024 && tiles128 <= 8 * g_num_sms;
    if (!small) {
        const int tiles256 = ceil_div(a.M, 2 * BM) * (a.N / 256);
        const bool legal = (!gather || g_pair_gather) && (p.vec_ok || a.sum_out) && g_epilogue_impl == 0 && a.N % 256 == 0 && g_num_sms >= 2 && (!a.mn_major || g_pair_mn);
        // (split-K weight gradients with a handful of 256 x 256 output tiles stay on the single-CTA kernel unless forced: every K split adds a
        // whole 256 KB tile with atomics, twice the atomic traffic of 128 x 128 tiles at the same SM count — measured 6.37 -> 6.55 ms per
        // training step, profiles/r2_train_pair_ab.txt)
        const bool enough = tiles256 * 4 >= g_num_sms;
        if (legal && (g_pair_mode == 2 || (g_pair_mode == 1 && enough))) {
            CUtensorMap ta, tb, tc;
            int rc;
            if (a.mn_major) {
                rc = make_tmap_bf16_mn(&ta, a.A, a.M, a.K, a.nplanes, a.nplanes > 1 ? a.a_plane_rows : a.K, a.lda);
                if (rc) return rc;
                rc = make_tmap_bf16_mn(&tb, a.W, a.N, a.K, a.nplanes, a.nplanes > 1 ? a.w_plane_rows : a.K, a.ldw);
            } else {
                rc = gather ? make_tmap_bf16_slots(&ta, a.A, a.K, a.slot_rows, a.arena_slots, a.lda)
                            : make_tmap_bf16_2d(&ta, a.A, a.K, a_rows, a.lda, BK, BM);
                if (rc) return rc;
                rc = make_tmap_bf16_2d(&tb, a.W, a.K, w_rows, a.ldw, BK, 128);
            }
            if (rc) return rc;
            tc = tb;
            if (p.tma_store) {
                rc = make_tmap_out(&tc, a.C, a.out_dtype, a.N, a.M, a.ldc);
                if (rc) return rc;
            }
            const bool epi2 = g_gemm_epi2 && p.num_kb * p.nseg <= 8 && !a.mn_major;
            return epi2 ? launch_tc_pair<5, 2>(ta, tb, tc, p, st) : launch_tc_pair<6, 1>(ta, tb, tc, p, st);
        }
    }
    const int bn = a.N <= 64 ? 64 : ((!small && a.N % 256 == 0 && (tiles128 > g_wide_tiles_min * g_num_sms / 2 || (a.mn_major && g_dw_wide))) ? 256 : 128);
    CUtensorMap ta, tb;
    int rc;
    if (a.mn_major) {
        rc = make_tmap_bf16_mn(&ta, a.A, a.M, a.K, a.nplanes, a.nplanes > 1 ? a.a_plane_rows : a.K, a.lda);
        if (rc) return rc;
        rc = make_tmap_bf16_mn(&tb, a.W, a.N, a.K, a.nplanes, a.nplanes > 1 ? a.w_plane_rows : a.K, a.ldw);
    } else {
        rc = gather ? make_tmap_bf16_slots(&ta, a.A, a.K, a.slot_rows, a.arena_slots, a.lda)
                    : make_tmap_bf16_2d(&ta, a.A, a.K, a_rows, a.lda, BK, BM);
        if (rc) return rc;
        rc = make_tmap_bf16_2d(&tb, a.W, a.K, w_rows, a.ldw, BK, bn);
    }
    if (rc) return rc;
    CUtensorMap tc = tb;
    if (p.tma_store) {
        rc = make_tmap_out(&tc, a.C, a.out_dtype, a.N, a.M, a.ldc);
        if (rc) return rc;
    }
    // short K loops are epilogue-bound: two epilogue warp sets, one pipeline stage fewer (see GemmSmem)
    const bool epi2 = g_gemm_epi2 && p.num_kb * p.nseg <= 8 && !a.mn_major;
    if (small) return launch_tc<128, 2, 1, 2>(ta, tb, tc, p, st);
    if (bn == 64) return launch_tc<64, 8>(ta, tb, tc, p, st);
    if (bn == 256) return epi2 ? launch_tc<256, 3, 2>(ta, tb, tc, p, st) : launch_tc<256, 4>(ta, tb, tc, p, st);
    return epi2 ? launch_tc<128, 5, 2>(ta, tb, tc, p, st) : launch_tc<128, 6>(ta, tb, tc, p, st);
}

}  // namespace stair

using namespace stair;

extern "C" int stair_set_gemm_impl(int impl) { g_gemm_impl = impl; return STAIR_OK; }
extern "C" int stair_set_gemm_pair_gather(int on) { g_pair_gather = on ? 1 : 0; return STAIR_OK; }
extern "C" int stair_set_gemm_pair_mn(int on) { g_pair_mn = on ? 1 : 0; return STAIR_OK; }
extern "C" int stair_set_gemm_pair(int mode) { g_pair_mode = mode < 0 ? 0 : (mode > 2 ? 2 : mode); return STAIR_OK; }
extern "C" int stair_set_gemm_small(int on) { g_gemm_small = on ? 1 : 0; return STAIR_OK; }
extern "C" int stair_set_gemm_epi2(int on) { g_gemm_epi2 = on ? 1 : 0; return STAIR_OK; }
extern "C" int stair_set_gemm_wide_min(int half_waves) { g_wide_tiles_min = half_waves < 0 ? 0 : half_waves; return STAIR_OK; }
extern "C" int stair_set_gemm_dw_wide(int on) { g_dw_wide = on ? 1 : 0; return STAIR_OK; }
extern "C" int stair_set_gemm_split_k(int on) { g_split_k = on ? 1 : 0; return STAIR_OK; }
extern "C" int stair_gemm_debug_timeline(unsigned long long* dev_buf) { g_dbg = dev_buf; return STAIR_OK; }
extern "C" int stair_set_gemm_epilogue(int impl) { g_epilogue_impl = impl; return STAIR_OK; }
extern "C" int stair_get_gemm_impl() { return g_gemm_impl; }
extern "C" int stair_gemm_error_flag() { return g_err_flag ? *g_err_flag : 0; }

// C = act(row_scale * (A · W^T) + bias).  A: bf16 [nplanes][a_plane_rows, lda], W: bf16 [nplanes][w_plane_rows, ldw].
extern "C" int stair_gemm_bf16(const void* A, long long lda, int a_plane_rows, const void* W, long long ldw, int w_plane_rows,
                               int nplanes, const float* bias, const float* row_scale, void* C, long long ldc, int out_dtype,
                               int M, int N, int K, int act, int accumulate, void* stream) {
    GemmArgs a;
    a.A = A; a.lda = lda; a.a_plane_rows = a_plane_rows; a.W = W; a.ldw = ldw; a.w_plane_rows = w_plane_rows; a.nplanes = nplanes;
    a.bias = bias; a.row_scale = row_scale; a.C = C; a.ldc = ldc; a.out_dtype = out_dtype; a.M = M; a.N = N; a.K = K;
    a.act = act; a.accumulate = accumulate;
    return launch_gemm(a, reinterpret_cast<cudaStream_t>(stream));
}

// Weight-gradient contraction C[M,N] (+)= A^T . W with A = [nplanes][a_plane_rows, lda] (element (k, m)), W = [nplanes][w_plane_rows, ldw]
// (element (k, n)): both operands are read in place as MN-major UMMA tiles (no transposed copies).  fp32 output.
extern "C" int stair_gemm_bf16_tn(const void* A, long long lda, int a_plane_rows, const void* W, long long ldw, int w_plane_rows, int nplanes,
                                  float* C, long long ldc, int M, int N, int K, int accumulate, void* stream) {
    GemmArgs a;
    a.A = A; a.lda = lda; a.a_plane_rows = a_plane_rows; a.W = W; a.ldw = ldw; a.w_plane_rows = w_plane_rows; a.nplanes = nplanes;
    a.C = C; a.ldc = ldc; a.out_dtype = STAIR_F32; a.M = M; a.N = N; a.K = K; a.accumulate = accumulate; a.mn_major = 1;
    return launch_gemm(a, reinterpret_cast<cudaStream_t>(stream));
}

// sum_out[i, :] = sum over the T rows of instance i of bf16(act(A . W^T + bias))  (GemmArgs::sum_out: Filter's frame sum as a GEMM epilogue)
extern "C" int stair_gemm_bf16_framesum(const void* A, long long lda, const void* W, long long ldw, const float* bias, void* sum_out, long long ld_sum,
                                        int M, int N, int K, int act, int T, void* stream) {
    GemmArgs a;
    a.A = A; a.lda = lda; a.W = W; a.ldw = ldw; a.w_plane_rows = N; a.bias = bias; a.M = M; a.N = N; a.K = K; a.act = act;
    a.out_dtype = STAIR_BF16; a.sum_out = sum_out; a.ld_sum = ld_sum; a.sum_T = T;
    return launch_gemm(a, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int stair_gemm_bf16_gather(const void* arena, long long ld, long long arena_slots, const int32_t* a_slots, int slot_rows,
                                      const void* W, long long ldw, const float* bias, const float* row_scale, void* C, long long ldc,
                                      int out_dtype, int M, int N, int K, int act, void* stream) {
    GemmArgs a;
    a.A = arena; a.lda = ld; a.arena_slots = arena_slots; a.a_slots = a_slots; a.slot_rows = slot_rows; a.W = W; a.ldw = ldw;
    a.bias = bias; a.row_scale = row_scale; a.C = C; a.ldc = ldc; a.out_dtype = out_dtype; a.M = M; a.N = N; a.K = K; a.act = act;
    return launch_gemm(a, reinterpret_cast<cudaStream_t>(stream));
}
