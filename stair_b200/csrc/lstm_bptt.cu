// Persistent fused BPTT of the BiLSTM encoders (bf16 training path): the whole backward time loop of nn.LSTM
// (video_nmn/module_net.py:39-47, 147-163; autograd of train_module.py:408) in ONE launch for both encoders and both directions.
// Replaces, per step and direction, a cell kernel + a recurrent GEMM launch (2 x (T + L_max) launches of 15-30 us each on the
// critical path of the training step).
//
// One CTA owns 64 questions of one (encoder, direction) for all steps, walking s = S-1 .. 0:
//   dh_s      = d(encoder output)[token row] (+ dh_rec from step s+1: TMEM); d question_feature is added to the rows of each
//               question's last step by a small kernel before the loop
//   gate math = the derivative of the cell (lstm.cu / train_kernels.cu lstm_cell_bwd_kernel) from the forward's blocked bf16
//               coefficient history (train_kernels.cuh lstm_hist_coef_off: six multiplies per unit, no transcendentals) and the
//               running dc (fp32, private, coalesced)
//   dG_s      -> bf16 rows [rows][8h] in token order (the operand of dW_ih / dW_hh / bias, executor_bwd.cu) and, as the A operand of
//               dh_rec_{s-1} = dG_s . W_hh, into shared memory: chunk c = the four gates of hidden units 64c .. 64c+63 = four k-blocks
//               (one per gate) of a [128 x 256] K-major SWIZZLE_128B tile, double-buffered by chunk parity
//   tcgen05.mma accumulates D[128 x h] over the 4 (or h/64) chunks while later chunks are still being produced; B tiles are rows of the
//   transposed W_hh copy (StairModel.wt, [h][4h]: box {64 k, 256 n} at k = gate*h + 64c) streamed through a TMA ring; D is
//   double-buffered in TMEM by step parity so the MMA of step s-1 never waits for the last reads of D_s.
//
// Warps (352 threads, 168 registers): 0-7 epilogue (all four TMEM lane quarters: rows 64-127 of the dG operand are a copy of rows 0-63,
// 4 column groups of 16 units per chunk), 8 TMA producer, 9 MMA issuer, 10 TMEM allocator.
// L2 prefetching was measured and removed twice: a warp running 1-3 steps ahead doubled the DRAM reads (the steps in flight of 148
// CTAs do not stay in L2) and prefetching 2-8 sub-block iterations ahead from the epilogue changed nothing (850 us either way).
#include "nmn_kernels.cuh"
#include "tc_ptx.cuh"
#include "train_kernels.cuh"

namespace stair {

namespace {

constexpr int LB_ROWS = 64;                        // questions per CTA (MMA M = 128, lanes 64-127 = a copy of lanes 0-63)
constexpr int LB_KB_BYTES = 128 * 64 * 2;          // one 64-wide k-block of the A tile (128 rows): 16 KiB
constexpr int LB_A_BYTES = 4 * LB_KB_BYTES;        // one chunk: 4 gates x 64 units
constexpr int LB_W_STAGE_BYTES = 256 * 64 * 2;     // one B tile (256 n x 64 k): 32 KiB
constexpr int LB_STAGES = 3;
constexpr int LB_THREADS = 352;                     // 11 warps: 8 epilogue + TMA + MMA + TMEM allocator (<= 3 warps per SM sub-partition: 168 registers)
constexpr int LB_EPI = 256;                        // epilogue threads
constexpr uint32_t LB_IDESC = make_idesc_bf16(128, 256);

struct BpttSeq {
    const bf16* coef_h;     // blocked bf16 coefficient history of the fused forward (train_kernels.cuh lstm_hist_coef_off)
    const float* dout;      // [rows][2h] gradient of the encoder output
    bf16* dxb;              // [rows][8h] gate pre-activation gradients, token order
    float* dc;              // running dc scratch [2][nblk][h][64]
    const int* q_off;       // text: [B+1]; video: null
    const int* order;       // text (optional): length-sorted schedule of the fused forward (lstm_fused.cu LstmSeq::order / soff): grid row r is
    const int* soff;        //   question order[r]; dxb rows (and the coefficient history, by grid row) in sorted order, dout rows in batch order
    int steps, B, h;
};
struct BpttParams { BpttSeq seq[2]; int* err_flag; };

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr) : "memory");
}
// operands of one sub-block iteration of the cell backward (8 hidden units of one question row)
struct BpttStage { uint4 co[LSTM_NCOEF]; float4 dh[2]; };
// j-th of the 8 bf16 values packed in a uint4
__device__ __forceinline__ float coef_at(const uint4& v, int j) {
    const uint32_t w = j < 2 ? v.x : j < 4 ? v.y : j < 6 ? v.z : v.w;
    return __uint_as_float((j & 1) ? (w & 0xFFFF0000u) : (w << 16));
}

__global__ void __launch_bounds__(LB_THREADS, 1)
lstm_bptt_kernel(const __grid_constant__ CUtensorMap tmW0, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmW3, const BpttParams p) {
    const BpttSeq sq = blockIdx.z == 0 ? p.seq[0] : p.seq[1];
    const int dir = blockIdx.y;
    const int wsel = blockIdx.z * 2 + dir;
    const int row0 = blockIdx.x * LB_ROWS;
    if (row0 >= sq.B) return;
    const int h = sq.h, NC = h / 64;
    const bool ragged = sq.q_off != nullptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Rows 64-127 of the dG operand hold a COPY of rows 0-63 (the MMA is M = 128 either way), so TMEM lanes 64-127 carry the same dh_rec
    // as lanes 0-63 and the 8 epilogue warps can sit on all four TMEM lane quarters = all four SM sub-partitions: quarters q and q + 2
    // share the rows (q & 1) * 32 .. + 31 and split the column groups.
    const bool is_epi = warp < 8;
    const int quarter = warp & 3, cgrp = (warp >> 2) * 2 + (quarter >> 1);          // column group 0..3: units 16*cgrp .. +15 of a chunk

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint8_t* sA = smem;                                            // [2][4 gates][128 x 64] bf16
    uint8_t* sW = smem + 2 * LB_A_BYTES;                           // [LB_STAGES][256 x 64] bf16
    uint64_t* w_full = reinterpret_cast<uint64_t*>(sW + LB_STAGES * LB_W_STAGE_BYTES);
    uint64_t* w_empty = w_full + LB_STAGES;
    uint64_t* a_ready = w_empty + LB_STAGES;        // [2] chunk buffer written by the epilogue
    uint64_t* a_empty = a_ready + 2;                // [2] MMAs reading the chunk buffer retired
    uint64_t* d_full = a_empty + 2;                 // [2] D of a step complete
    uint64_t* d_empty = d_full + 2;                 // [2] epilogue done reading D of a step
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(d_empty + 2);
    int* s_steps = reinterpret_cast<int*>(tmem_ptr_smem + 1);

    if (threadIdx.x == 0) {
        for (int s = 0; s < LB_STAGES; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&a_ready[a], LB_EPI); mbar_init(&a_empty[a], 1);
            mbar_init(&d_full[a], 1); mbar_init(&d_empty[a], LB_EPI);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        *s_steps = ragged ? 0 : sq.steps;
    }
    if (warp == 10) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_ptr_smem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 2 * LB_A_BYTES / 16; i += LB_THREADS) reinterpret_cast<uint4*>(sA)[i] = make_uint4(0, 0, 0, 0);   // rows 64-127 stay 0
    __syncthreads();
    if (ragged && threadIdx.x < LB_ROWS) {
        const int r = row0 + threadIdx.x;
        const int* off = sq.soff ? sq.soff : sq.q_off;
        if (r < sq.B) atomicMax(s_steps, __ldg(off + r + 1) - __ldg(off + r));
    }
    fence_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const int S = *s_steps;

    if (warp == 8) {
        if (lane == 0) {
            // ===================== TMA producer: B tiles of W_hh^T, (chunk, gate) order, every step that feeds an earlier one ==========
            int stage = 0; uint32_t phase = 0;
            for (int s = S - 1; s >= 1; --s)
                for (int c = 0; c < NC; ++c)
                    for (int g = 0; g < 4; ++g) {
                        mbar_wait(&w_empty[stage], phase ^ 1, p.err_flag, 401);
                        mbar_arrive_expect_tx(&w_full[stage], LB_W_STAGE_BYTES);
                        uint8_t* dst = sW + stage * LB_W_STAGE_BYTES;
                        const int k0 = g * h + c * 64;
                        if (wsel == 0) tma_load_2d(dst, &tmW0, &w_full[stage], k0, 0);
                        else if (wsel == 1) tma_load_2d(dst, &tmW1, &w_full[stage], k0, 0);
                        else if (wsel == 2) tma_load_2d(dst, &tmW2, &w_full[stage], k0, 0);
                        else tma_load_2d(dst, &tmW3, &w_full[stage], k0, 0);
                        if (++stage == LB_STAGES) { stage = 0; phase ^= 1; }
                    }
        }
    } else if (warp == 9) {
        if (lane == 0) {
            // ===================== MMA issuer: D_s[128, h] = dG_s[128, 4h] . W_hh  (chunk by chunk as the epilogue produces dG_s) ==========
            int stage = 0; uint32_t phase = 0;
            uint32_t ar_phase[2] = {0, 0};
            int it = 0;                                            // step counter (D buffer / phase bookkeeping)
            for (int s = S - 1; s >= 1; --s, ++it) {
                const int db = s & 1;
                // D buffer db was last read by the epilogue of step s+1 ... it holds D of step s+2: wait until those reads are done
                if (it >= 2) mbar_wait(&d_empty[db], static_cast<uint32_t>(((it - 2) >> 1) & 1), p.err_flag, 402);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(db * 256);
                for (int c = 0; c < NC; ++c) {
                    const int ab = c & 1;
                    mbar_wait(&a_ready[ab], ar_phase[ab], p.err_flag, 403);
                    ar_phase[ab] ^= 1;
                    tcgen05_fence_after();
                    const uint32_t abase = smem_u32(sA + ab * LB_A_BYTES);
                    for (int g = 0; g < 4; ++g) {
                        mbar_wait(&w_full[stage], phase, p.err_flag, 404);
                        tcgen05_fence_after();
                        const uint64_t adesc = make_umma_desc_kmajor_sw128(abase + g * LB_KB_BYTES);
                        const uint64_t bdesc = make_umma_desc_kmajor_sw128(smem_u32(sW + stage * LB_W_STAGE_BYTES));
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, LB_IDESC, (c | g | k) != 0 ? 1u : 0u);
                        umma_commit(&w_empty[stage]);
                        if (++stage == LB_STAGES) { stage = 0; phase ^= 1; }
                    }
                    umma_commit(&a_empty[ab]);
                }
                umma_commit(&d_full[db]);
            }
        }
    } else if (is_epi) {
        // ===================== cell backward: thread = question row, 16 units per chunk (two sub-blocks of 8) =========================
        // The step is a chain of 2 NC sub-block iterations, each needing ~100 bytes per thread from HBM (coefficients, output gradient)
        // and L2 (dc): the operands of iteration k+1 are requested before iteration k is processed (two register stages), so the
        // recurrence waits on the tensor core, not on memory.
        const int row = (quarter & 1) * 32 + lane;
        const int pos = row0 + row;                                 // grid row; carries question order[pos] when the text is length-sorted
        const bool valid = pos < sq.B;
        const int grow = (valid && sq.order) ? __ldg(sq.order + pos) : pos;
        int base = 0, xbase = 0, L = sq.steps;                      // first dout row (batch order) / first dxb row (schedule order)
        if (ragged) {
            base = valid ? __ldg(sq.q_off + grow) : 0; L = valid ? __ldg(sq.q_off + grow + 1) - base : 0;
            xbase = (valid && sq.soff) ? __ldg(sq.soff + pos) : base;
        } else xbase = base = grow * sq.steps;
        if (!valid) L = 0;
        constexpr uint32_t DUP = 64u * 128u;                       // byte offset of the copy of a row in the dG operand (row + 64)
        const int nblk = (sq.B + LB_ROWS - 1) / LB_ROWS;
        float* dcblk = sq.dc + (static_cast<long long>(dir) * nblk + blockIdx.x) * (static_cast<long long>(h) * LB_ROWS) + row * 4;   // [unit/4][row][4]
        const long long RB = (sq.B + 127) / 128 * 4;
        const long long hist_rb = (static_cast<long long>(dir) * RB + (pos >> 5)) * (h >> 3);
        const long long hist_step = 2 * RB * (h >> 3);
        const uint32_t sA0 = smem_u32(sA);
        const uint32_t rowoff = static_cast<uint32_t>(row) * 128u;
        const uint32_t sw = static_cast<uint32_t>(row & 7);

        auto load_stage = [&](BpttStage& st, int s, int c, int sb) {
            if (s < 0 || s >= L) return;
            const int u0 = c * 64 + cgrp * 16 + sb * 8;
            const bf16* cp = sq.coef_h + (s * hist_step + hist_rb + (u0 >> 3)) * (LSTM_NCOEF * 256) + lane * 8;
#pragma unroll
            for (int k = 0; k < LSTM_NCOEF; ++k) st.co[k] = ld_stream_v4(cp + k * 256);
            const float* drow = sq.dout + (static_cast<long long>(base) + (dir == 0 ? s : L - 1 - s)) * 2 * h + dir * h + u0;
            st.dh[0] = __ldg(reinterpret_cast<const float4*>(drow));
            st.dh[1] = __ldg(reinterpret_cast<const float4*>(drow + 4));
        };
        // running dc of one sub-block (written by this thread one step ago: L2).  One register set: the next iteration's values are
        // requested as soon as the current ones are consumed, half an iteration ahead of their use
        float4 dcr[2];
        auto load_dc = [&](int s, int c, int sb) {
            if (s < 0 || s + 1 >= L) return;
            const int u0 = c * 64 + cgrp * 16 + sb * 8;
            dcr[0] = *reinterpret_cast<const float4*>(dcblk + (u0 / 4) * (LB_ROWS * 4));
            dcr[1] = *reinterpret_cast<const float4*>(dcblk + (u0 / 4 + 1) * (LB_ROWS * 4));
        };
        int aw0 = 0, aw1 = 0;                                      // writes into each chunk buffer so far
        auto process = [&](const BpttStage& st, int s, int c, int sb, int dbuf, int ns, int nc, int nsb) {
            const int ab = c & 1;
            const int u0 = c * 64 + cgrp * 16 + sb * 8;
            const bool active = s < L, has_next = s + 1 < L;       // has_next: this row was active at step s+1, dh_rec / dc carry over
            uint32_t dr[8];
            if (s < S - 1) {                                        // .sync.aligned: the whole warp, converged
                tmem_ld8(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(dbuf * 256 + u0), dr);
                tmem_ld_wait();
            }
            const uint32_t ch = static_cast<uint32_t>((u0 & 63) >> 3);      // A operand of dh_rec_{s-1}: k-block = gate, 16-byte chunk = (u0 % 64) / 8
            const uint32_t a0 = sA0 + static_cast<uint32_t>(ab * LB_A_BYTES) + rowoff + ((ch ^ sw) << 4);
            if (active) {
                float dh8[8] = {st.dh[0].x, st.dh[0].y, st.dh[0].z, st.dh[0].w, st.dh[1].x, st.dh[1].y, st.dh[1].z, st.dh[1].w};
                float dct8[8] = {dcr[0].x, dcr[0].y, dcr[0].z, dcr[0].w, dcr[1].x, dcr[1].y, dcr[1].z, dcr[1].w};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (has_next) dh8[j] += __uint_as_float(dr[j]);
                    dct8[j] = fmaf(dh8[j], coef_at(st.co[LSTM_CO_A], j), has_next ? dct8[j] : 0.f);
                }
                load_dc(ns, nc, nsb);
                if (s > 0) {
                    float dcn[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) dcn[j] = dct8[j] * coef_at(st.co[LSTM_CO_F], j);
                    *reinterpret_cast<float4*>(dcblk + (u0 / 4) * (LB_ROWS * 4)) = make_float4(dcn[0], dcn[1], dcn[2], dcn[3]);
                    *reinterpret_cast<float4*>(dcblk + (u0 / 4 + 1) * (LB_ROWS * 4)) = make_float4(dcn[4], dcn[5], dcn[6], dcn[7]);
                }
                bf16* xrow = sq.dxb + (static_cast<long long>(xbase) + (dir == 0 ? s : L - 1 - s)) * 8 * h + dir * 4 * h + u0;
                // gate by gate: multiply, pack, store (token-order row + shared-memory operand) — short live ranges
                auto emit = [&](int g, const float (&x)[8], const uint4& co) {
                    uint4 v;
                    v.x = pack_bf16(x[0] * coef_at(co, 0), x[1] * coef_at(co, 1)); v.y = pack_bf16(x[2] * coef_at(co, 2), x[3] * coef_at(co, 3));
                    v.z = pack_bf16(x[4] * coef_at(co, 4), x[5] * coef_at(co, 5)); v.w = pack_bf16(x[6] * coef_at(co, 6), x[7] * coef_at(co, 7));
                    *reinterpret_cast<uint4*>(xrow + g * h) = v;
                    if (s >= 1) {
                        st_shared_v4(a0 + g * LB_KB_BYTES, v.x, v.y, v.z, v.w);
                        st_shared_v4(a0 + g * LB_KB_BYTES + DUP, v.x, v.y, v.z, v.w);
                    }
                };
                emit(0, dct8, st.co[LSTM_CO_BI]);
                emit(1, dct8, st.co[LSTM_CO_BF]);
                emit(2, dct8, st.co[LSTM_CO_BG]);
                emit(3, dh8, st.co[LSTM_CO_BO]);
            } else if (s >= 1) {
#pragma unroll
                for (int g = 0; g < 4; ++g) { st_shared_v4(a0 + g * LB_KB_BYTES, 0u, 0u, 0u, 0u); st_shared_v4(a0 + g * LB_KB_BYTES + DUP, 0u, 0u, 0u, 0u); }
            }
        };

        BpttStage st0, st1;
        load_stage(st0, S - 1, 0, 0);
        dcr[0] = dcr[1] = make_float4(0.f, 0.f, 0.f, 0.f);
        int it = 0;
        for (int s = S - 1; s >= 0; --s, ++it) {
            const int dbuf = (s + 1) & 1;                           // D of step s+1
            if (s < S - 1) {
                mbar_wait(&d_full[dbuf], static_cast<uint32_t>(((it - 1) >> 1) & 1), p.err_flag, 405);
                tcgen05_fence_after();
            }
            for (int c = 0; c < NC; ++c) {
                const int ab = c & 1;
                load_stage(st1, s, c, 1);
                if (s >= 1) {
                    // the MMAs that read the previous contents of this chunk buffer must have retired (one commit per use, in order)
                    const int aw = ab ? aw1 : aw0;
                    if (aw > 0) mbar_wait(&a_empty[ab], static_cast<uint32_t>((aw - 1) & 1), p.err_flag, 407);
                    if (ab) ++aw1; else ++aw0;
                }
                process(st0, s, c, 0, dbuf, s, c, 1);
                const int ns = c + 1 < NC ? s : s - 1, nc = c + 1 < NC ? c + 1 : 0;
                load_stage(st0, ns, nc, 0);
                process(st1, s, c, 1, dbuf, ns, nc, 0);
                if (s >= 1) {
                    fence_async_smem();
                    mbar_arrive(&a_ready[ab]);
                }
            }
            if (s < S - 1) {
                tcgen05_fence_before();
                mbar_arrive(&d_empty[dbuf]);
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 10) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// d question_feature = gradient of the final hidden state of both directions (module_net.py:160-163): direction 0 ends on a question's
// last token, direction 1 on its first
__global__ void add_qfeat_grad_kernel(float* __restrict__ dtok, const float* __restrict__ dqfeat, const int* __restrict__ q_off, int B, int h) {
    const long long total = static_cast<long long>(B) * 2 * h;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int b = static_cast<int>(i / (2 * h)), j = static_cast<int>(i % (2 * h));
        const int base = __ldg(q_off + b), L = __ldg(q_off + b + 1) - base;
        if (L <= 0) continue;
        dtok[(static_cast<long long>(base) + (j < h ? L - 1 : 0)) * 2 * h + j] += dqfeat[i];
    }
}

}  // namespace

bool lstm_bptt_fused_ok(int precision, int h) { return precision == STAIR_BF16 && h >= 64 && h <= 256 && (h % 64) == 0; }

// index 0 = video encoder (T steps), 1 = text encoder (ragged, L_max steps).  whhT_* = transposed W_hh copies [h][4h] (StairModel.wt).
int launch_lstm_bptt_fused(const LstmBptt& a, int B, int h, int T, int L_max, const int* q_off, int* err_flag, cudaStream_t st,
                           const int* text_order, const int* text_soff) {
    if (B <= 0) return STAIR_OK;
    if (!text_order || !text_soff) text_order = text_soff = nullptr;
    BpttParams p;
    p.err_flag = err_flag;
    // length-sorted text: the text blocks come first in the grid (longest first), the short video blocks fill in behind them (lstm_fused.cu)
    const int first = (text_order && L_max > T) ? 1 : 0;          // (more frames than words: the video blocks are the long ones and stay first)
    const void* whhT[4];
    for (int i = 0; i < 2; ++i) {
        const int e = i == 0 ? first : 1 - first;
        BpttSeq& s = p.seq[i];
        s.coef_h = reinterpret_cast<const bf16*>(a.gates[e]); s.dout = a.dout[e]; s.dxb = a.dxb[e];
        s.dc = a.dc[e]; s.q_off = e == 1 ? q_off : nullptr; s.steps = e == 0 ? T : L_max; s.B = B; s.h = h;
        s.order = e == 1 ? text_order : nullptr; s.soff = e == 1 ? text_soff : nullptr;
        whhT[2 * i] = a.whhT[2 * e]; whhT[2 * i + 1] = a.whhT[2 * e + 1];
    }
    if (a.dqfeat) {
        add_qfeat_grad_kernel<<<static_cast<int>((static_cast<long long>(B) * 2 * h + 255) / 256), 256, 0, st>>>(a.dout[1], a.dqfeat, q_off, B, h);
        STAIR_CHECK_LAUNCH();
    }
    CUtensorMap tm[4];
    for (int i = 0; i < 4; ++i) STAIR_TRY(make_tmap_bf16_2d(&tm[i], whhT[i], 4ULL * h, h, 4ULL * h, 64, 256));
    const int smem = 2 * LB_A_BYTES + LB_STAGES * LB_W_STAGE_BYTES + 256 + 1024;
    static int configured = 0;
    if (configured < smem) {
        if (cudaFuncSetAttribute(lstm_bptt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return STAIR_ERR_CUDA;
        configured = smem;
    }
    dim3 grid((B + LB_ROWS - 1) / LB_ROWS, 2, 2);
    lstm_bptt_kernel<<<grid, LB_THREADS, smem, st>>>(tm[0], tm[1], tm[2], tm[3], p);
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

}  // namespace stair
