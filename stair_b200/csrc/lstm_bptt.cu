// Persistent fused BPTT of the BiLSTM encoders (bf16 training path): the whole backward time loop of nn.LSTM
// (video_nmn/module_net.py:39-47, 147-163; autograd of train_module.py:408) in ONE launch for both encoders and both directions.
// Replaces, per step and direction, a cell kernel + a recurrent GEMM launch (2 x (T + L_max) launches of 15-30 us each on the
// critical path of the training step).
//
// One CTA owns 64 questions of one (encoder, direction) for all steps, walking s = S-1 .. 0:
//   dh_s      = d(encoder output)[token row] (+ dh_rec from step s+1: TMEM) (+ d question_feature at a question's last step)
//   gate math = the derivative of the cell (lstm.cu / train_kernels.cu lstm_cell_bwd_kernel) from the forward's blocked history
//               (gates i,f,g,o post-activation and cell states, lstm_hist_*_off) and the running dc (fp32, private, coalesced)
//   dG_s      -> bf16 rows [rows][8h] in token order (the operand of dW_ih / dW_hh / bias, executor_bwd.cu) and, as the A operand of
//               dh_rec_{s-1} = dG_s . W_hh, into shared memory: chunk c = the four gates of hidden units 64c .. 64c+63 = four k-blocks
//               (one per gate) of a [128 x 256] K-major SWIZZLE_128B tile, double-buffered by chunk parity
//   tcgen05.mma accumulates D[128 x h] over the 4 (or h/64) chunks while later chunks are still being produced; B tiles are rows of the
//   transposed W_hh copy (StairModel.wt, [h][4h]: box {64 k, 256 n} at k = gate*h + 64c) streamed through a TMA ring; D is
//   double-buffered in TMEM by step parity so the MMA of step s-1 never waits for the last reads of D_s.
//
// Warps (512 threads, 128 registers): 2 TMA producer, 3 MMA issuer, 6 TMEM allocator,
// epilogue = the 8 warps whose TMEM lane quarter (warp % 4) is 0 or 1 (rows 0-63), 4 column groups of 16 units per chunk.
// (An L2-prefetch warp running 1-3 steps ahead was measured and removed: 795 -> 875-895 us and 1.53 -> 2.8-3.1 GB of DRAM reads at
// B = 4096 — the steps in flight of 148 CTAs do not stay in L2.)
#include "nmn_kernels.cuh"
#include "tc_ptx.cuh"
#include "train_kernels.cuh"

namespace stair {

namespace {

constexpr int LB_ROWS = 64;                        // questions per CTA (MMA M = 128, lanes 64-127 unused)
constexpr int LB_KB_BYTES = 128 * 64 * 2;          // one 64-wide k-block of the A tile (128 rows): 16 KiB
constexpr int LB_A_BYTES = 4 * LB_KB_BYTES;        // one chunk: 4 gates x 64 units
constexpr int LB_W_STAGE_BYTES = 256 * 64 * 2;     // one B tile (256 n x 64 k): 32 KiB
constexpr int LB_STAGES = 3;
constexpr int LB_THREADS = 512;
constexpr int LB_EPI = 256;                        // epilogue threads
constexpr uint32_t LB_IDESC = make_idesc_bf16(128, 256);

struct BpttSeq {
    const float* gates_h;   // blocked history (lstm_hist_gate_off)
    const float* c_h;       // blocked history (lstm_hist_c_off)
    const float* dout;      // [rows][2h] gradient of the encoder output
    const float* dqfeat;    // text: [B][2h] gradient of question_feature (added at a question's last step); video: null
    bf16* dxb;              // [rows][8h] gate pre-activation gradients, token order
    float* dc;              // running dc scratch [2][nblk][h][64]
    const int* q_off;       // text: [B+1]; video: null
    int steps, B, h;
};
struct BpttParams { BpttSeq seq[2]; int* err_flag; };

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

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
    const bool is_epi = (warp & 3) < 2;
    const int quarter = warp & 3, cgrp = warp >> 2;          // column group 0..3: units 16*cgrp .. +15 of a chunk

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
    if (warp == 6) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_ptr_smem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 2 * LB_A_BYTES / 16; i += LB_THREADS) reinterpret_cast<uint4*>(sA)[i] = make_uint4(0, 0, 0, 0);   // rows 64-127 stay 0
    __syncthreads();
    if (ragged && threadIdx.x < LB_ROWS) {
        const int r = row0 + threadIdx.x;
        if (r < sq.B) atomicMax(s_steps, __ldg(sq.q_off + r + 1) - __ldg(sq.q_off + r));
    }
    fence_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const int S = *s_steps;

    if (warp == 2) {
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
    } else if (warp == 3) {
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
        // ===================== cell backward: thread = question row, 16 units per chunk =============================================
        const int row = quarter * 32 + lane;
        const int grow = row0 + row;
        const bool valid = grow < sq.B;
        int base = 0, L = sq.steps;
        if (ragged) { base = valid ? __ldg(sq.q_off + grow) : 0; L = valid ? __ldg(sq.q_off + grow + 1) - base : 0; }
        else base = grow * sq.steps;
        if (!valid) L = 0;
        const int nblk = (sq.B + LB_ROWS - 1) / LB_ROWS;
        float* dcblk = sq.dc + (static_cast<long long>(dir) * nblk + blockIdx.x) * (static_cast<long long>(h) * LB_ROWS) + row * 4;   // [unit/4][row][4]
        const long long RB = (sq.B + 127) / 128 * 4;
        const long long hist_rb = (static_cast<long long>(dir) * RB + (grow >> 5)) * (h >> 3);
        const long long hist_step = 2 * RB * (h >> 3);
        const uint32_t sA0 = smem_u32(sA);
        const uint32_t rowoff = static_cast<uint32_t>(row) * 128u;
        const uint32_t sw = static_cast<uint32_t>(row & 7);
        int aw[2] = {0, 0};                                        // writes into each chunk buffer so far
        int it = 0;
        for (int s = S - 1; s >= 0; --s, ++it) {
            const bool active = s < L;
            const bool has_next = s + 1 < L;                       // this row was active at step s+1: dh_rec / dc carry over
            const bool final_step = s == L - 1;
            const long long tokrow = static_cast<long long>(base) + (dir == 0 ? s : L - 1 - s);
            const float* drow = sq.dout + tokrow * 2 * h + dir * h;
            bf16* xrow = sq.dxb + tokrow * 8 * h + dir * 4 * h;
            const int dbuf = (s + 1) & 1;                           // D of step s+1
            if (s < S - 1) {
                mbar_wait(&d_full[dbuf], static_cast<uint32_t>(((it - 1) >> 1) & 1), p.err_flag, 405);
                tcgen05_fence_after();
            }
            for (int c = 0; c < NC; ++c) {
                const int ab = c & 1;
                if (s >= 1) {
                    // the MMAs that read the previous contents of this chunk buffer must have retired (one commit per use, in order)
                    if (aw[ab] > 0) mbar_wait(&a_empty[ab], static_cast<uint32_t>((aw[ab] - 1) & 1), p.err_flag, 407);
                    ++aw[ab];
                }
#pragma unroll
                for (int sb = 0; sb < 2; ++sb) {
                    const int u0 = c * 64 + cgrp * 16 + sb * 8;
                    uint32_t dr[8];
                    if (s < S - 1) {                                // .sync.aligned: the whole warp, converged
                        tmem_ld8(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(dbuf * 256 + u0), dr);
                        tmem_ld_wait();
                    }
                    float dpi[8], dpf[8], dpg[8], dpo[8];
                    if (active) {
                        const float* gp = sq.gates_h + (s * hist_step + hist_rb + (u0 >> 3)) * 1024 + lane * 8;
                        const float* cc = sq.c_h + (s * hist_step + hist_rb + (u0 >> 3)) * 256 + lane * 8;
                        float ig[8], fg[8], gg[8], og[8], ccur[8], cprev[8], dh[8], dcv[8];
                        *reinterpret_cast<float4*>(ig) = *reinterpret_cast<const float4*>(gp);
                        *reinterpret_cast<float4*>(ig + 4) = *reinterpret_cast<const float4*>(gp + 4);
                        *reinterpret_cast<float4*>(fg) = *reinterpret_cast<const float4*>(gp + 256);
                        *reinterpret_cast<float4*>(fg + 4) = *reinterpret_cast<const float4*>(gp + 260);
                        *reinterpret_cast<float4*>(gg) = *reinterpret_cast<const float4*>(gp + 512);
                        *reinterpret_cast<float4*>(gg + 4) = *reinterpret_cast<const float4*>(gp + 516);
                        *reinterpret_cast<float4*>(og) = *reinterpret_cast<const float4*>(gp + 768);
                        *reinterpret_cast<float4*>(og + 4) = *reinterpret_cast<const float4*>(gp + 772);
                        *reinterpret_cast<float4*>(ccur) = *reinterpret_cast<const float4*>(cc);
                        *reinterpret_cast<float4*>(ccur + 4) = *reinterpret_cast<const float4*>(cc + 4);
                        if (s > 0) {
                            const float* cpp = cc - hist_step * 256;
                            *reinterpret_cast<float4*>(cprev) = *reinterpret_cast<const float4*>(cpp);
                            *reinterpret_cast<float4*>(cprev + 4) = *reinterpret_cast<const float4*>(cpp + 4);
                        }
                        *reinterpret_cast<float4*>(dh) = *reinterpret_cast<const float4*>(drow + u0);
                        *reinterpret_cast<float4*>(dh + 4) = *reinterpret_cast<const float4*>(drow + u0 + 4);
                        if (has_next) {
                            *reinterpret_cast<float4*>(dcv) = *reinterpret_cast<const float4*>(dcblk + (u0 / 4) * (LB_ROWS * 4));
                            *reinterpret_cast<float4*>(dcv + 4) = *reinterpret_cast<const float4*>(dcblk + (u0 / 4 + 1) * (LB_ROWS * 4));
                        }
                        if (final_step && sq.dqfeat) {
                            const float* qr = sq.dqfeat + static_cast<long long>(grow) * 2 * h + dir * h + u0;
#pragma unroll
                            for (int j = 0; j < 8; ++j) dh[j] += qr[j];
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float dhj = dh[j];
                            if (has_next) dhj += __uint_as_float(dr[j]);
                            const float tc = tanhf(ccur[j]);
                            const float dct = (has_next ? dcv[j] : 0.f) + dhj * og[j] * (1.f - tc * tc);
                            const float cp = s > 0 ? cprev[j] : 0.f;
                            dpi[j] = dct * gg[j] * ig[j] * (1.f - ig[j]);
                            dpf[j] = dct * cp * fg[j] * (1.f - fg[j]);
                            dpg[j] = dct * ig[j] * (1.f - gg[j] * gg[j]);
                            dpo[j] = dhj * tc * og[j] * (1.f - og[j]);
                            dcv[j] = dct * fg[j];
                        }
                        *reinterpret_cast<float4*>(dcblk + (u0 / 4) * (LB_ROWS * 4)) = make_float4(dcv[0], dcv[1], dcv[2], dcv[3]);
                        *reinterpret_cast<float4*>(dcblk + (u0 / 4 + 1) * (LB_ROWS * 4)) = make_float4(dcv[4], dcv[5], dcv[6], dcv[7]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) { dpi[j] = 0.f; dpf[j] = 0.f; dpg[j] = 0.f; dpo[j] = 0.f; }
                    }
                    uint4 q[4];
                    q[0] = make_uint4(pack_bf16(dpi[0], dpi[1]), pack_bf16(dpi[2], dpi[3]), pack_bf16(dpi[4], dpi[5]), pack_bf16(dpi[6], dpi[7]));
                    q[1] = make_uint4(pack_bf16(dpf[0], dpf[1]), pack_bf16(dpf[2], dpf[3]), pack_bf16(dpf[4], dpf[5]), pack_bf16(dpf[6], dpf[7]));
                    q[2] = make_uint4(pack_bf16(dpg[0], dpg[1]), pack_bf16(dpg[2], dpg[3]), pack_bf16(dpg[4], dpg[5]), pack_bf16(dpg[6], dpg[7]));
                    q[3] = make_uint4(pack_bf16(dpo[0], dpo[1]), pack_bf16(dpo[2], dpo[3]), pack_bf16(dpo[4], dpo[5]), pack_bf16(dpo[6], dpo[7]));
                    if (active) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) *reinterpret_cast<uint4*>(xrow + g * h + u0) = q[g];
                    }
                    if (s >= 1) {                                   // A operand of dh_rec_{s-1}: k-block = gate, 16-byte chunk = (u0 % 64) / 8
                        const uint32_t ch = static_cast<uint32_t>((u0 & 63) >> 3);
                        const uint32_t a0 = sA0 + static_cast<uint32_t>(ab * LB_A_BYTES) + rowoff + ((ch ^ sw) << 4);
#pragma unroll
                        for (int g = 0; g < 4; ++g) st_shared_v4(a0 + g * LB_KB_BYTES, q[g].x, q[g].y, q[g].z, q[g].w);
                    }
                }
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
    if (warp == 6) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

}  // namespace

bool lstm_bptt_fused_ok(int precision, int h) { return precision == STAIR_BF16 && h >= 64 && h <= 256 && (h % 64) == 0; }

// index 0 = video encoder (T steps), 1 = text encoder (ragged, L_max steps).  whhT_* = transposed W_hh copies [h][4h] (StairModel.wt).
int launch_lstm_bptt_fused(const LstmBptt& a, int B, int h, int T, int L_max, const int* q_off, int* err_flag, cudaStream_t st) {
    if (B <= 0) return STAIR_OK;
    BpttParams p;
    p.err_flag = err_flag;
    for (int e = 0; e < 2; ++e) {
        BpttSeq& s = p.seq[e];
        s.gates_h = a.gates[e]; s.c_h = a.c[e]; s.dout = a.dout[e]; s.dqfeat = e == 1 ? a.dqfeat : nullptr; s.dxb = a.dxb[e];
        s.dc = a.dc[e]; s.q_off = e == 1 ? q_off : nullptr; s.steps = e == 0 ? T : L_max; s.B = B; s.h = h;
    }
    CUtensorMap tm[4];
    for (int i = 0; i < 4; ++i) STAIR_TRY(make_tmap_bf16_2d(&tm[i], a.whhT[i], 4ULL * h, h, 4ULL * h, 64, 256));
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
