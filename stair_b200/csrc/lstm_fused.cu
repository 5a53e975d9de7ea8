// Persistent fused BiLSTM recurrence (bf16 path): the whole time loop of nn.LSTM (video_nmn/module_net.py:39-47,
// 147-163) in ONE launch for both encoders and both directions.
//
// One CTA owns 64 (product) or 128 questions of one (encoder, direction) for all T (or L) steps:
//   h_{t-1} lives in shared memory as the bf16 A operand (128 x h, K-major, SWIZZLE_128B, double-buffered),
//   W_hh (gate-interleaved so that a 256-column chunk = 64 hidden units x {i,f,g,o}) streams from L2 through a TMA ring,
//   tcgen05.mma accumulates the 128 x 256 gate pre-activations of a chunk in TMEM (two chunks in flight),
//   8 epilogue warps add the precomputed input projection (xproj), apply the cell non-linearities, update c (fp32, global,
//   L2-resident), write h_t to the encoder output and straight back into the other shared-memory h buffer.
// No per-step launches, no [B,4h] round trips: per step only xproj rows and h rows touch HBM.
//
// 128-row form: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4.. cell epilogue; 64-row form (product): see the
// kernel's template comment.
#include "nmn_kernels.cuh"
#include "tc_ptx.cuh"
#include "train_kernels.cuh"
#include <cstdlib>

namespace stair {

namespace {

constexpr int LF_ROWS = 128;
constexpr int LF_KB_BYTES = LF_ROWS * 64 * 2;      // one 64-wide k-block of h: 16 KiB
constexpr int LF_W_STAGE_BYTES = 256 * 64 * 2;     // one W tile (256 gate columns x 64 k): 32 KiB
constexpr int LF_STAGES = 3;
constexpr uint32_t LF_IDESC = make_idesc_bf16(128, 256);

struct LstmSeq {
    const bf16* xproj;      // [rows, 8h]: W_ih x + b for both directions (fwd gates | reverse gates)
    float* c;               // [2][B][h] cell state scratch (fp32)
    bf16* out;              // video: [B*T, 2h] (VID arena slots 0..B-1) ; text: token_feature [n_tok, 2h]
    bf16* final_h;          // text: question_feature [B, 2h] (h_n of both directions) ; video: null
    const int* q_off;       // text: [B+1] token offsets (ragged) ; video: null
    const int* order;       // text, inference (optional): [B] question ids in descending length — grid row r is question order[r], so a CTA's
    const int* soff;        //   questions have similar lengths and it stops at its own longest.  soff [B+1]: token offsets in THAT order: xproj is
                            //   staged physically sorted (rows soff[r] .. of position r; text_sort_kernel), so a block's input rows stay one
                            //   contiguous range (an index-only sort scatters them over the whole projection and measured slower).  Outputs
                            //   go to the questions' own rows (q_off).  null = batch order
    int steps;              // video: T ; text: L_max
    int B, h;
    // training (HIST): per-step history for BPTT = the six bf16 cell-derivative coefficients of train_kernels.cuh (lstm_hist_coef_off:
    // [step][dir][32-row block][8-unit block][coefficient][row][8 units]) so that a warp's stores are contiguous (a row-major history =
    // 32 scattered 16-byte pieces per store instruction, measured 3x slower).  The running cell state stays in the private fp32 scratch
    // `c` (the backward never reads c).  h goes to hs_h [rows][2h] bf16 in token order, shifted by one step (row of token r holds the
    // state BEFORE r was consumed): the second operand of the dW_hh contraction.
    bf16* coef_h; bf16* hs_h; long long hs_dir;
};

struct LstmFusedParams {
    LstmSeq seq[2];
    int* err_flag;
    volatile unsigned int* dbg;   // debug progress words (pinned host memory), null in production
    int prefetch;                 // 1 = the L2 prefetch warp runs two steps ahead of the cell epilogue (STAIR_LSTM_PF=1; off by default: since
                                  // the chunk's operands are requested before the MMA wait it only adds DRAM reads, 405 -> 495 MB, 400 -> 403 us)
};
#define LF_DBG(slot, val) do { if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) { p.dbg[slot] = (val); __threadfence_system(); } } while (0)

__device__ __forceinline__ float fast_tanh(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// sigmoid(x) = 0.5 + 0.5 tanh(x / 2): one MUFU op (the cell epilogue is MUFU-bound: 5 transcendentals per hidden unit)
__device__ __forceinline__ float fast_sigmoid(float x) { return fmaf(0.5f, fast_tanh(0.5f * x), 0.5f); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr) : "memory");
}
// 32-byte global accesses (sm_100 LDG/STG .256): one request per gate for a thread's 16 units instead of two — every lane of the cell
// epilogue addresses its own question's row, so each access is its own LSU wavefront, and the wavefronts are the kernel's most loaded unit
__device__ __forceinline__ void ldg256_nc(const void* p, uint4& a, uint4& b) {
    asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint4& a, const uint4& b) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :: "l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
    uint4 r;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr) : "memory");
    return r;
}
__device__ __forceinline__ void unpack8(const uint4& r, float (&f)[8]) {
    const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(hh[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}

// CG = column groups of a chunk's 64 hidden units (each epilogue thread takes 64 / CG of them).
// VROWS = questions per CTA.  128: all four TMEM lane quarters carry distinct rows, warps 0-3 are the role warps, 4.. the epilogue
// (CG warps per quarter, threads = 128 + 128 * CG).  64 (product, CG = 4; CG = 8 kept for comparison): the MMA is still M = 128 and
// rows 64-127 of the h operand hold a COPY of rows 0-63, so TMEM lanes 64-127 repeat lanes 0-63 and the 2 CG epilogue warps (warps
// 0 .. 2 CG - 1) sit on all four lane quarters = all four SM sub-partitions; quarters q and q + 2 share rows and split the column
// groups.  Half the per-step epilogue latency per question block of the 128-row form (the recurrence is a chain of L sequential
// steps, so that latency is the kernel's critical path) and twice as many CTAs to fill the SMs the video blocks free after T steps.
// Role warps follow the epilogue warps: TMA, MMA, TMEM allocator, (optional) L2 prefetch; 12 warps -> up to 168 registers per thread.
template <int CG, bool HIST, int VROWS>
__global__ void __launch_bounds__(VROWS == 64 ? 64 * CG + 128 : 128 + 128 * CG, 1)
lstm_fused_kernel(const __grid_constant__ CUtensorMap tmW0, const __grid_constant__ CUtensorMap tmW1,
                  const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmW3, const LstmFusedParams p) {
    const LstmSeq sq = blockIdx.z == 0 ? p.seq[0] : p.seq[1];      // by value: a runtime index into param space forces a local copy
    const int dir = blockIdx.y;
    const int wsel = blockIdx.z * 2 + dir;          // which W_hh map (never form a runtime-selected pointer to a param-space map)
    static_assert(VROWS == 128 || (VROWS == 64 && (CG == 4 || CG == 8)), "64-row blocks use 4 or 8 column groups: half of them per pair of TMEM quarters");
    const int row0 = blockIdx.x * VROWS;
    if (row0 >= sq.B) return;
    const int h = sq.h, NC = h / 64;              // chunks of 64 hidden units == k-blocks of h
    const bool ragged = sq.q_off != nullptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int LF_THREADS = VROWS == 64 ? 64 * CG + 128 : 128 + 128 * CG;
    constexpr int SBN = 8 / CG;                   // 8-unit sub-blocks per thread per chunk
    constexpr int EPI_THREADS = VROWS * CG;       // epilogue threads (arrival count of the tmem_empty / h_ready barriers)
    // role warps and the (TMEM quarter, column group) of an epilogue warp
    // 64-row blocks: the MMA is still M = 128, and rows 64-127 of the h operand hold a COPY of rows 0-63, so TMEM lanes 64-127 carry the
    // same gate pre-activations as lanes 0-63.  That lets the 8 epilogue warps sit on all four TMEM lane quarters = all four SM
    // sub-partitions (warp % 4): quarters q and q + 2 share the rows (q & 1) * 32 .. + 31 and split the column groups.  With the epilogue
    // on quarters 0 and 1 only (zero rows 64-127), two schedulers and half of the MUFU units did all the cell math.
    constexpr int W_EPI64 = 2 * CG;               // 64-row blocks: epilogue warps 0 .. 2 CG - 1, then the four role warps
    constexpr int W_TMA = VROWS == 64 ? W_EPI64 : 0, W_MMA = VROWS == 64 ? W_EPI64 + 1 : 1, W_ALLOC = VROWS == 64 ? W_EPI64 + 2 : 2,
                  W_PREF = VROWS == 64 ? W_EPI64 + 3 : 3;
    const bool is_epi = VROWS == 64 ? (warp < W_EPI64) : (warp >= 4);
    const int quarter = warp & 3;
    const int halfsel = VROWS == 64 ? ((warp >> 2) * 2 + (quarter >> 1)) : ((warp - 4) >> 2);      // halfsel = column group 0..CG-1

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    const int hbuf_bytes = NC * LF_KB_BYTES;
    uint8_t* sH = smem;                                            // [2][NC][128 x 64] bf16
    uint8_t* sW = smem + 2 * hbuf_bytes;                           // [LF_STAGES][256 x 64] bf16
    uint64_t* w_full = reinterpret_cast<uint64_t*>(sW + LF_STAGES * LF_W_STAGE_BYTES);
    uint64_t* w_empty = w_full + LF_STAGES;
    uint64_t* tmem_full = w_empty + LF_STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* h_ready = tmem_empty + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(h_ready + 4);     // h_ready[kb]: k-block kb (= chunk kb) of h_s is in smem
    int* s_steps = reinterpret_cast<int*>(tmem_ptr_smem + 1);

    if (threadIdx.x == 0) {
        for (int s = 0; s < LF_STAGES; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], EPI_THREADS); }
        for (int a = 0; a < 4; ++a) mbar_init(&h_ready[a], EPI_THREADS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        *s_steps = ragged ? 0 : sq.steps;
    }
    if (warp == W_ALLOC) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_ptr_smem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 2 * hbuf_bytes / 16; i += LF_THREADS) reinterpret_cast<uint4*>(sH)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    if (ragged && threadIdx.x < VROWS) {                            // this CTA only runs as many steps as its longest question
        const int r = row0 + threadIdx.x;
        const int* off = sq.soff ? sq.soff : sq.q_off;
        if (r < sq.B) atomicMax(s_steps, __ldg(off + r + 1) - __ldg(off + r));
    }
    fence_async_smem();                                            // zero-filled h buffers visible to the tensor core (async proxy)
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const int S = *s_steps;

    if (warp == W_TMA) {
        if (lane == 0) {
            // ===================== TMA producer: W_hh tiles, (chunk, k-block) order, every step after the first ==========
            int stage = 0; uint32_t phase = 0;
            for (int s = 1; s < S; ++s)
                for (int c = 0; c < NC; ++c)
                    for (int kb = 0; kb < NC; ++kb) {
                        LF_DBG(0, 0x10000u + s * 256 + c * 16 + kb);
                        mbar_wait(&w_empty[stage], phase ^ 1, p.err_flag, 201);
                        mbar_arrive_expect_tx(&w_full[stage], LF_W_STAGE_BYTES);
                        uint8_t* dst = sW + stage * LF_W_STAGE_BYTES;
                        if (wsel == 0) tma_load_2d(dst, &tmW0, &w_full[stage], kb * 64, c * 256);
                        else if (wsel == 1) tma_load_2d(dst, &tmW1, &w_full[stage], kb * 64, c * 256);
                        else if (wsel == 2) tma_load_2d(dst, &tmW2, &w_full[stage], kb * 64, c * 256);
                        else tma_load_2d(dst, &tmW3, &w_full[stage], kb * 64, c * 256);
                        if (++stage == LF_STAGES) { stage = 0; phase ^= 1; }
                    }
        }
    } else if (warp == W_MMA) {
        if (lane == 0) {
            // ===================== MMA issuer: gates[128, 256c..] = h_{s-1}[128, h] . W_hh[chunk]^T ============================
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int s = 1; s < S; ++s) {
                LF_DBG(1, 0x20000u + s * 256);
                const uint32_t hb = smem_u32(sH + (s & 1) * hbuf_bytes);
                for (int c = 0; c < NC; ++c) {
                    mbar_wait(&tmem_empty[acc], acc_phase ^ 1, p.err_flag, 203);
                    tcgen05_fence_after();
                    const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * 256);
                    for (int kb = 0; kb < NC; ++kb) {
                        LF_DBG(1, 0x24000u + s * 256 + c * 16 + kb);
                        // k-block kb of h_{s-1} is complete as soon as the cell epilogue of chunk kb of step s-1 is: the first
                        // chunk of a step starts accumulating while the last chunks of the previous step are still in their epilogue
                        if (c == 0) mbar_wait(&h_ready[kb], static_cast<uint32_t>((s - 1) & 1), p.err_flag, 202);
                        mbar_wait(&w_full[stage], phase, p.err_flag, 204);
                        tcgen05_fence_after();
                        const uint64_t adesc = make_umma_desc_kmajor_sw128(hb + kb * LF_KB_BYTES);
                        const uint64_t bdesc = make_umma_desc_kmajor_sw128(smem_u32(sW + stage * LF_W_STAGE_BYTES));
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, LF_IDESC, (kb | k) != 0 ? 1u : 0u);
                        umma_commit(&w_empty[stage]);
                        if (++stage == LF_STAGES) { stage = 0; phase ^= 1; }
                    }
                    umma_commit(&tmem_full[acc]);
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
    } else if (warp == W_PREF) {
        // ===================== L2 prefetcher: the input-projection rows of step s+2 (they do not depend on the recurrence) ========
        // xproj (134 + 268 MB at B=4096) does not fit in L2; without this every epilogue load is an HBM-latency miss.
        for (int s = 0; s < S && p.prefetch; ++s) {
            // pace: rows of step s are requested once step s-3 is complete (two steps ahead of the cell epilogue).  A parity wait on
            // a phase that is already two behind simply returns one phase later; a later phase of that parity always exists here.
            if (s >= 3) mbar_wait(&h_ready[NC - 1], static_cast<uint32_t>((s - 3) & 1), p.err_flag, 206);
            for (int r = lane; r < VROWS; r += 32) {
                const int grow = row0 + r;
                if (grow >= sq.B) continue;
                int base, L = sq.steps;
                if (ragged) { const int* off = sq.soff ? sq.soff : sq.q_off; base = __ldg(off + grow); L = __ldg(off + grow + 1) - base; }
                else base = grow * sq.steps;
                if (s >= L) continue;
                const bf16* xrow = sq.xproj + (static_cast<long long>(base) + (dir == 0 ? s : L - 1 - s)) * 8 * h + dir * 4 * h;
                for (int b = 0; b < 4 * h * 2; b += 128) prefetch_l2(reinterpret_cast<const char*>(xrow) + b);
            }
        }
    } else if (is_epi) {
        // ===================== cell epilogue: CG warps per TMEM lane quarter, each takes 64 / CG of a chunk's 64 units =============
        const int row = (VROWS == 64 ? (quarter & 1) : quarter) * 32 + lane;
        const int pos = row0 + row;                                 // grid row; the question it carries is order[pos] when the text is length-sorted
        const bool valid = pos < sq.B;
        const int grow = (valid && sq.order) ? __ldg(sq.order + pos) : pos;
        int base = 0, xbase = 0, L = sq.steps;                      // first output row / first input-projection row of the question
        if (ragged) {
            base = valid ? __ldg(sq.q_off + grow) : 0; L = valid ? __ldg(sq.q_off + grow + 1) - base : 0;
            xbase = (valid && sq.soff) ? __ldg(sq.soff + pos) : base;
        } else xbase = base = grow * sq.steps;
        constexpr uint32_t DUP = VROWS == 64 ? 64u * 128u : 0u;      // byte offset of the copy of a row in the h operand (row + 64)
        // cell state scratch, private to this CTA, laid out [unit/4][row][4] so that a warp's float4 accesses are contiguous
        const int nblk = (sq.B + VROWS - 1) / VROWS;
        float* cblk = sq.c + (static_cast<long long>(dir) * nblk + blockIdx.x) * (static_cast<long long>(h) * VROWS) + row * 4;
        const long long RB = (sq.B + 127) / 128 * 4;                                      // 32-row blocks per (step, direction)
        const long long hist_rb = (static_cast<long long>(dir) * RB + (pos >> 5)) * (h >> 3);       // + step * 2 * RB * (h/8); then + unit/8 (by grid row)
        const long long hist_step = 2 * RB * (h >> 3);
        auto c_ptr = [&](int, int unit, int q) -> float* { return cblk + (unit / 4 + q) * (VROWS * 4); };     // q-th float4 of 8 units
        const uint32_t sH0 = smem_u32(sH);
        const uint32_t rowoff = static_cast<uint32_t>(row) * 128u;
        const uint32_t sw = static_cast<uint32_t>(row & 7);
        int acc = 0; uint32_t acc_phase = 0;
        constexpr bool GS_CHUNK_C = !HIST && CG == 8;                // the 16-warp comparison form loads the cell state per chunk
        float4 cnext[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};     // cell state of the next sub-block's 8 units
        // (Requesting the input-projection pieces one sub-block ahead as well — two alternating register sets — measured SLOWER, 1.26 ->
        // 1.345 ms per forward: a thread's two sub-blocks share 32-byte sectors and separate requests double the L2 traffic.  They are
        // requested per chunk, as one 32-byte load per gate: profiles/r3_lstm_sorted_ncu_summary.txt.)
        constexpr bool WIDE = SBN % 2 == 0;       // 32-byte input-projection loads / output stores for pairs of sub-blocks
        for (int s = 0; s < S; ++s) {
            const bool active = valid && s < L;
            const int tstep = dir == 0 ? s : L - 1 - s;
            const long long tokrow = static_cast<long long>(base) + tstep;
            const bf16* xrow = sq.xproj + (static_cast<long long>(xbase) + tstep) * 8 * h + dir * 4 * h;
            bf16* orow = sq.out + tokrow * 2 * h + dir * h;
            const bool last = ragged && s == L - 1;
            const uint32_t h_src = sH0 + static_cast<uint32_t>((s & 1) * hbuf_bytes);
            const uint32_t h_dst = sH0 + static_cast<uint32_t>(((s + 1) & 1) * hbuf_bytes);
            for (int c = 0; c < NC; ++c) {
                // all global operands of this chunk (input projection, previous cell state) are requested before waiting for the
                // tensor core, so their latency overlaps the MMA of this chunk instead of serialising 4x per chunk
                uint4 xq[SBN][4];
                if (active) {
                    if constexpr (WIDE) {
#pragma unroll
                        for (int sb = 0; sb < SBN; sb += 2) {          // a pair of sub-blocks = 16 units = one 32-byte sector per gate
                            const int u0 = c * 64 + halfsel * (SBN * 8) + sb * 8;
#pragma unroll
                            for (int g = 0; g < 4; ++g) ldg256_nc(xrow + g * h + u0, xq[sb][g], xq[sb + 1][g]);
                        }
                    } else {
#pragma unroll
                        for (int sb = 0; sb < SBN; ++sb) {
                            const int u0 = c * 64 + halfsel * (SBN * 8) + sb * 8;
#pragma unroll
                            for (int g = 0; g < 4; ++g) xq[sb][g] = __ldg(reinterpret_cast<const uint4*>(xrow + g * h + u0));      // L1-allocating: the next sub-block reads the other half of the sector (no-allocate measured 1.6x slower)
                        }
                    }
                    if (GS_CHUNK_C && s > 0) {
#pragma unroll
                        for (int q = 0; q < 2; ++q) cnext[q] = *reinterpret_cast<const float4*>(c_ptr(s - 1, c * 64 + halfsel * (SBN * 8), q));
                    }
                }
                if (s > 0) {
                    mbar_wait(&tmem_full[acc], acc_phase, p.err_flag, 205);
                    tcgen05_fence_after();
                }
                uint4 o_even = make_uint4(0, 0, 0, 0);                  // packed h of the even sub-block of a pair (WIDE)
#pragma unroll
                for (int sb = 0; sb < SBN; ++sb) {
                    const int u0 = c * 64 + halfsel * (SBN * 8) + sb * 8;    // 8 hidden units u0 .. u0+7 (one 16-byte chunk of the h row)
                    // GS (16 epilogue warps, 96 registers): the gates are taken from TMEM two at a time -- i and g first (their product is
                    // all the cell update needs of them), then f and o into the same registers -- so that no 32-value gate set is ever live
                    constexpr bool GS = !HIST && CG == 8;
                    uint32_t gi[8], gf[GS ? 1 : 8], gg[8], go[GS ? 1 : 8];
                    const uint32_t t = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                                       static_cast<uint32_t>(acc * 256 + halfsel * (SBN * 8) + sb * 8);
                    if (s > 0) {                                      // tcgen05.ld / wait::ld are .sync.aligned: the whole warp, converged
                        if constexpr (GS) { tmem_ld8(t, gi); tmem_ld8(t + 128, gg); }
                        else { tmem_ld8(t, gi); tmem_ld8(t + 64, gf); tmem_ld8(t + 128, gg); tmem_ld8(t + 192, go); }
                        tmem_ld_wait();
                    }
                    const uint32_t ch = static_cast<uint32_t>(halfsel * SBN + sb);
                    const uint32_t a0 = static_cast<uint32_t>(c) * LF_KB_BYTES + rowoff + ((ch ^ sw) << 4);
                    if constexpr (GS) {
                        float ig_g[8];
                        if (active) {
                            float fi[8], fg[8];
                            unpack8(xq[sb][0], fi); unpack8(xq[sb][2], fg);
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                float pi = fi[j], pg = fg[j];
                                if (s > 0) { pi += __uint_as_float(gi[j]); pg += __uint_as_float(gg[j]); }
                                ig_g[j] = fast_sigmoid(pi) * fast_tanh(pg);
                            }
                        }
                        if (s > 0) { tmem_ld8(t + 64, gi); tmem_ld8(t + 192, gg); tmem_ld_wait(); }      // f -> gi, o -> gg (converged again)
                        if (active) {
                            float ff[8], fo[8], cn[8];
                            unpack8(xq[sb][1], ff); unpack8(xq[sb][3], fo);
                            uint32_t hp[4];
#pragma unroll
                            for (int jp = 0; jp < 4; ++jp) {
                                float hv[2];
#pragma unroll
                                for (int e = 0; e < 2; ++e) {
                                    const int j = 2 * jp + e;
                                    float pf = ff[j], po = fo[j], cp = 0.0f;
                                    if (s > 0) {
                                        pf += __uint_as_float(gi[j]); po += __uint_as_float(gg[j]);
                                        cp = j < 4 ? (&cnext[0].x)[j] : (&cnext[1].x)[j - 4];
                                    }
                                    const float cc = fast_sigmoid(pf) * cp + ig_g[j];
                                    cn[j] = cc;
                                    hv[e] = fast_sigmoid(po) * fast_tanh(cc);
                                }
                                hp[jp] = pack_bf16(hv[0], hv[1]);
                            }
#pragma unroll
                            for (int q = 0; q < 2; ++q)
                                *reinterpret_cast<float4*>(c_ptr(s, u0, q)) = make_float4(cn[4 * q], cn[4 * q + 1], cn[4 * q + 2], cn[4 * q + 3]);
                            *reinterpret_cast<uint4*>(orow + u0) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
                            if (last) *reinterpret_cast<uint4*>(sq.final_h + static_cast<long long>(grow) * 2 * h + dir * h + u0) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
                            st_shared_v4(h_dst + a0, hp[0], hp[1], hp[2], hp[3]);
                            if (DUP) st_shared_v4(h_dst + a0 + DUP, hp[0], hp[1], hp[2], hp[3]);
                        } else {
                            const uint4 p0 = ld_shared_v4(h_src + a0);
                            st_shared_v4(h_dst + a0, p0.x, p0.y, p0.z, p0.w);
                            if (DUP) st_shared_v4(h_dst + a0 + DUP, p0.x, p0.y, p0.z, p0.w);
                        }
                    } else if (active) {
                        float fi[8], ff[8], fg[8], fo[8], cn[8];
                        unpack8(xq[sb][0], fi); unpack8(xq[sb][1], ff); unpack8(xq[sb][2], fg); unpack8(xq[sb][3], fo);
                        const float cprev[8] = {cnext[0].x, cnext[0].y, cnext[0].z, cnext[0].w, cnext[1].x, cnext[1].y, cnext[1].z, cnext[1].w};
                        // rolling prefetch of the cell state, one sub-block ahead ACROSS chunk and step boundaries: the state of the next
                        // sub-block in processing order was written by this thread at least a chunk ago (the first sub-block of the next step:
                        // in this step's first chunk).  Requested at the start of a chunk (right before the wait for the tensor core, which is
                        // short because the cell epilogue is the slower side) the L2 round trip was the kernel's largest single stall
                        // (ncu source page: 27 % of the epilogue warps' samples on the first use of this load).
                        {
                            const bool last_sb = sb == SBN - 1, last_chunk = c == NC - 1;
                            const int nu0 = !last_sb ? u0 + 8 : (!last_chunk ? (c + 1) * 64 + halfsel * (SBN * 8) : halfsel * (SBN * 8));
                            const bool need = (last_sb && last_chunk) ? (s + 1 < S && NC * SBN > 1) : (s > 0);     // (a single sub-block per step: after its store, below)
                            if (need) {
#pragma unroll
                                for (int q = 0; q < 2; ++q) cnext[q] = *reinterpret_cast<const float4*>(c_ptr(s, nu0, q));
                            }
                        }
                        uint32_t hp[4];                               // h of the 8 units as bf16 pairs
                        if constexpr (HIST) {
                            // training: pairs are packed as they are produced (h and the six coefficients), so the 48 coefficient values of
                            // a sub-block are never live at once (the all-at-once form spilled 160 bytes: 805 -> 707 us at B = 4096)
                            uint32_t cop[HIST ? LSTM_NCOEF : 1][4];       // HIST: the backward's coefficients (train_kernels.cuh), packed in pairs
#pragma unroll
                            for (int jp = 0; jp < 4; ++jp) {
                                float hv[2], co[LSTM_NCOEF][2];
#pragma unroll
                                for (int e = 0; e < 2; ++e) {
                                    const int j = 2 * jp + e;
                                    float pi = fi[j], pf = ff[j], pg = fg[j], po = fo[j], cp = 0.0f;
                                    if (s > 0) {
                                        pi += __uint_as_float(gi[j]); pf += __uint_as_float(gf[j]);
                                        pg += __uint_as_float(gg[j]); po += __uint_as_float(go[j]);
                                        cp = cprev[j];
                                    }
                                    const float ig = fast_sigmoid(pi), fgt = fast_sigmoid(pf), gg2 = fast_tanh(pg), og = fast_sigmoid(po);
                                    const float cc = fgt * cp + ig * gg2;
                                    cn[j] = cc;
                                    const float tc = fast_tanh(cc);
                                    hv[e] = og * tc;
                                    if (HIST) {
                                        co[LSTM_CO_A][e] = og * (1.0f - tc * tc);
                                        co[LSTM_CO_BI][e] = gg2 * ig * (1.0f - ig);
                                        co[LSTM_CO_BF][e] = cp * fgt * (1.0f - fgt);
                                        co[LSTM_CO_BG][e] = ig * (1.0f - gg2 * gg2);
                                        co[LSTM_CO_BO][e] = tc * og * (1.0f - og);
                                        co[LSTM_CO_F][e] = fgt;
                                    }
                                }
                                hp[jp] = pack_bf16(hv[0], hv[1]);
                                if (HIST) {
#pragma unroll
                                    for (int k = 0; k < LSTM_NCOEF; ++k) cop[k][jp] = pack_bf16(co[k][0], co[k][1]);
                                }
                            }
#pragma unroll
                            for (int q = 0; q < 2; ++q)
                                *reinterpret_cast<float4*>(c_ptr(s, u0, q)) = make_float4(cn[4 * q], cn[4 * q + 1], cn[4 * q + 2], cn[4 * q + 3]);
                            if (HIST) {
                                bf16* gh = sq.coef_h + (s * hist_step + hist_rb + (u0 >> 3)) * (LSTM_NCOEF * 256) + lane * 8;      // [coefficient][row][8 units]
#pragma unroll
                                for (int k = 0; k < LSTM_NCOEF; ++k)
                                    *reinterpret_cast<uint4*>(gh + k * 256) = make_uint4(cop[k][0], cop[k][1], cop[k][2], cop[k][3]);
                            }
                        } else {
                            // inference: all 8 units in flight at once (more ILP; measured 453 vs 473 us for the pairwise form)
                            float hn[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                float pi = fi[j], pf = ff[j], pg = fg[j], po = fo[j], cp = 0.0f;
                                if (s > 0) {
                                    pi += __uint_as_float(gi[j]); pf += __uint_as_float(gf[j]);
                                    pg += __uint_as_float(gg[j]); po += __uint_as_float(go[j]);
                                    cp = cprev[j];
                                }
                                const float ig = fast_sigmoid(pi), fgt = fast_sigmoid(pf), gg2 = fast_tanh(pg), og = fast_sigmoid(po);
                                const float cc = fgt * cp + ig * gg2;
                                cn[j] = cc;
                                hn[j] = og * fast_tanh(cc);
                            }
#pragma unroll
                            for (int q = 0; q < 2; ++q)
                                *reinterpret_cast<float4*>(c_ptr(s, u0, q)) = make_float4(cn[4 * q], cn[4 * q + 1], cn[4 * q + 2], cn[4 * q + 3]);
                            hp[0] = pack_bf16(hn[0], hn[1]); hp[1] = pack_bf16(hn[2], hn[3]); hp[2] = pack_bf16(hn[4], hn[5]); hp[3] = pack_bf16(hn[6], hn[7]);
                        }
                        if (SBN == 1 && NC == 1 && s + 1 < S) {              // the only sub-block of the step: its new state was just stored
#pragma unroll
                            for (int q = 0; q < 2; ++q) cnext[q] = make_float4(cn[4 * q], cn[4 * q + 1], cn[4 * q + 2], cn[4 * q + 3]);
                        }
                        uint4 o0;
                        o0.x = hp[0]; o0.y = hp[1]; o0.z = hp[2]; o0.w = hp[3];
                        if constexpr (WIDE) {                         // outputs of a pair of sub-blocks leave as 32-byte stores
                            if (sb & 1) {
                                stg256(orow + u0 - 8, o_even, o0);
                                if (HIST && s + 1 < L) stg256(sq.hs_h + (static_cast<long long>(xbase) + tstep + (dir == 0 ? 1 : -1)) * 2 * h + dir * h + u0 - 8, o_even, o0);
                                if (last) stg256(sq.final_h + static_cast<long long>(grow) * 2 * h + dir * h + u0 - 8, o_even, o0);
                            } else o_even = o0;
                        } else {
                        *reinterpret_cast<uint4*>(orow + u0) = o0;
                        // h_s is the *previous* state of the next step's token: stored at that token's row ([rows][2h], zero where a direction
                        // starts), so dW_hh = dGates^T . h_prev is one contraction over token rows with both operands in the same (schedule) order
                        if (HIST && s + 1 < L) *reinterpret_cast<uint4*>(sq.hs_h + (static_cast<long long>(xbase) + tstep + (dir == 0 ? 1 : -1)) * 2 * h + dir * h + u0) = o0;
                        if (last) *reinterpret_cast<uint4*>(sq.final_h + static_cast<long long>(grow) * 2 * h + dir * h + u0) = o0;
                        }
                        st_shared_v4(h_dst + a0, o0.x, o0.y, o0.z, o0.w);
                        if (DUP) st_shared_v4(h_dst + a0 + DUP, o0.x, o0.y, o0.z, o0.w);
                    } else {
                        // finished (or padding) row: carry h forward unchanged so the next step's MMA reads a defined operand
                        const uint4 p0 = ld_shared_v4(h_src + a0);
                        st_shared_v4(h_dst + a0, p0.x, p0.y, p0.z, p0.w);
                        if (DUP) st_shared_v4(h_dst + a0 + DUP, p0.x, p0.y, p0.z, p0.w);
                    }
                }
                if (s > 0) {
                    tcgen05_fence_before();
                    mbar_arrive(&tmem_empty[acc]);
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
                fence_async_smem();                                 // this chunk of h_s (generic-proxy stores) -> visible to tcgen05.mma
                mbar_arrive(&h_ready[c]);
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == W_ALLOC) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// Length-sorted text schedule (inference): counting sort of the questions by descending length, the token offsets in that order and the
// source row of every sorted token row.  One block: B and n_tok are a few thousand / tens of thousands of integers.  Which of two questions
// of EQUAL length comes first is left to the atomics: every question's rows are computed independently, the results do not depend on it.
constexpr int TS_MAX_LEN = 1023;
// 256 threads (a few thousand registers, 300 bytes of shared memory at L_max = 24): the block has to fit NEXT TO a resident 226 KB GEMM CTA — it runs on a side lane underneath the
// video projection; a 1024-thread block waited for the GEMM to drain and put the text staging on the critical path.
__global__ void __launch_bounds__(256)
text_sort_kernel(const int* __restrict__ q_off, int B, int L_max, int* __restrict__ order, int* __restrict__ soff) {
    extern __shared__ int ts_smem[];                               // 3 (L_max + 1) integers, per key = L_max - length (dynamic: the video
    const int nk = L_max + 1;                                      // projection's CTAs leave < 1 KB of shared memory per SM)
    int *cnt = ts_smem, *pos0 = ts_smem + nk, *tok0 = ts_smem + 2 * nk;
    for (int k = threadIdx.x; k < nk; k += blockDim.x) cnt[k] = 0;
    __syncthreads();
    for (int q = threadIdx.x; q < B; q += blockDim.x) {
        int len = __ldg(q_off + q + 1) - __ldg(q_off + q);
        len = len < 0 ? 0 : (len > L_max ? L_max : len);
        atomicAdd(&cnt[L_max - len], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int p = 0, t = 0;
        for (int k = 0; k < nk; ++k) { pos0[k] = p; tok0[k] = t; p += cnt[k]; t += cnt[k] * (L_max - k); cnt[k] = 0; }
        soff[B] = t;
    }
    __syncthreads();
    for (int q = threadIdx.x; q < B; q += blockDim.x) {
        int len = __ldg(q_off + q + 1) - __ldg(q_off + q);
        len = len < 0 ? 0 : (len > L_max ? L_max : len);
        const int k = L_max - len;
        const int i = atomicAdd(&cnt[k], 1);
        order[pos0[k] + i] = q;
        soff[pos0[k] + i] = tok0[k] + i * len;
    }
}
// tok_src[soff[p] + s] = q_off[order[p]] + s: one warp per sorted position
__global__ void text_src_kernel(const int* __restrict__ q_off, int B, const int* __restrict__ order, const int* __restrict__ soff,
                                int* __restrict__ tok_src) {
    const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (p >= B) return;
    const int q = __ldg(order + p), dst = __ldg(soff + p), src = __ldg(q_off + q), len = __ldg(soff + p + 1) - dst;
    for (int s = lane; s < len; s += 32) tok_src[dst + s] = src + s;
}

}  // namespace

static volatile unsigned int* g_lstm_dbg = nullptr;
static int g_lstm_cg = 2;
static int g_lstm_rows = 64;      // questions per CTA of the fused recurrence (128 = round-1 layout)
bool lstm_fused_ok(int precision, int h) { return precision == STAIR_BF16 && h >= 64 && h <= 256 && (h % 64) == 0; }

// seq 0 = video (T steps), seq 1 = text (ragged, L_max steps); either may be disabled with steps = 0.
int launch_lstm_fused(const void* xproj_v, void* vid_out, int T, const void* whh_v_f, const void* whh_v_r,
                      const void* xproj_t, void* tokfeat, void* qfeat, const int* q_off, int L_max, const void* whh_t_f,
                      const void* whh_t_r, float* c_scratch, int B, int h, int run_video, int run_text, int* err_flag, cudaStream_t st,
                      const LstmHist* hist, const int* text_order, const int* text_soff) {
    if (B <= 0 || (!run_video && !run_text)) return STAIR_OK;
    if (!text_order || !text_soff) text_order = text_soff = nullptr;      // (training: the history, hs_h and the BPTT follow the same schedule)
    {   // the cell epilogue reads the projections and writes its outputs with 32-byte accesses
        const void* al[] = {xproj_v, vid_out, xproj_t, tokfeat, qfeat, hist ? hist->hs[0] : nullptr, hist ? hist->hs[1] : nullptr};
        for (const void* q : al) if (reinterpret_cast<uintptr_t>(q) & 31) return STAIR_ERR_ARG;
    }
    LstmFusedParams p;
    p.err_flag = err_flag;
    p.dbg = g_lstm_dbg;
    static int pf = -1;
    if (pf < 0) { const char* e = getenv("STAIR_LSTM_PF"); pf = e ? atoi(e) : 0; }
    p.prefetch = pf;
    LstmSeq v; v.xproj = reinterpret_cast<const bf16*>(xproj_v); v.c = c_scratch; v.out = reinterpret_cast<bf16*>(vid_out);
    v.final_h = nullptr; v.q_off = nullptr; v.order = v.soff = nullptr; v.steps = T; v.B = B; v.h = h;
    LstmSeq t; t.xproj = reinterpret_cast<const bf16*>(xproj_t); t.c = c_scratch + 2LL * ((B + LF_ROWS - 1) / LF_ROWS) * LF_ROWS * h; t.out = reinterpret_cast<bf16*>(tokfeat);
    t.final_h = reinterpret_cast<bf16*>(qfeat); t.q_off = q_off; t.order = text_order; t.soff = text_soff; t.steps = L_max; t.B = B; t.h = h;
    v.coef_h = hist ? reinterpret_cast<bf16*>(hist->gates[0]) : nullptr; v.hs_h = hist ? hist->hs[0] : nullptr; v.hs_dir = hist ? hist->hs_dir[0] : 0;
    t.coef_h = hist ? reinterpret_cast<bf16*>(hist->gates[1]) : nullptr; t.hs_h = hist ? hist->hs[1] : nullptr; t.hs_dir = hist ? hist->hs_dir[1] : 0;
    const void* w[4];
    int nseq = 0;
    // Block dispatch follows the linear block index (x, then y, then z = sequence).  Batch order: video first, the text blocks (all L_max
    // steps long) fill the SMs the video blocks free.  Length-sorted text: text first, longest blocks first (x = 0), and the short video
    // blocks fill in behind the text blocks as those finish (longest-processing-time order: 0.40 -> 0.30 ms makespan at B = 4096).
    // With more frames than words (I3D: T = 64) the video blocks are the long ones and stay first.
    const bool text_first = text_order && L_max > T;
    if (run_text && text_first) { p.seq[nseq] = t; w[2 * nseq] = whh_t_f; w[2 * nseq + 1] = whh_t_r; ++nseq; }
    if (run_video) { p.seq[nseq] = v; w[2 * nseq] = whh_v_f; w[2 * nseq + 1] = whh_v_r; ++nseq; }
    if (run_text && !text_first) { p.seq[nseq] = t; w[2 * nseq] = whh_t_f; w[2 * nseq + 1] = whh_t_r; ++nseq; }
    if (nseq == 1) { p.seq[1] = p.seq[0]; w[2] = w[0]; w[3] = w[1]; }
    CUtensorMap tm[4];
    for (int i = 0; i < 4; ++i) STAIR_TRY(make_tmap_bf16_2d(&tm[i], w[i], h, 4ULL * h, h, 64, 256));
    const int smem = 2 * (h / 64) * LF_KB_BYTES + LF_STAGES * LF_W_STAGE_BYTES + 256 + 1024;
    // variants: 0 = 128 rows, 2 column groups (round-1 default); 1 = 128 rows, 4 column groups (comparison); 2 = training history,
    // 128 rows; 3 = 64 rows per CTA, 4 column groups; 4 = training history, 64 rows
    // 5 / 6 = 64 rows, 8 column groups (16 epilogue warps, 96 registers; comparison), inference / training history
    static int configured[7] = {0, 0, 0, 0, 0, 0, 0};
    const bool r64 = g_lstm_rows == 64;
    const int vi = hist ? (r64 ? (g_lstm_cg == 8 ? 6 : 4) : 2) : (r64 ? (g_lstm_cg == 8 ? 5 : 3) : (g_lstm_cg == 4 ? 1 : 0));
    if (configured[vi] < smem) {
        cudaError_t e;
        switch (vi) {
        case 6: e = cudaFuncSetAttribute(lstm_fused_kernel<8, true, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); break;
        case 5: e = cudaFuncSetAttribute(lstm_fused_kernel<8, false, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); break;
        case 4: e = cudaFuncSetAttribute(lstm_fused_kernel<4, true, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); break;
        case 3: e = cudaFuncSetAttribute(lstm_fused_kernel<4, false, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); break;
        case 2: e = cudaFuncSetAttribute(lstm_fused_kernel<2, true, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); break;
        case 1: e = cudaFuncSetAttribute(lstm_fused_kernel<4, false, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); break;
        default: e = cudaFuncSetAttribute(lstm_fused_kernel<2, false, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); break;
        }
        if (e != cudaSuccess) return STAIR_ERR_CUDA;
        configured[vi] = smem;
    }
    const int vrows = r64 ? 64 : 128;
    dim3 grid((B + vrows - 1) / vrows, 2, nseq);
    switch (vi) {
    case 6: lstm_fused_kernel<8, true, 64><<<grid, 640, smem, st>>>(tm[0], tm[1], tm[2], tm[3], p); break;
    case 5: lstm_fused_kernel<8, false, 64><<<grid, 640, smem, st>>>(tm[0], tm[1], tm[2], tm[3], p); break;
    case 4: lstm_fused_kernel<4, true, 64><<<grid, 384, smem, st>>>(tm[0], tm[1], tm[2], tm[3], p); break;
    case 3: lstm_fused_kernel<4, false, 64><<<grid, 384, smem, st>>>(tm[0], tm[1], tm[2], tm[3], p); break;
    case 2: lstm_fused_kernel<2, true, 128><<<grid, 128 + 128 * 2, smem, st>>>(tm[0], tm[1], tm[2], tm[3], p); break;
    case 1: lstm_fused_kernel<4, false, 128><<<grid, 128 + 128 * 4, smem, st>>>(tm[0], tm[1], tm[2], tm[3], p); break;
    default: lstm_fused_kernel<2, false, 128><<<grid, 128 + 128 * 2, smem, st>>>(tm[0], tm[1], tm[2], tm[3], p); break;
    }
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

bool text_sort_ok(int L_max) { return L_max >= 1 && L_max <= TS_MAX_LEN; }
int launch_text_sort(const int* q_off, int B, int L_max, int* order, int* soff, int* tok_src, cudaStream_t st) {
    if (B <= 0) return STAIR_OK;
    if (!text_sort_ok(L_max)) return STAIR_ERR_ARG;
    text_sort_kernel<<<1, 256, 3 * (L_max + 1) * sizeof(int), st>>>(q_off, B, L_max, order, soff);
    STAIR_CHECK_LAUNCH();
    text_src_kernel<<<(B + 3) / 4, 128, 0, st>>>(q_off, B, order, soff, tok_src);
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

}  // namespace stair

extern "C" int stair_lstm_debug(unsigned int* pinned_buf) { stair::g_lstm_dbg = pinned_buf; return STAIR_OK; }
extern "C" int stair_lstm_rows(int rows) { if (rows != 64 && rows != 128) return STAIR_ERR_ARG; stair::g_lstm_rows = rows; return STAIR_OK; }
extern "C" int stair_lstm_colgroups(int cg) { if (cg != 2 && cg != 4 && cg != 8) return STAIR_ERR_ARG; stair::g_lstm_cg = cg; return STAIR_OK; }
