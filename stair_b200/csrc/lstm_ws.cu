// Weight-stationary BiLSTM recurrence (bf16 path, h = 256): nn.LSTM's time loop (video_nmn/module_net.py:39-47,147-163) with W_hh
// RESIDENT in shared memory.
//
// Why a second recurrence kernel: in lstm_fused.cu a CTA owns 64 questions and ALL 4h gate columns, so it re-streams the whole W_hh
// (512 KB) from L2 every time step — 5 us of a 12.6 us step, and the M = 128 MMA carries only 64 distinct rows.  Here the split is the
// other way round: a CLUSTER of four CTAs serves one direction of one encoder; CTA c of the cluster owns hidden units [64 c, 64 c + 64)
// = the 256 gate columns (i, f, g, o x 64 units) of chunk c, and keeps that 256 x 256 slice of W_hh (128 KB, gate-interleaved copy
// STAIR_W_*_WHHI_*) in shared memory for the whole kernel.  Per step and 128-question block a CTA then needs only the block's h_{t-1}
// tile (128 x 256 bf16 = 64 KB) instead of 512 KB of weights: 16x less operand streaming per question-step, and all 128 MMA rows are
// distinct questions.
//
// h is exchanged through a small dense global buffer hx[dir][parity][question][h] (L2-resident, 16 MB at B = 4096): the epilogue of
// chunk c writes its 64 units of h_t there (and to the encoder output), every CTA of the cluster TMA-loads the full 128 x 256 tile of
// the block for step t + 1.  Synchronisation is per (block, step): an mbarrier per local block in every CTA, on which each epilogue warp
// of each of the four CTAs arrives (release.cluster, remote arrive through mapa) once its slice of h_t is in global memory; the TMA
// producer waits (acquire.cluster), issues a generic->async proxy fence and loads.  h_t goes to parity t & 1: by the time anybody writes
// h_{t+2}, every CTA has finished loading h_t (it had to, to produce the h_{t+1} slice that the writer's own MMA waited for).
// A cluster walks (step, local block) in lexicographic order, so while block b waits for its peers' slices the CTA works on its
// other blocks; TMEM is double-buffered (2 x 256 columns), so the MMA of the next block overlaps the cell epilogue of this one.
//
// Roles per CTA (384 threads): warps 0-7 cell epilogue (TMEM lane quarter = warp % 4, 32-unit column group = warp / 4), warp 8 TMA
// producer, warp 9 tcgen05.mma issuer, warp 10 TMEM allocator.
// The cell arithmetic, the BPTT history written in training (HIST) and all global layouts are those of lstm_fused.cu, so the two kernels
// are interchangeable (tests/test_forward_gpu.py::test_fused_lstm_matches_stepwise, tests/test_train_gpu.py).
#include "nmn_kernels.cuh"
#include "tc_ptx.cuh"
#include "train_kernels.cuh"
#include <cstdlib>
#include <cstdio>

namespace stair {

namespace {

constexpr int WS_ROWS = 128;                       // questions per block
constexpr int WS_H = 256;                          // hidden units per direction (H = 512)
constexpr int WS_NC = WS_H / 64;                   // chunks = CTAs per cluster = k-blocks of h
constexpr int WS_W_KB_BYTES = 256 * 64 * 2;        // one k-block of the CTA's W slice: 256 gate rows x 64 k = 32 KiB
constexpr int WS_A_KB_BYTES = WS_ROWS * 64 * 2;    // one k-block of a block's h tile: 16 KiB
constexpr int WS_A_STAGES = 4;
constexpr int WS_MAX_LB = 16;                      // local blocks per cluster
constexpr uint32_t WS_IDESC = make_idesc_bf16(128, 256);
constexpr int WS_SMEM = WS_NC * WS_W_KB_BYTES + WS_A_STAGES * WS_A_KB_BYTES + 1024 /*barriers, step table*/ + 1024 /*alignment*/;
__host__ __device__ constexpr int ws_threads(int ew) { return (ew + 4) * 32; }

struct WsParams {
    const bf16* xproj;      // [rows, 8h]: W_ih x + b for both directions (fwd gates | reverse gates)
    float* c;               // cell state scratch (fp32), [dir][block][chunk][unit/4][row][4]
    bf16* out;              // video: [B*T, 2h] ; text: token_feature [n_tok, 2h]
    bf16* final_h;          // text: question_feature [B, 2h] ; video: null
    const int* q_off;       // text: [B+1] token offsets (ragged) ; video: null
    bf16* hx;               // exchange buffer [2 dirs][2 parities][nblk * 128][h]
    int steps;              // video: T ; text: L_max
    int B, nblk;
    bf16* coef_h; bf16* hs_h;      // training history (HIST), layouts of lstm_fused.cu / train_kernels.cuh
    int* err_flag;
};

__device__ __forceinline__ float ws_tanh(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ws_sigmoid(float x) { return fmaf(0.5f, ws_tanh(0.5f * x), 0.5f); }
__device__ __forceinline__ void ws_tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ws_unpack8(const uint4& r, float (&f)[8]) {
    const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(hh[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
// wait on an mbarrier whose arrivals come from other CTAs of the cluster, bounded like mbar_wait.  The spin is RELAXED and one
// acquire fence follows the successful try: an acquire.cluster try_wait invalidates L1 on every iteration (CCTL.IVALL, 1.2 M times per
// launch in the first version), which is where the epilogue warps of the same SM keep the other halves of their xproj sectors.
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity, int* err_flag, int code, bool acquire = true) {
    uint32_t spins = 0;
    uint64_t t0 = 0;
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.relaxed.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) {
            if (acquire) asm volatile("fence.acq_rel.cluster;" ::: "memory");
            return;
        }
        if ((++spins & 0x3FFu) == 0) {
            uint64_t now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000ull) {
                if (err_flag) atomicExch(err_flag, code);
                __threadfence_system();
                __trap();
            }
        }
    }
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async.global;" ::: "memory"); }     // generic <-> async proxy, global state space (the h exchange buffer)

// EW = cell-epilogue warps.  The cell epilogue is a latency chain (TMEM read -> 5 transcendentals -> stores); with 8 warps (2 per
// scheduler) it runs at 23 % of the issue slots.  The inference form processes a sub-block of 8 units GATE BY GATE (i, g -> i*g; f, o ->
// c, h) so that at most two gates' operands are live: it fits 16 warps (4 per scheduler) in 96 registers without spilling.  The training
// form (HIST: six coefficients per unit) keeps 8 warps and all four gates in flight.
template <bool HIST, int EW>
__global__ void __cluster_dims__(WS_NC, 1, 1) __launch_bounds__(ws_threads(EW), 1)
lstm_ws_kernel(const __grid_constant__ CUtensorMap tmW0, const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmHX,
               const WsParams p) {
    constexpr int h = WS_H;
    constexpr int WS_EPI_WARPS = EW;
    constexpr int WS_THREADS = ws_threads(EW);
    const uint32_t chunk = cluster_ctarank();                      // 64-unit chunk this CTA owns
    const int cl = blockIdx.x / WS_NC, ncl = gridDim.x / WS_NC;
    const int dir = cl & 1, g = cl >> 1, Gd = ncl >> 1;            // clusters alternate directions; g-th cluster of its direction
    const bool ragged = p.q_off != nullptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int nlb = 0;
    for (int b = g; b < p.nblk && nlb < WS_MAX_LB; b += Gd) ++nlb;  // local blocks: g, g + Gd, ...

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint8_t* sW = smem;                                            // [4 kb][256 x 64] bf16, resident
    uint8_t* sA = smem + WS_NC * WS_W_KB_BYTES;                    // [WS_A_STAGES][128 x 64] bf16 ring of h k-blocks
    uint64_t* w_full = reinterpret_cast<uint64_t*>(sA + WS_A_STAGES * WS_A_KB_BYTES);
    uint64_t* a_full = w_full + 1;
    uint64_t* a_empty = a_full + WS_A_STAGES;
    uint64_t* tmem_full = a_empty + WS_A_STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* hready = tmem_empty + 2;                             // [WS_MAX_LB]: h_s of local block lb complete (all 4 chunks), phase = s
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(hready + WS_MAX_LB);
    int* s_steps = reinterpret_cast<int*>(tmem_ptr_smem + 1);      // [WS_MAX_LB] steps of each local block (its longest question)

    if (threadIdx.x == 0) {
        mbar_init(w_full, 1);
        for (int s = 0; s < WS_A_STAGES; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], WS_EPI_WARPS * 32); }
        for (int lb = 0; lb < WS_MAX_LB; ++lb) mbar_init(&hready[lb], WS_NC * WS_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < WS_MAX_LB) s_steps[threadIdx.x] = (ragged || threadIdx.x >= nlb) ? 0 : p.steps;
    if (warp == WS_EPI_WARPS + 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_ptr_smem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    if (ragged) {                                                  // a block runs as many steps as its longest question
        for (int i = threadIdx.x; i < nlb * WS_ROWS; i += WS_THREADS) {
            const int r = (g + (i / WS_ROWS) * Gd) * WS_ROWS + (i % WS_ROWS);
            if (r < p.B) atomicMax(&s_steps[i / WS_ROWS], __ldg(p.q_off + r + 1) - __ldg(p.q_off + r));
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();                                            // every CTA's barriers are initialised before any remote arrive
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    int S = 0;
    for (int lb = 0; lb < nlb; ++lb) S = max(S, s_steps[lb]);
    const long long hx_rows = static_cast<long long>(p.nblk) * WS_ROWS;

    if (warp == WS_EPI_WARPS) {
        if (lane == 0) {
            // ===================== TMA producer: the resident W slice once, then h_{s-1} tiles block by block =====================
            mbar_arrive_expect_tx(w_full, WS_NC * WS_W_KB_BYTES);
            for (int kb = 0; kb < WS_NC; ++kb) {
                if (dir == 0) tma_load_2d(sW + kb * WS_W_KB_BYTES, &tmW0, w_full, kb * 64, static_cast<int>(chunk) * 256);
                else tma_load_2d(sW + kb * WS_W_KB_BYTES, &tmW1, w_full, kb * 64, static_cast<int>(chunk) * 256);
            }
            int stage = 0; uint32_t phase = 0;
            for (int s = 1; s < S; ++s)
                for (int lb = 0; lb < nlb; ++lb) {
                    if (s >= s_steps[lb]) continue;
                    const int b = g + lb * Gd;
                    mbar_wait_cluster(&hready[lb], static_cast<uint32_t>((s - 1) & 1), p.err_flag, 301);     // h_{s-1}[b]: all four slices written
                    fence_proxy_async_all();                       // generic-proxy global writes (other CTAs) -> this CTA's TMA reads
                    const int row = static_cast<int>(((dir * 2 + ((s - 1) & 1)) * hx_rows) + static_cast<long long>(b) * WS_ROWS);
                    for (int kb = 0; kb < WS_NC; ++kb) {
                        mbar_wait(&a_empty[stage], phase ^ 1, p.err_flag, 302);
                        mbar_arrive_expect_tx(&a_full[stage], WS_A_KB_BYTES);
                        tma_load_2d(sA + stage * WS_A_KB_BYTES, &tmHX, &a_full[stage], kb * 64, row);
                        if (++stage == WS_A_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
        }
    } else if (warp == WS_EPI_WARPS + 1) {
        if (lane == 0) {
            // ===================== MMA issuer: gates[128, 256] = h_{s-1}[128, 256] . W_chunk[256, 256]^T =====================
            mbar_wait(w_full, 0, p.err_flag, 303);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int s = 1; s < S; ++s)
                for (int lb = 0; lb < nlb; ++lb) {
                    if (s >= s_steps[lb]) continue;
                    mbar_wait(&tmem_empty[acc], acc_phase ^ 1, p.err_flag, 304);
                    tcgen05_fence_after();
                    const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * 256);
                    for (int kb = 0; kb < WS_NC; ++kb) {
                        mbar_wait(&a_full[stage], phase, p.err_flag, 305);
                        tcgen05_fence_after();
                        const uint64_t adesc = make_umma_desc_kmajor_sw128(smem_u32(sA + stage * WS_A_KB_BYTES));
                        const uint64_t bdesc = make_umma_desc_kmajor_sw128(smem_u32(sW + kb * WS_W_KB_BYTES));
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, WS_IDESC, (kb | k) != 0 ? 1u : 0u);
                        umma_commit(&a_empty[stage]);
                        if (++stage == WS_A_STAGES) { stage = 0; phase ^= 1; }
                    }
                    umma_commit(&tmem_full[acc]);
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
        }
    } else if (warp == WS_EPI_WARPS + 3) {
        // ===================== L2 prefetcher: the chunk's input-projection columns of block-step (s, lb), one round ahead =====================
        // xproj (402 MB at B = 4096) streams from HBM exactly once; the cell epilogue is a dependent chain per step, so an HBM miss
        // (~0.7 us) in front of every sub-block is exposed.  This warp requests the 128 rows x 4 gates x 128 bytes of a block-step as soon
        // as h_{s-1} of that block is complete (the same event its MMA waits for), i.e. one exchange + MMA ahead of the epilogue.
        for (int s = 0; s < S; ++s)
            for (int lb = 0; lb < nlb; ++lb) {
                if (s >= s_steps[lb]) continue;
                if (s > 0) mbar_wait_cluster(&hready[lb], static_cast<uint32_t>((s - 1) & 1), p.err_flag, 307, false);
                const int b = g + lb * Gd;
                for (int r = lane; r < WS_ROWS; r += 32) {
                    const int grow = b * WS_ROWS + r;
                    if (grow >= p.B) continue;
                    int base, L = p.steps;
                    if (ragged) { base = __ldg(p.q_off + grow); L = __ldg(p.q_off + grow + 1) - base; }
                    else base = grow * p.steps;
                    if (s >= L) continue;
                    const bf16* xrow = p.xproj + (static_cast<long long>(base) + (dir == 0 ? s : L - 1 - s)) * 8 * h + dir * 4 * h + static_cast<int>(chunk) * 64;
#pragma unroll
                    for (int gte = 0; gte < 4; ++gte) asm volatile("prefetch.global.L2 [%0];" ::"l"(xrow + gte * h));
                }
            }
    } else if (warp < WS_EPI_WARPS) {
        // ===================== cell epilogue: thread = (question row, 32 of the chunk's 64 units) =====================
        const int quarter = warp & 3, cg = warp >> 2;
        const int row = quarter * 32 + lane;
        constexpr int UPT = 64 / (EW / 4);                         // hidden units per thread and block-step
        constexpr int SBN = UPT / 8;                               // 8-unit sub-blocks per thread
        const long long RB = (p.B + 127) / 128 * 4;                // 32-row blocks per (step, direction) of the BPTT history
        const long long hist_step = 2 * RB * (h >> 3);
        int acc = 0; uint32_t acc_phase = 0;
        for (int s = 0; s < S; ++s)
            for (int lb = 0; lb < nlb; ++lb) {
                if (s >= s_steps[lb]) continue;
                const int b = g + lb * Gd;
                const int grow = b * WS_ROWS + row;
                const bool valid = grow < p.B;
                int base = 0, L = p.steps;
                if (ragged) { base = valid ? __ldg(p.q_off + grow) : 0; L = valid ? __ldg(p.q_off + grow + 1) - base : 0; }
                else base = grow * p.steps;
                const bool active = valid && s < L;
                const long long tokrow = static_cast<long long>(base) + (dir == 0 ? s : L - 1 - s);
                const bf16* xrow = p.xproj + tokrow * 8 * h + dir * 4 * h;
                bf16* orow = p.out + tokrow * 2 * h + dir * h;
                const bool last = ragged && s == L - 1;
                float* cblk = p.c + ((static_cast<long long>(dir) * p.nblk + b) * WS_NC + chunk) * (64LL * WS_ROWS) + row * 4;     // [unit/4][row][4]
                bf16* hxrow = p.hx + ((dir * 2 + (s & 1)) * hx_rows + grow) * h;
                const long long hist_rb = (static_cast<long long>(dir) * RB + (grow >> 5)) * (h >> 3);
                const int ul0 = cg * UPT;                            // first local unit of this thread
                if constexpr (!HIST) {
                    // ---- inference: gate-serial sub-blocks (low register footprint, 16 warps) ----
                    (void)hist_rb; (void)hist_step;
                    if (s > 0) {
                        mbar_wait(&tmem_full[acc], acc_phase, p.err_flag, 306);
                        tcgen05_fence_after();
                    }
#pragma unroll
                    for (int sb = 0; sb < SBN; ++sb) {
                        const int ul = ul0 + sb * 8;
                        const int u0 = static_cast<int>(chunk) * 64 + ul;
                        uint4 xi, xf, xg, xo;
                        float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f), c1 = c0;
                        if (active) {
                            xi = __ldg(reinterpret_cast<const uint4*>(xrow + u0));
                            xg = __ldg(reinterpret_cast<const uint4*>(xrow + 2 * h + u0));
                            xf = __ldg(reinterpret_cast<const uint4*>(xrow + h + u0));
                            xo = __ldg(reinterpret_cast<const uint4*>(xrow + 3 * h + u0));
                            if (s > 0) {
                                c0 = *reinterpret_cast<const float4*>(cblk + (ul / 4) * (WS_ROWS * 4));
                                c1 = *reinterpret_cast<const float4*>(cblk + (ul / 4 + 1) * (WS_ROWS * 4));
                            }
                        }
                        const uint32_t t = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * 256 + ul);
                        uint32_t ra[8], rb[8];
                        float a[8], fx[8];
                        if (s > 0) { ws_tmem_ld8(t, ra); ws_tmem_ld8(t + 128, rb); tmem_ld_wait(); }     // gates i and g (.sync.aligned: whole warp)
                        if (active) {
                            ws_unpack8(xi, fx);
#pragma unroll
                            for (int j = 0; j < 8; ++j) a[j] = ws_sigmoid(s > 0 ? fx[j] + __uint_as_float(ra[j]) : fx[j]);
                            ws_unpack8(xg, fx);
#pragma unroll
                            for (int j = 0; j < 8; ++j) a[j] *= ws_tanh(s > 0 ? fx[j] + __uint_as_float(rb[j]) : fx[j]);
                        }
                        if (s > 0) { ws_tmem_ld8(t + 64, ra); ws_tmem_ld8(t + 192, rb); tmem_ld_wait(); }    // gates f and o
                        if (active) {
                            const float cprev[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
                            float cn[8];
                            ws_unpack8(xf, fx);
#pragma unroll
                            for (int j = 0; j < 8; ++j) cn[j] = fmaf(ws_sigmoid(s > 0 ? fx[j] + __uint_as_float(ra[j]) : fx[j]), cprev[j], a[j]);
                            *reinterpret_cast<float4*>(cblk + (ul / 4) * (WS_ROWS * 4)) = make_float4(cn[0], cn[1], cn[2], cn[3]);
                            *reinterpret_cast<float4*>(cblk + (ul / 4 + 1) * (WS_ROWS * 4)) = make_float4(cn[4], cn[5], cn[6], cn[7]);
                            ws_unpack8(xo, fx);
#pragma unroll
                            for (int j = 0; j < 8; ++j) a[j] = ws_sigmoid(s > 0 ? fx[j] + __uint_as_float(rb[j]) : fx[j]) * ws_tanh(cn[j]);
                            const uint4 o0 = make_uint4(pack_bf16(a[0], a[1]), pack_bf16(a[2], a[3]), pack_bf16(a[4], a[5]), pack_bf16(a[6], a[7]));
                            *reinterpret_cast<uint4*>(orow + u0) = o0;
                            *reinterpret_cast<uint4*>(hxrow + u0) = o0;          // the next step's A operand (all four CTAs of the cluster load it)
                            if (last) *reinterpret_cast<uint4*>(p.final_h + static_cast<long long>(grow) * 2 * h + dir * h + u0) = o0;
                        }
                    }
                } else {
                // ---- training (history): all four gates of a sub-block in flight, operands of two sub-blocks double-buffered ----
                // operands of the first two sub-blocks are requested before waiting for the tensor core
                uint4 xq[2][4];
                float4 cq[2][2];
                if (active) {
#pragma unroll
                    for (int sb = 0; sb < 2; ++sb) {
                        const int u0 = static_cast<int>(chunk) * 64 + ul0 + sb * 8;
#pragma unroll
                        for (int gte = 0; gte < 4; ++gte) xq[sb][gte] = __ldg(reinterpret_cast<const uint4*>(xrow + gte * h + u0));
                        if (s > 0) {
#pragma unroll
                            for (int q = 0; q < 2; ++q) cq[sb][q] = *reinterpret_cast<const float4*>(cblk + ((ul0 + sb * 8) / 4 + q) * (WS_ROWS * 4));
                        }
                    }
                }
                if (s > 0) {
                    mbar_wait(&tmem_full[acc], acc_phase, p.err_flag, 306);
                    tcgen05_fence_after();
                }
#pragma unroll
                for (int sb = 0; sb < SBN; ++sb) {
                    const int ul = ul0 + sb * 8;                     // local unit (0..63) of the 8 units of this sub-block
                    const int u0 = static_cast<int>(chunk) * 64 + ul; // global hidden unit
                    uint32_t gi[8], gf[8], gg[8], go[8];
                    if (s > 0) {                                     // tcgen05.ld / wait::ld are .sync.aligned: the whole warp, converged
                        const uint32_t t = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * 256 + ul);
                        ws_tmem_ld8(t, gi); ws_tmem_ld8(t + 64, gf); ws_tmem_ld8(t + 128, gg); ws_tmem_ld8(t + 192, go);
                        tmem_ld_wait();
                    }
                    if (active) {
                        float fi[8], ff[8], fg[8], fo[8], cn[8];
                        ws_unpack8(xq[sb & 1][0], fi); ws_unpack8(xq[sb & 1][1], ff); ws_unpack8(xq[sb & 1][2], fg); ws_unpack8(xq[sb & 1][3], fo);
                        const float cprev[8] = {cq[sb & 1][0].x, cq[sb & 1][0].y, cq[sb & 1][0].z, cq[sb & 1][0].w,
                                                cq[sb & 1][1].x, cq[sb & 1][1].y, cq[sb & 1][1].z, cq[sb & 1][1].w};
                        if (sb + 2 < SBN) {                          // refill this slot with the operands of sub-block sb + 2
                            const int un = u0 + 16;
#pragma unroll
                            for (int gte = 0; gte < 4; ++gte) xq[sb & 1][gte] = __ldg(reinterpret_cast<const uint4*>(xrow + gte * h + un));
                            if (s > 0) {
#pragma unroll
                                for (int q = 0; q < 2; ++q) cq[sb & 1][q] = *reinterpret_cast<const float4*>(cblk + ((ul + 16) / 4 + q) * (WS_ROWS * 4));
                            }
                        }
                        uint32_t hp[4];
                        if constexpr (HIST) {
                            uint32_t cop[LSTM_NCOEF][4];
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
                                    const float ig = ws_sigmoid(pi), fgt = ws_sigmoid(pf), gg2 = ws_tanh(pg), og = ws_sigmoid(po);
                                    const float cc = fgt * cp + ig * gg2;
                                    cn[j] = cc;
                                    const float tc = ws_tanh(cc);
                                    hv[e] = og * tc;
                                    co[LSTM_CO_A][e] = og * (1.0f - tc * tc);
                                    co[LSTM_CO_BI][e] = gg2 * ig * (1.0f - ig);
                                    co[LSTM_CO_BF][e] = cp * fgt * (1.0f - fgt);
                                    co[LSTM_CO_BG][e] = ig * (1.0f - gg2 * gg2);
                                    co[LSTM_CO_BO][e] = tc * og * (1.0f - og);
                                    co[LSTM_CO_F][e] = fgt;
                                }
                                hp[jp] = pack_bf16(hv[0], hv[1]);
#pragma unroll
                                for (int k = 0; k < LSTM_NCOEF; ++k) cop[k][jp] = pack_bf16(co[k][0], co[k][1]);
                            }
                            bf16* gh = p.coef_h + (s * hist_step + hist_rb + (u0 >> 3)) * (LSTM_NCOEF * 256) + lane * 8;      // [coefficient][row][8 units]
#pragma unroll
                            for (int k = 0; k < LSTM_NCOEF; ++k)
                                *reinterpret_cast<uint4*>(gh + k * 256) = make_uint4(cop[k][0], cop[k][1], cop[k][2], cop[k][3]);
                        } else {
                            float hn[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                float pi = fi[j], pf = ff[j], pg = fg[j], po = fo[j], cp = 0.0f;
                                if (s > 0) {
                                    pi += __uint_as_float(gi[j]); pf += __uint_as_float(gf[j]);
                                    pg += __uint_as_float(gg[j]); po += __uint_as_float(go[j]);
                                    cp = cprev[j];
                                }
                                const float ig = ws_sigmoid(pi), fgt = ws_sigmoid(pf), gg2 = ws_tanh(pg), og = ws_sigmoid(po);
                                const float cc = fgt * cp + ig * gg2;
                                cn[j] = cc;
                                hn[j] = og * ws_tanh(cc);
                            }
                            hp[0] = pack_bf16(hn[0], hn[1]); hp[1] = pack_bf16(hn[2], hn[3]); hp[2] = pack_bf16(hn[4], hn[5]); hp[3] = pack_bf16(hn[6], hn[7]);
                        }
#pragma unroll
                        for (int q = 0; q < 2; ++q)
                            *reinterpret_cast<float4*>(cblk + (ul / 4 + q) * (WS_ROWS * 4)) = make_float4(cn[4 * q], cn[4 * q + 1], cn[4 * q + 2], cn[4 * q + 3]);
                        const uint4 o0 = make_uint4(hp[0], hp[1], hp[2], hp[3]);
                        *reinterpret_cast<uint4*>(orow + u0) = o0;
                        *reinterpret_cast<uint4*>(hxrow + u0) = o0;          // the next step's A operand (all four CTAs of the cluster load it)
                        if (HIST && s + 1 < L) *reinterpret_cast<uint4*>(p.hs_h + (tokrow + (dir == 0 ? 1 : -1)) * 2 * h + dir * h + u0) = o0;
                        if (last) *reinterpret_cast<uint4*>(p.final_h + static_cast<long long>(grow) * 2 * h + dir * h + u0) = o0;
                    }
                }
                }
                if (s > 0) {
                    tcgen05_fence_before();
                    mbar_arrive(&tmem_empty[acc]);
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
                // this warp's slice of h_s[b] is in global memory: tell the TMA producers of all four CTAs (finished / padding rows simply
                // do not write: a row of the GEMM only feeds the same row of the gates, and nothing reads a finished row again)
                if (s + 1 < s_steps[lb]) {
                    fence_proxy_async_all();
                    __syncwarp();
                    if (lane == 0) {
#pragma unroll
                        for (int r = 0; r < WS_NC; ++r) mbar_arrive_cluster(mapa_shared(smem_u32(&hready[lb]), static_cast<uint32_t>(r)));
                    }
                }
            }
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();                                            // no CTA exits while a peer may still arrive on its barriers
    if (warp == WS_EPI_WARPS + 2) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}


}  // namespace

static int lstm_ws_default() {                       // STAIR_LSTM_WS=0|1 overrides the default at library load (A/B runs)
    const char* e = getenv("STAIR_LSTM_WS");
    return (e && (e[0] == '0' || e[0] == '1')) ? e[0] - '0' : 0;
}
static int g_lstm_ws = lstm_ws_default();      // 1 = weight-stationary cluster kernel when eligible, 0 = lstm_fused.cu always

bool lstm_ws_ok(int precision, int h, int B) {
    return g_lstm_ws && precision == STAIR_BF16 && h == WS_H && B > 0;
}
long long lstm_ws_hx_bytes(int B) { return 2LL * 2 * ((B + WS_ROWS - 1) / WS_ROWS) * WS_ROWS * WS_H * 2; }

// One encoder (both directions).  q_off == nullptr: video (T steps for everybody); else text (ragged, L_max steps).
// c_scratch: >= 2 * nblk * 128 * h floats; hx: >= lstm_ws_hx_bytes(B) bytes.
int launch_lstm_ws(const void* xproj, void* out, void* final_h, const int* q_off, int steps, const void* whh_f, const void* whh_r,
                   float* c_scratch, void* hx, int B, int h, int* err_flag, cudaStream_t st, bf16* coef_h, bf16* hs_h) {
    if (B <= 0 || steps <= 0) return STAIR_OK;
    if (h != WS_H) return STAIR_ERR_UNSUPPORTED;
    static int max_clusters = -1;
    static bool configured[2] = {false, false};
    const bool hist = coef_h != nullptr;
    if (!configured[hist ? 1 : 0]) {
        cudaError_t e = hist ? cudaFuncSetAttribute(lstm_ws_kernel<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM)
                             : cudaFuncSetAttribute(lstm_ws_kernel<false, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM);
        if (e != cudaSuccess) return STAIR_ERR_CUDA;
        configured[hist ? 1 : 0] = true;
    }
    if (max_clusters < 0) {
        // clusters of four 200 KB CTAs that can be resident at once (GPC boundaries: fewer than SMs / 4)
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(WS_NC * 64); cfg.blockDim = dim3(ws_threads(16)); cfg.dynamicSmemBytes = WS_SMEM;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = WS_NC; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = 0;
        if (cudaFuncSetAttribute(lstm_ws_kernel<false, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM) != cudaSuccess || cudaOccupancyMaxActiveClusters(&n, lstm_ws_kernel<false, 16>, &cfg) != cudaSuccess || n < 2) { cudaGetLastError(); n = 32; }
        max_clusters = n;
    }
    const int nblk = (B + WS_ROWS - 1) / WS_ROWS;
    int Gd = max_clusters / 2;                                      // clusters per direction
    if (Gd > nblk) Gd = nblk;
    if (Gd < 1) Gd = 1;
    if ((nblk + Gd - 1) / Gd > WS_MAX_LB) Gd = (nblk + WS_MAX_LB - 1) / WS_MAX_LB;      // more clusters than are resident at once: still correct (clusters are independent)
    WsParams p;
    p.xproj = reinterpret_cast<const bf16*>(xproj); p.c = c_scratch; p.out = reinterpret_cast<bf16*>(out); p.final_h = reinterpret_cast<bf16*>(final_h);
    p.q_off = q_off; p.hx = reinterpret_cast<bf16*>(hx); p.steps = steps; p.B = B; p.nblk = nblk; p.coef_h = coef_h; p.hs_h = hs_h; p.err_flag = err_flag;
    CUtensorMap tw0, tw1, thx;
    STAIR_TRY(make_tmap_bf16_2d(&tw0, whh_f, h, 4ULL * h, h, 64, 256));
    STAIR_TRY(make_tmap_bf16_2d(&tw1, whh_r, h, 4ULL * h, h, 64, 256));
    STAIR_TRY(make_tmap_bf16_2d(&thx, hx, h, 4ULL * nblk * WS_ROWS, h, 64, WS_ROWS));
    const int grid = WS_NC * 2 * Gd;
    static bool said = false;
    if (!said && getenv("STAIR_DEBUG")) { fprintf(stderr, "lstm_ws: max resident clusters %d, clusters per direction %d, blocks %d, grid %d\n", max_clusters, Gd, nblk, grid); said = true; }
    if (hist) lstm_ws_kernel<true, 8><<<grid, ws_threads(8), WS_SMEM, st>>>(tw0, tw1, thx, p);
    else lstm_ws_kernel<false, 16><<<grid, ws_threads(16), WS_SMEM, st>>>(tw0, tw1, thx, p);
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

}  // namespace stair

extern "C" int stair_set_lstm_ws(int on) { stair::g_lstm_ws = on ? 1 : 0; return STAIR_OK; }
