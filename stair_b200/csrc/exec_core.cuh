// Core of the batched interpreter shared by the forward (executor.cu) and backward (executor_bwd.cu) entry points:
// scratch plan, execution context, GEMM helpers and the per-group forward (video_nmn/module_net.py:94-133 per module).
#pragma once
#include "nmn_kernels.cuh"

namespace stair {
namespace ex {


extern thread_local long long t_last_launches;   // kernels launched by the last entry-point call on this thread
extern int g_lstm_impl;                     // 0 = fused persistent recurrence when eligible, 1 = per-step GEMM + cell kernels
extern int g_text_sort;                     // 1 = inference text recurrence over length-sorted questions (default)
extern int g_fuse_sum;                      // 1 = Filter's frame sum in the epilogue of its second Linear (inference)
constexpr int LANES = 8;                 // independent groups of one wave execute concurrently on up to LANES streams (main + side)
constexpr long long ROW_CAP = 65536;     // frame rows per chunk of a VID-typed group
constexpr long long VEC_CAP = 16384;     // instances per chunk of a VEC-typed group / decoder chunk

inline long long align_up(long long x, long long a) { return (x + a - 1) / a * a; }

struct Plan {
    // encoder phase
    long long xv_in, xv, xq_in, xq, g, c, hs, hx, tsort;
    // module / decoder phase (aliases the encoder regions; everything is stream-ordered)
    long long s0, s1, s2, pl, vp, v01, ats, a0;
    long long total;
    long long mod_bytes; // size of one lane's module scratch (lane l lives at l * mod_bytes)
    int nc_vid;          // instances per VID chunk
    long long R;         // rows per VID chunk
    long long PR;        // rows of the staging plane buffer
};

inline void make_plan(const StairModel& m, const StairBatch& b, Plan* p) {
    const int np = m.precision == STAIR_F32 ? 3 : 1;
    const long long esz = m.precision == STAIR_F32 ? 4 : 2;
    const long long H = m.H, h = m.H / 2, T = b.T, B = b.B;
    int max_count = 1;
    for (int g = 0; g < b.n_groups; ++g) max_count = b.groups[g].count > max_count ? b.groups[g].count : max_count;
    long long nc = ROW_CAP / (T > 0 ? T : 1);
    if (nc < 1) nc = 1;
    if (nc > max_count) nc = max_count;
    p->nc_vid = static_cast<int>(nc);
    p->R = nc * T;
    p->PR = p->R + 2 * nc;
    long long o = 0;
    auto take = [&](long long bytes) { long long r = o; o = align_up(o + bytes, 1024); return r; };
    // encoder regions
    const bool stage_video = !(b.video_dtype == STAIR_BF16 && np == 1 && (m.V % 8) == 0);
    p->xv_in = take(stage_video ? np * B * T * m.V_ld * 2 : 0);
    p->xv = take(B * T * 4 * H * esz);
    p->xq_in = take(static_cast<long long>(np) * b.n_tok * m.text_ld * 2);
    p->xq = take(static_cast<long long>(b.n_tok) * 4 * H * esz);
    p->g = take(2 * B * 4 * h * 4);
    p->c = take(2 * 2 * ((B + 127) / 128 * 128) * h * 4);     // video + text cell states (the fused kernel runs both encoders at once)
    p->hs = take(np * 2 * B * h * 2);
    p->hx = take(lstm_ws_ok(m.precision, static_cast<int>(h), static_cast<int>(B)) ? 2 * lstm_ws_hx_bytes(static_cast<int>(B)) : 0);   // h exchange of the weight-stationary recurrence (video | text)
    p->tsort = take((2LL * B + 1 + b.n_tok) * 4);               // length-sorted text schedule: order [B] | soff [B+1] | tok_src [n_tok]
    const long long enc_total = o;
    // module regions
    o = 0;
    const long long Rv = VEC_CAP < max_count ? VEC_CAP : max_count;
    const long long Rd = VEC_CAP < B ? VEC_CAP : B;
    const long long Rvd = Rv > Rd ? Rv : Rd;
    p->s0 = take(p->R * H * esz);
    p->s1 = take(p->R * H * esz);
    p->s2 = take(p->PR * H * esz);
    p->pl = take(np * p->PR * H * 2);
    p->vp = take(np * Rvd * 3 * H * 2);
    p->v01 = take(Rvd * 2 * H * esz);
    p->ats = take(p->R * (T > 2 ? T : 2) * 4);
    p->a0 = take(p->R * 4);
    p->mod_bytes = o;
    p->total = LANES * o > enc_total ? LANES * o : enc_total;
}

struct Ctx {
    const StairModel& m;
    const StairBatch& b;
    const StairBuffers& buf;
    cudaStream_t st;
    Plan plan;
    char* ws;
    int T, H, h, np, adt, esz;
    StairItabLayout il;
    const int *perm, *out_slot, *arg0, *arg1, *arg2, *pos_q, *span_s, *span_e;
    // dropout (training entry points only; 0 = off): the next GEMM consumes `pending` (set by drop_next)
    float drop_p = 0.0f;
    unsigned long long drop_seed = 0;
    DropSpec pending;
    // training with saved module activations: chunk k of the schedule (groups in order, chunks of a group in order) keeps its scratch
    // (S0/S1/S2/VP/V0/...) at act_base + k * plan.mod_bytes instead of the shared per-lane scratch, so the backward reads it back
    // instead of re-running the chunk's forward
    char* act_base = nullptr;
    // inference entry points only (stair_nmn_forward / stair_op_forward): intermediates nobody reads afterwards may be skipped, e.g. the
    // frame sum of Filter is taken in the epilogue of its second Linear and the [n T, H] activation is never written
    bool inference = false;

    template <typename P> P* at(long long off) const { return reinterpret_cast<P*>(ws + off); }
    const void* W(int id) const { return m.w[id]; }
    const float* Wf(int id) const { return reinterpret_cast<const float*>(m.w[id]); }
    char* act_ptr(void* base, long long elem_off) const { return reinterpret_cast<char*>(base) + elem_off * esz; }
};

// dropout site = id of the Linear whose (activated) output is dropped; row0 = global row id of the GEMM's first row
inline void drop_next(Ctx& c, int site, long long row0) {
    if (c.drop_p > 0.0f) c.pending = make_drop(c.drop_p, c.drop_seed, site, row0);
}
inline DropSpec take_drop(Ctx& c) { const DropSpec d = c.pending; c.pending = DropSpec(); return d; }

// C = act(row_scale * (A_planes . W^T) + bias)
inline int gemm_planes(Ctx& c, const bf16* A, long long lda, long long a_plane_rows, int M, int N, int K, int wid, int bid, int act,
                const float* row_scale, void* C, int cdt, long long ldc) {
    GemmArgs a;
    a.A = A; a.lda = lda; a.a_plane_rows = static_cast<int>(a_plane_rows); a.nplanes = c.np;
    a.W = c.W(wid); a.ldw = align_up(K, 8); a.w_plane_rows = N;
    a.bias = bid >= 0 ? c.Wf(bid) : nullptr; a.row_scale = row_scale;
    a.C = C; a.ldc = ldc; a.out_dtype = cdt; a.M = M; a.N = N; a.K = K; a.act = act;
    a.drop = take_drop(c);
    return launch_gemm(a, c.st);
}

// A = contiguous activation rows [M, K] (act dtype)
inline int gemm_act(Ctx& c, const void* A, int M, int N, int K, int wid, int bid, int act, const float* row_scale, void* C, int cdt, long long ldc) {
    if (c.np == 1) return gemm_planes(c, reinterpret_cast<const bf16*>(A), K, 0, M, N, K, wid, bid, act, row_scale, C, cdt, ldc);
    if (M > c.plan.PR) return STAIR_ERR_CAPACITY;
    bf16* pl = c.at<bf16>(c.plan.pl);
    STAIR_TRY(launch_stage_rows(STAIR_F32, A, K, nullptr, 1, 1, pl, K, M, 3, M, K, c.st));
    return gemm_planes(c, pl, K, M, M, N, K, wid, bid, act, row_scale, C, cdt, ldc);
}

// A = frame rows of n VID slots gathered from the arena (TMA gather when possible, staging copy otherwise)
inline int gemm_vid(Ctx& c, const int* slots, int n, int N, int wid, int bid, int act, const float* row_scale, void* C, int cdt, long long ldc) {
    const int M = n * c.T, K = c.H;
    if (c.np == 1 && gemm_gather_ok(c.T)) {
        GemmArgs a;
        a.A = c.buf.vid; a.lda = K; a.arena_slots = c.buf.vid_slots; a.a_slots = slots; a.slot_rows = c.T;
        a.W = c.W(wid); a.ldw = K; a.w_plane_rows = N; a.bias = bid >= 0 ? c.Wf(bid) : nullptr; a.row_scale = row_scale;
        a.C = C; a.ldc = ldc; a.out_dtype = cdt; a.M = M; a.N = N; a.K = K; a.act = act;
        a.drop = take_drop(c);
        return launch_gemm(a, c.st);
    }
    if (M > c.plan.PR) return STAIR_ERR_CAPACITY;
    bf16* pl = c.at<bf16>(c.plan.pl);
    STAIR_TRY(launch_stage_rows(c.adt, c.buf.vid, K, slots, c.T, c.T, pl, K, M, c.np, M, K, c.st));
    return gemm_planes(c, pl, K, M, M, N, K, wid, bid, act, row_scale, C, cdt, ldc);
}

// A = gathered VEC rows (rps consecutive rows per index)
inline int gemm_vec_rows(Ctx& c, const int* idx, int rps, int n, int N, int wid, int bid, int act, void* C, int cdt, long long ldc) {
    const int M = n * rps, K = c.H;
    if (M > c.plan.PR) return STAIR_ERR_CAPACITY;
    bf16* pl = c.at<bf16>(c.plan.pl);
    STAIR_TRY(launch_stage_rows(c.adt, c.buf.vec, K, idx, rps, 1, pl, K, M, c.np, M, K, c.st));
    return gemm_planes(c, pl, K, M, M, N, K, wid, bid, act, nullptr, C, cdt, ldc);
}

constexpr int MAX_SCHED_GROUPS = 96;     // groups with their own completion event (dependency-driven scheduling); more -> wave scheduling
struct LaneStreams { cudaStream_t side[LANES - 1]; cudaEvent_t fork, join[LANES - 1], done[MAX_SCHED_GROUPS]; bool ok = false; };
// The only CUDA objects the library owns: per calling thread, LANES - 1 side streams + LANES events used to run the independent groups of
// a schedule wave concurrently.  Created by stair_init() (or lazily by the first forward of a thread), destroyed by stair_shutdown().
inline LaneStreams& lane_state() {
    static thread_local LaneStreams ls;
    return ls;
}
inline void lane_streams_destroy() {
    LaneStreams& ls = lane_state();
    if (!ls.ok) return;
    for (int l = 0; l < LANES - 1; ++l) { cudaStreamDestroy(ls.side[l]); cudaEventDestroy(ls.join[l]); }
    for (int g = 0; g < MAX_SCHED_GROUPS; ++g) cudaEventDestroy(ls.done[g]);
    cudaEventDestroy(ls.fork);
    ls.ok = false;
}
inline LaneStreams* lane_streams() {
    LaneStreams& ls = lane_state();
    if (!ls.ok) {
        for (int l = 0; l < LANES - 1; ++l) {
            if (cudaStreamCreateWithFlags(&ls.side[l], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
            if (cudaEventCreateWithFlags(&ls.join[l], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        }
        if (cudaEventCreateWithFlags(&ls.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        for (int g = 0; g < MAX_SCHED_GROUPS; ++g)
            if (cudaEventCreateWithFlags(&ls.done[g], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        ls.ok = true;
    }
    return &ls;
}

extern int g_lanes;      // 1 = everything on the caller's stream; up to LANES
extern int g_timeline;
// Debug marks on the caller's stream (stair_debug_timeline(1)): 0 forward start, 1 after the video projection, 2 after the text projection,
// 3 after the recurrence, 4 after the grouping join, 5 after the module phase, 6 after the decoder (stair_debug_phase_marks reads them)
struct PhaseMarks { cudaEvent_t ev[8]; bool ok = false; };
inline PhaseMarks& phase_marks() { static PhaseMarks p; return p; }
inline void phase_mark(cudaStream_t st, int i) {
    if (!g_timeline) return;
    PhaseMarks& pm = phase_marks();
    if (!pm.ok) { for (int k = 0; k < 8; ++k) cudaEventCreate(&pm.ev[k]); pm.ok = true; }
    cudaEventRecord(pm.ev[i], st);
}

// ---- encoders (module_net.py:147-163) --------------------------------------------------------------------------
inline int run_encoders(Ctx& c, int phases) {
    const StairModel& m = c.m; const StairBatch& b = c.b;
    const int B = b.B, T = c.T, H = c.H, h = c.h;
    const long long rows_v = static_cast<long long>(B) * T;
    if (rows_v * 1 > 0x7fffffffLL || static_cast<long long>(b.n_tok) > 0x7fffffffLL) return STAIR_ERR_CAPACITY;
    float* g = c.at<float>(c.plan.g);
    float* cs = c.at<float>(c.plan.c);
    bf16* hs = c.at<bf16>(c.plan.hs);
    // The fp32 -> bf16 staging of the packed question tokens (HBM-bound, ~24 us at B = 4096) only depends on the inputs: with both encoders
    // requested it runs on a side lane underneath the video projection GEMM (whose 200 KB CTAs leave the SMs' thread slots free).
    bool text_staged = false;
    const bool fused = lstm_fused_ok(m.precision, h) && g_lstm_impl == 0 && c.W(STAIR_W_VENC_WHHI_F) && c.W(STAIR_W_TENC_WHHI_F);
    const bool ws = fused && lstm_ws_ok(m.precision, h, B);
    // Inference: the text recurrence runs over the questions in descending length (a 64-question block stops at its own longest question
    // instead of the batch's).  The sort only permutes which block computes a question; the projection input is staged in that order
    // (gather by tok_src) so that a block's input rows stay contiguous.  Outputs keep the batch's token order.
    const bool tsort = (phases & STAIR_FWD_ENCODE_TEXT) && c.inference && fused && !ws && g_text_sort && text_sort_ok(b.L_max) && b.n_tok > 0;
    const bool tsort_given = tsort && b.q_order && b.q_soff && b.tok_src;      // the caller's schedule (collate); else a device counting sort
    const int* t_order = !tsort ? nullptr : (tsort_given ? b.q_order : c.at<int>(c.plan.tsort));
    const int* t_soff = !tsort ? nullptr : (tsort_given ? b.q_soff : t_order + B);
    const int* t_src = !tsort ? nullptr : (tsort_given ? b.tok_src : t_soff + B + 1);
    auto device_sort = [&](cudaStream_t st) {
        int* o = c.at<int>(c.plan.tsort);
        return launch_text_sort(b.q_off, B, b.L_max, o, o + B, o + 2 * B + 1, st);
    };
    if ((phases & STAIR_FWD_ENCODE_VIDEO) && (phases & STAIR_FWD_ENCODE_TEXT) && c.inference && g_lanes > 2) {
        if (LaneStreams* ls = lane_streams()) {
            if (cudaEventRecord(ls->join[2], c.st) != cudaSuccess || cudaStreamWaitEvent(ls->side[1], ls->join[2], 0) != cudaSuccess) return STAIR_ERR_CUDA;
            if (tsort && !tsort_given) STAIR_TRY(device_sort(ls->side[1]));
            STAIR_TRY(launch_stage_rows(b.question_dtype, b.question, m.text_size, t_src, 1, 1, c.at<bf16>(c.plan.xq_in), m.text_ld, b.n_tok, c.np, b.n_tok,
                                        m.text_size, ls->side[1]));
            if (cudaEventRecord(ls->join[1], ls->side[1]) != cudaSuccess) return STAIR_ERR_CUDA;
            text_staged = true;
        }
    }
    // video input projection: [B*T, V] x [V, 8h] for both directions at once
    if (phases & STAIR_FWD_ENCODE_VIDEO) {
        const bf16* A; long long lda, apr;
        const bool direct = b.video_dtype == STAIR_BF16 && c.np == 1 && (m.V % 8) == 0;
        if (direct) { A = reinterpret_cast<const bf16*>(b.video); lda = m.V; apr = 0; }
        else {
            bf16* in = c.at<bf16>(c.plan.xv_in);
            STAIR_TRY(launch_stage_rows(b.video_dtype, b.video, m.V, nullptr, 1, 1, in, m.V_ld, rows_v, c.np, rows_v, m.V, c.st));
            A = in; lda = m.V_ld; apr = rows_v;
        }
        GemmArgs a;
        a.A = A; a.lda = lda; a.a_plane_rows = static_cast<int>(apr); a.nplanes = c.np; a.W = c.W(STAIR_W_VENC_WIH); a.ldw = m.V_ld;
        a.w_plane_rows = 4 * H; a.bias = c.Wf(STAIR_W_VENC_B); a.C = c.at<void>(c.plan.xv); a.ldc = 4 * H; a.out_dtype = c.adt;
        a.M = static_cast<int>(rows_v); a.N = 4 * H; a.K = m.V;
        STAIR_TRY(launch_gemm(a, c.st));
        phase_mark(c.st, 1);
    }
    if ((phases & STAIR_FWD_ENCODE_VIDEO) && !fused) {
    if (cudaMemsetAsync(cs, 0, sizeof(float) * 2 * B * h, c.st) != cudaSuccess) return STAIR_ERR_CUDA;
    for (int s = 0; s < T; ++s) {
        if (s > 0)
            for (int d = 0; d < 2; ++d)
                STAIR_TRY(gemm_planes(c, hs + static_cast<long long>(d) * B * h, h, 2LL * B, B, 4 * h, h,
                                      d == 0 ? STAIR_W_VENC_WHH_F : STAIR_W_VENC_WHH_R, -1, STAIR_ACT_NONE, nullptr,
                                      g + static_cast<long long>(d) * B * 4 * h, STAIR_F32, 4 * h));
        STAIR_TRY(launch_lstm_cell_video(c.adt, c.at<void>(c.plan.xv), g, cs, hs, c.np, c.adt, c.buf.vid, B, T, h, s, c.st));
    }
    }
    // weight-stationary cluster recurrence (lstm_ws.cu) when eligible (h = 256): one launch per encoder, W_hh resident in shared memory
    const long long nblk128 = (B + 127) / 128;
    char* hx = c.at<char>(c.plan.hx);
    if (ws && (phases & STAIR_FWD_ENCODE_VIDEO))
        STAIR_TRY(launch_lstm_ws(c.at<void>(c.plan.xv), c.buf.vid, nullptr, nullptr, T, c.W(STAIR_W_VENC_WHHI_F), c.W(STAIR_W_VENC_WHHI_R), cs, hx,
                                 B, h, err_flag_ptr(), c.st));
    if (!(phases & STAIR_FWD_ENCODE_TEXT)) {
        if (fused && !ws && (phases & STAIR_FWD_ENCODE_VIDEO))
            return launch_lstm_fused(c.at<void>(c.plan.xv), c.buf.vid, T, c.W(STAIR_W_VENC_WHHI_F), c.W(STAIR_W_VENC_WHHI_R), nullptr,
                                     nullptr, nullptr, nullptr, 0, nullptr, nullptr, cs, B, h, 1, 0, err_flag_ptr(), c.st);
        return STAIR_OK;
    }
    // text input projection over the packed tokens of all questions
    {
        bf16* in = c.at<bf16>(c.plan.xq_in);
        if (text_staged) { if (cudaStreamWaitEvent(c.st, lane_streams()->join[1], 0) != cudaSuccess) return STAIR_ERR_CUDA; }
        else {
            if (tsort && !tsort_given) STAIR_TRY(device_sort(c.st));
            STAIR_TRY(launch_stage_rows(b.question_dtype, b.question, m.text_size, t_src, 1, 1, in, m.text_ld, b.n_tok, c.np, b.n_tok, m.text_size, c.st));
        }
        GemmArgs a;
        a.A = in; a.lda = m.text_ld; a.a_plane_rows = b.n_tok; a.nplanes = c.np; a.W = c.W(STAIR_W_TENC_WIH); a.ldw = m.text_ld;
        a.w_plane_rows = 4 * H; a.bias = c.Wf(STAIR_W_TENC_B); a.C = c.at<void>(c.plan.xq); a.ldc = 4 * H; a.out_dtype = c.adt;
        a.M = b.n_tok; a.N = 4 * H; a.K = m.text_size;
        STAIR_TRY(launch_gemm(a, c.st));
        phase_mark(c.st, 2);
    }
    if (ws)
        return launch_lstm_ws(c.at<void>(c.plan.xq), c.buf.tokfeat, c.buf.qfeat, b.q_off, b.L_max, c.W(STAIR_W_TENC_WHHI_F), c.W(STAIR_W_TENC_WHHI_R),
                              cs + 2 * nblk128 * 128 * h, hx + lstm_ws_hx_bytes(B), B, h, err_flag_ptr(), c.st);
    if (fused)
        return launch_lstm_fused(c.at<void>(c.plan.xv), c.buf.vid, T, c.W(STAIR_W_VENC_WHHI_F), c.W(STAIR_W_VENC_WHHI_R),
                                 c.at<void>(c.plan.xq), c.buf.tokfeat, c.buf.qfeat, b.q_off, b.L_max, c.W(STAIR_W_TENC_WHHI_F),
                                 c.W(STAIR_W_TENC_WHHI_R), cs, B, h, (phases & STAIR_FWD_ENCODE_VIDEO) ? 1 : 0, 1, err_flag_ptr(), c.st, nullptr,
                                 t_order, t_soff);
    if (cudaMemsetAsync(cs, 0, sizeof(float) * 2 * B * h, c.st) != cudaSuccess) return STAIR_ERR_CUDA;
    for (int s = 0; s < b.L_max; ++s) {
        if (s > 0)
            for (int d = 0; d < 2; ++d)
                STAIR_TRY(gemm_planes(c, hs + static_cast<long long>(d) * B * h, h, 2LL * B, B, 4 * h, h,
                                      d == 0 ? STAIR_W_TENC_WHH_F : STAIR_W_TENC_WHH_R, -1, STAIR_ACT_NONE, nullptr,
                                      g + static_cast<long long>(d) * B * 4 * h, STAIR_F32, 4 * h));
        STAIR_TRY(launch_lstm_cell_text(c.adt, c.at<void>(c.plan.xq), g, cs, hs, c.np, c.adt, c.buf.tokfeat, c.buf.qfeat, b.q_off, B, h, s, c.st));
    }
    return STAIR_OK;
}

// ---- one chunk of one group --------------------------------------------------------------------------------------
// p = sorted position of the chunk's first instance, n = instances, ob = first output index, ab = first aux index
inline int run_chunk(Ctx& c, const StairGroup& g, int p, int n, int ob, int ab) {
    const int T = c.T, H = c.H, dt = c.adt;
    const int *a0 = c.arg0 + p, *a1 = c.arg1 + p, *a2 = c.arg2 + p;
    void* S0 = c.at<void>(c.plan.s0); void* S1 = c.at<void>(c.plan.s1); void* S2 = c.at<void>(c.plan.s2);
    bf16* VP = c.at<bf16>(c.plan.vp);
    void* V0 = c.at<void>(c.plan.v01);
    float* att = c.buf.att;
    void* vid_out = c.act_ptr(c.buf.vid, static_cast<long long>(ob) * T * H);
    void* vec_out = c.act_ptr(c.buf.vec, static_cast<long long>(ob) * H);
    const long long pT = static_cast<long long>(p) * T;        // global row id of the chunk's first frame row (dropout masks)
    switch (g.op) {
    case STAIR_OP_WORD:
        return launch_word_embed(dt, c.buf.tokfeat, c.b.q_off, c.pos_q + p, c.span_s + p, c.span_e + p, c.buf.vec, ob, n, H, c.st);
    case STAIR_OP_LOCALIZE: {                                   // modules.py:194-217; variant = K-1
        const int K = g.variant + 1;
        drop_next(c, STAIR_W_LOC_V0_W, pT);
        STAIR_TRY(gemm_vid(c, a0, n, H, STAIR_W_LOC_V0_W, STAIR_W_LOC_V0_B, STAIR_ACT_RELU, nullptr, S0, dt, H));
        STAIR_TRY(gemm_act(c, S0, n * T, H, H, STAIR_W_LOC_V1_W, STAIR_W_LOC_V1_B, STAIR_ACT_NONE, nullptr, S1, dt, H));
        STAIR_TRY(gemm_vec_rows(c, a1, K, n, H, STAIR_W_LOC_K_W, STAIR_W_LOC_K_B, STAIR_ACT_NONE, S2, dt, H));
        return launch_cos_att(dt, S1, S2, K, T, H, att, ob, n, c.st);
    }
    case STAIR_OP_TEMPORAL: {                                   // modules.py:310-327; variant = mode*2 + (K-1)
        const int mode = g.variant >> 1, K = (g.variant & 1) + 1;
        const float* params[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        if (mode > 0) for (int j = 0; j < 6; ++j) params[j] = c.Wf(STAIR_W_TEMP_REL_BEFORE + 6 * (mode - 1) + j);
        STAIR_TRY(launch_temporal_relate(att, a1, K, mode, c.m.conv_k, params, att, ab, n, T, c.st));
        // dense(r[t] * feat[t]) = relu(r[t] * (W feat[t]) + b): the gate is the GEMM's row scale
        drop_next(c, STAIR_W_TEMP_D_W, pT);
        STAIR_TRY(gemm_vid(c, a0, n, H, STAIR_W_TEMP_D_W, STAIR_W_TEMP_D_B, STAIR_ACT_RELU, att + static_cast<long long>(ab) * T, S0, dt, H));
        return launch_layernorm(dt, S0, c.Wf(STAIR_W_TEMP_LN_G), c.Wf(STAIR_W_TEMP_LN_B), vid_out, static_cast<long long>(n) * T, H, c.st);
    }
    case STAIR_OP_FILTER: {                                     // modules.py:361-378; variant 0 repr,1 actions,2 objects,3 relations
        const int w = STAIR_W_FILT_REPR + 4 * g.variant;
        drop_next(c, w, pT);
        STAIR_TRY(gemm_vid(c, a0, n, H, w, w + 1, STAIR_ACT_RELU, nullptr, S0, dt, H));
        // tensor keyword: nn.Softmax() over a size-1 dim makes the attention exactly 1.0 -> plain sum over frames
        if (c.inference && g_fuse_sum && c.np == 1 && c.drop_p == 0.0f && gemm_sum_epilogue_ok(T)) {
            GemmArgs a;                                              // second Linear + ReLU + sum over the T frames of an instance in one kernel
            a.A = S0; a.lda = H; a.W = c.W(w + 2); a.ldw = H; a.w_plane_rows = H; a.bias = c.Wf(w + 3); a.M = n * T; a.N = H; a.K = H;
            a.act = STAIR_ACT_RELU; a.out_dtype = dt; a.sum_out = S2; a.ld_sum = H; a.sum_T = T;
            STAIR_TRY(launch_gemm(a, c.st));
        } else {
            drop_next(c, w + 2, pT);
            STAIR_TRY(gemm_act(c, S0, n * T, H, H, w + 2, w + 3, STAIR_ACT_RELU, nullptr, S1, dt, H));
            STAIR_TRY(launch_sum_T(dt, S1, S2, n, T, H, c.st));
        }
        STAIR_TRY(gemm_act(c, S2, n, H, H, STAIR_W_FILT_D_W, STAIR_W_FILT_D_B, STAIR_ACT_RELU, nullptr, vec_out, dt, H));
        if (g.head) return launch_l2norm(dt, c.buf.vec, ob, c.buf.head_vec, ab, n, H, c.st);
        return STAIR_OK;
    }
    case STAIR_OP_FILTERFRAME: {                                // modules.py:398-414; variant 0 repr,1 relations,2 actions
        const int w = STAIR_W_FF_REPR + 4 * g.variant;
        drop_next(c, w, pT);
        STAIR_TRY(gemm_vid(c, a0, n, H, w, w + 1, STAIR_ACT_RELU, nullptr, S0, dt, H));
        drop_next(c, w + 2, pT);
        STAIR_TRY(gemm_act(c, S0, n * T, H, H, w + 2, w + 3, STAIR_ACT_RELU, nullptr, S1, dt, H));
        const float* gate = nullptr;
        if (g.variant == 0) {
            float* A0 = c.at<float>(c.plan.a0);
            STAIR_TRY(launch_ff_attn(dt, S1, c.buf.vec, a1, c.Wf(STAIR_W_FF_ATT_W), c.Wf(STAIR_W_FF_ATT_B), A0, n, T, H, c.st));
            gate = A0;      // dense(a[t] * x[t]) = relu(a[t] * (W x[t]) + b)
        }
        drop_next(c, STAIR_W_FF_D_W, pT);
        STAIR_TRY(gemm_act(c, S1, n * T, H, H, STAIR_W_FF_D_W, STAIR_W_FF_D_B, STAIR_ACT_RELU, gate, vid_out, dt, H));
        if (g.head)
            return gemm_act(c, vid_out, n * T, c.m.O, H, STAIR_W_FF_HEAD_W, STAIR_W_FF_HEAD_B, STAIR_ACT_NONE, nullptr,
                            c.buf.head_ff + static_cast<long long>(ab) * T * c.m.O, STAIR_F32, c.m.O);
        return STAIR_OK;
    }
    case STAIR_OP_HASITEM:                                      // modules.py:131-138
        drop_next(c, STAIR_W_HAS0_W, pT);
        STAIR_TRY(gemm_vid(c, a0, n, H, STAIR_W_HAS0_W, STAIR_W_HAS0_B, STAIR_ACT_RELU, nullptr, S0, dt, H));
        STAIR_TRY(launch_rowdot_sigmoid(dt, S0, c.Wf(STAIR_W_HAS1_W), c.Wf(STAIR_W_HAS1_B), att, ob, n, T, H, c.st));
        if (c.drop_p > 0.0f)                                     // Sigmoid -> Dropout (modules.py:129): the output map itself is dropped
            return launch_drop_rows(att + static_cast<long long>(ob) * T, static_cast<long long>(n) * T, make_drop(c.drop_p, c.drop_seed, STAIR_W_HAS1_W, pT), c.st);
        return STAIR_OK;
    case STAIR_OP_EXISTSFRAME:                                  // (keyword, feat) modules.py:169-178
        return launch_existsframe(dt, c.buf.vid, a1, c.buf.vec, a0, att, ob, n, T, H, c.st);
    case STAIR_OP_RELATE:                                       // variant 0 forward, 1 backward; modules.py:423-435
        // variants 2 / 3: the operand is a 2-D [1, T] map -> the reference adds the constant beta[0]: softmax_T(attn), sign 0
        return launch_relate(att, a0, c.Wf(STAIR_W_REL_BETA), g.variant >= 2 ? 0 : (g.variant == 0 ? 1 : -1), att, ob, n, T, c.st);
    case STAIR_OP_ATTNVIDEO:
        return launch_attnvideo(dt, c.buf.vid, a0, att, a1, ob, n, T, H, c.st);
    case STAIR_OP_AND:
    case STAIR_OP_XORFRAME: {                                   // variant 0: VEC rows, 1: [T] maps, 2: [2,T] maps
        const int op = g.op == STAIR_OP_AND ? STAIR_BIN_MIN : STAIR_BIN_ABSDIFF;
        if (g.variant == 0) return launch_binary(dt, c.buf.vec, a0, a1, ob, H, H, op, n, c.st);
        return launch_binary(STAIR_F32, att, a0, a1, ob, T, g.variant * T, op, n, c.st);
    }
    case STAIR_OP_CHOOSE:
        return launch_choose(dt, c.buf.vec, a0, a1, a2, ob, n, H, c.st);
    case STAIR_OP_ARRAY2:
        return launch_array2(dt, c.buf.vec, a0, a1, ob, n, H, c.st);
    case STAIR_OP_COMPARE:
    case STAIR_OP_EQUALS: {                                     // modules.py:15-37
        const bool eq = g.op == STAIR_OP_EQUALS;
        STAIR_TRY(launch_concat_vec(dt, c.buf.vec, a0, a1, STAIR_CAT_PAIR, VP, n, c.np, n, H, c.st));
        STAIR_TRY(gemm_planes(c, VP, 2 * H, n, n, H, 2 * H, eq ? STAIR_W_EQUALS_W : STAIR_W_COMPARE_W, eq ? STAIR_W_EQUALS_B : STAIR_W_COMPARE_B,
                              STAIR_ACT_RELU, nullptr, vec_out, dt, H));
        if (eq && g.head)
            return launch_small_head(dt, c.buf.vec, ob, c.Wf(STAIR_W_EQUALS_HEAD_W), c.Wf(STAIR_W_EQUALS_HEAD_B), 1, c.buf.head_small, ab, n, H, c.st);
        return STAIR_OK;
    }
    case STAIR_OP_XOR:                                          // modules.py:59-72
        STAIR_TRY(launch_concat_vec(dt, c.buf.vec, a0, a1, STAIR_CAT_XOR, VP, n, c.np, n, H, c.st));
        STAIR_TRY(gemm_planes(c, VP, 3 * H, n, n, H, 3 * H, STAIR_W_XOR_W, STAIR_W_XOR_B, STAIR_ACT_RELU, nullptr, vec_out, dt, H));
        if (g.head) return launch_small_head(dt, c.buf.vec, ob, c.Wf(STAIR_W_XOR_HEAD_W), c.Wf(STAIR_W_XOR_HEAD_B), 2, c.buf.head_small, ab, n, H, c.st);
        return STAIR_OK;
    case STAIR_OP_EXISTS:                                       // (keyword, feat) modules.py:141-159
        STAIR_TRY(launch_concat_vec(dt, c.buf.vec, a0, a1, STAIR_CAT_EXISTS, VP, n, c.np, n, H, c.st));
        drop_next(c, STAIR_W_EXISTS0_W, p);
        STAIR_TRY(gemm_planes(c, VP, 3 * H, n, n, H, 3 * H, STAIR_W_EXISTS0_W, STAIR_W_EXISTS0_B, STAIR_ACT_RELU, nullptr, V0, dt, H));
        drop_next(c, STAIR_W_EXISTS1_W, p);
        STAIR_TRY(gemm_act(c, V0, n, H, H, STAIR_W_EXISTS1_W, STAIR_W_EXISTS1_B, STAIR_ACT_RELU, nullptr, vec_out, dt, H));
        if (g.head) return launch_small_head(dt, c.buf.vec, ob, c.Wf(STAIR_W_EXISTS_HEAD_W), c.Wf(STAIR_W_EXISTS_HEAD_B), 2, c.buf.head_small, ab, n, H, c.st);
        return STAIR_OK;
    case STAIR_OP_TOACTION:                                     // (action, keyword) modules.py:102-120
        STAIR_TRY(launch_concat_vec(dt, c.buf.vec, a0, a1, STAIR_CAT_PAIR, VP, n, c.np, n, H, c.st));
        drop_next(c, STAIR_W_TOACT0_W, p);
        STAIR_TRY(gemm_planes(c, VP, 2 * H, n, n, H, 2 * H, STAIR_W_TOACT0_W, STAIR_W_TOACT0_B, STAIR_ACT_RELU, nullptr, V0, dt, H));
        STAIR_TRY(gemm_act(c, V0, n, H, H, STAIR_W_TOACT1_W, STAIR_W_TOACT1_B, STAIR_ACT_RELU, nullptr, vec_out, dt, H));
        if (g.head) return launch_l2norm(dt, c.buf.vec, ob, c.buf.head_vec, ab, n, H, c.st);
        return STAIR_OK;
    case STAIR_OP_SUPERLATIVE: {                                // (mode, actions, feat) modules.py:233-248
        // variant = is_min + 2*kind ; kind 0: one VEC action, 1: Array2 (two rows), 2: [T,H] frame features as T actions
        const int is_min = g.variant & 1, kind = g.variant >> 1;
        const int K = kind == 0 ? 1 : (kind == 1 ? 2 : T);
        drop_next(c, STAIR_W_LOC_V0_W, pT);
        STAIR_TRY(gemm_vid(c, a1, n, H, STAIR_W_LOC_V0_W, STAIR_W_LOC_V0_B, STAIR_ACT_RELU, nullptr, S0, dt, H));
        STAIR_TRY(gemm_act(c, S0, n * T, H, H, STAIR_W_LOC_V1_W, STAIR_W_LOC_V1_B, STAIR_ACT_NONE, nullptr, S1, dt, H));
        if (kind == 2) STAIR_TRY(gemm_vid(c, a0, n, H, STAIR_W_LOC_K_W, STAIR_W_LOC_K_B, STAIR_ACT_NONE, nullptr, S2, dt, H));
        else STAIR_TRY(gemm_vec_rows(c, a0, K, n, H, STAIR_W_LOC_K_W, STAIR_W_LOC_K_B, STAIR_ACT_NONE, S2, dt, H));
        float* ats = c.at<float>(c.plan.ats);
        STAIR_TRY(launch_cos_att(dt, S1, S2, K, T, H, ats, 0, n, c.st));
        STAIR_TRY(launch_super_mix(dt, ats, K, T, H, is_min, kind == 2 ? c.buf.vid : c.buf.vec, a0, kind == 2 ? T : 1, V0, n, c.st));
        STAIR_TRY(gemm_act(c, V0, n, H, H, STAIR_W_SUP_D_W, STAIR_W_SUP_D_B, STAIR_ACT_RELU, nullptr, vec_out, dt, H));
        if (g.head) return launch_l2norm(dt, c.buf.vec, ob, c.buf.head_vec, ab, n, H, c.st);
        return STAIR_OK;
    }
    default:
        return STAIR_ERR_LAYOUT;
    }
}

inline bool op_is_vid_sized(int op) {
    switch (op) {
    case STAIR_OP_LOCALIZE: case STAIR_OP_TEMPORAL: case STAIR_OP_FILTER: case STAIR_OP_FILTERFRAME: case STAIR_OP_HASITEM:
    case STAIR_OP_SUPERLATIVE: case STAIR_OP_ATTNVIDEO: case STAIR_OP_EXISTSFRAME:
        return true;
    default:
        return false;
    }
}

inline int group_cap(const Ctx& c, const StairGroup& g) { return op_is_vid_sized(g.op) ? c.plan.nc_vid : static_cast<int>(VEC_CAP); }
inline long long group_chunks(const Ctx& c, const StairGroup& g) { const int cap = group_cap(c, g); return (g.count + cap - 1) / cap; }
// index of the first chunk of group gi in schedule order (saved-activation slots)
inline long long chunk_base(const Ctx& c, int gi) {
    long long k = 0;
    for (int g = 0; g < gi; ++g) k += group_chunks(c, c.b.groups[g]);
    return k;
}
inline long long total_chunks(const Ctx& c) { return chunk_base(c, c.b.n_groups); }

inline int run_group(Ctx& c, const StairGroup& g, int gi) {
    const int cap = group_cap(c, g);
    long long k = c.act_base ? chunk_base(c, gi) : 0;
    char* const ws0 = c.ws;
    for (int done = 0; done < g.count; done += cap, ++k) {
        const int n = g.count - done < cap ? g.count - done : cap;
        if (c.act_base) c.ws = c.act_base + k * c.plan.mod_bytes;
        const int rc = run_chunk(c, g, g.node_off + done, n, g.out_base + done * g.out_mult, g.aux_base >= 0 ? g.aux_base + done : -1);
        c.ws = ws0;
        if (rc != STAIR_OK) return rc;
    }
    return STAIR_OK;
}

// rough cost of a group for the lane assignment (launch-latency floor per kernel + work)
inline long long group_cost(const StairGroup& g, int T) {
    int launches;
    switch (g.op) {
    case STAIR_OP_LOCALIZE: launches = 5; break;
    case STAIR_OP_SUPERLATIVE: launches = 7; break;
    case STAIR_OP_FILTER: case STAIR_OP_FILTERFRAME: launches = 4; break;
    case STAIR_OP_TEMPORAL: case STAIR_OP_EXISTS: case STAIR_OP_TOACTION: launches = 3; break;
    case STAIR_OP_HASITEM: case STAIR_OP_COMPARE: case STAIR_OP_EQUALS: case STAIR_OP_XOR: launches = 2; break;
    default: launches = 1;
    }
    return launches * 1000LL + static_cast<long long>(g.count) * (op_is_vid_sized(g.op) ? T : 1) / 16;
}

// Debug timeline of the dependency-driven module phase (stair_debug_timeline): timing events around every group, read back by
// stair_debug_timeline_read.  Off by default; the events perturb the schedule by a fraction of a microsecond per group.
struct Timeline { cudaEvent_t origin, t0[MAX_SCHED_GROUPS], t1[MAX_SCHED_GROUPS]; int lane[MAX_SCHED_GROUPS], op[MAX_SCHED_GROUPS], count[MAX_SCHED_GROUPS], variant[MAX_SCHED_GROUPS]; int n = 0; bool ok = false; };
inline Timeline& timeline_state() { static Timeline t; return t; }
extern int g_timeline;

extern int g_dep_sched;  // 1 = schedule the module groups by data dependency when the batch carries group_deps

// Dependency-driven module phase.  Wave scheduling (run_modules below) joins all lanes after every schedule wave, so a wave lasts as long
// as its slowest lane and the lanes idle in between; but a group only needs the groups whose outputs it reads.  Here every group gets
// a completion event; a group is placed on the lane of its latest-finishing producer when that producer is the lane's tail (the chain
// continues in stream order, no event wait) and on the least-loaded lane otherwise, waiting only for the events of its own producers.
// The critical path becomes the longest dependency chain of the layout mix instead of the sum of the per-wave maxima.
inline int run_modules_dep(Ctx& c, LaneStreams* ls) {
    const int ng = c.b.n_groups;
    const int lanes = g_lanes < LANES ? g_lanes : LANES;
    int lane_of[MAX_SCHED_GROUPS], tail[LANES];
    long long lane_end[LANES], finish[MAX_SCHED_GROUPS];
    bool used[LANES];
    for (int l = 0; l < lanes; ++l) { tail[l] = -1; lane_end[l] = 0; used[l] = false; }
    Timeline* tl = nullptr;
    if (g_timeline) {
        tl = &timeline_state();
        if (!tl->ok) {
            cudaEventCreate(&tl->origin);
            for (int g = 0; g < MAX_SCHED_GROUPS; ++g) { cudaEventCreate(&tl->t0[g]); cudaEventCreate(&tl->t1[g]); }
            tl->ok = true;
        }
        tl->n = ng;
        cudaEventRecord(tl->origin, c.st);                         // before the fork: every lane starts after it
    }
    if (cudaEventRecord(ls->fork, c.st) != cudaSuccess) return STAIR_ERR_CUDA;
    for (int g = 0; g < ng; ++g) {
        const int* deps = c.b.group_deps + static_cast<long long>(g) * STAIR_MAX_GROUP_DEPS;
        const bool all = deps[0] == -2;
        long long ready = 0;
        int latest = -1;
        for (int d = 0; d < (all ? g : STAIR_MAX_GROUP_DEPS); ++d) {
            const int pg = all ? d : deps[d];
            if (pg < 0) break;
            if (pg >= g) return STAIR_ERR_LAYOUT;
            if (finish[pg] >= ready) { ready = finish[pg]; latest = pg; }
        }
        // lane choice: continue the producer's chain when it is its lane's tail, else the lane that is free first
        int lane = -1;
        if (latest >= 0 && tail[lane_of[latest]] == latest) lane = lane_of[latest];
        else {
            lane = 0;
            for (int l = 1; l < lanes; ++l) if (lane_end[l] < lane_end[lane]) lane = l;
        }
        Ctx lc = c;
        cudaStream_t st = lane == 0 ? c.st : ls->side[lane - 1];
        if (lane > 0) { lc.st = st; lc.ws = c.ws + static_cast<long long>(lane) * c.plan.mod_bytes; }
        if (lane > 0 && !used[lane]) {
            if (cudaStreamWaitEvent(st, ls->fork, 0) != cudaSuccess) return STAIR_ERR_CUDA;
            used[lane] = true;
        }
        for (int d = 0; d < (all ? g : STAIR_MAX_GROUP_DEPS); ++d) {
            const int pg = all ? d : deps[d];
            if (pg < 0) break;
            if (lane_of[pg] != lane && cudaStreamWaitEvent(st, ls->done[pg], 0) != cudaSuccess) return STAIR_ERR_CUDA;
        }
        if (tl) { cudaEventRecord(tl->t0[g], st); tl->lane[g] = lane; tl->op[g] = c.b.groups[g].op; tl->count[g] = c.b.groups[g].count; tl->variant[g] = c.b.groups[g].variant; }
        STAIR_TRY(run_group(lc, c.b.groups[g], g));
        if (tl) cudaEventRecord(tl->t1[g], st);
        if (cudaEventRecord(ls->done[g], st) != cudaSuccess) return STAIR_ERR_CUDA;
        lane_of[g] = lane;
        tail[lane] = g;
        const long long start = ready > lane_end[lane] ? ready : lane_end[lane];
        finish[g] = lane_end[lane] = start + group_cost(c.b.groups[g], c.T);
    }
    for (int l = 1; l < lanes; ++l)
        if (used[l]) {
            if (cudaEventRecord(ls->join[l - 1], ls->side[l - 1]) != cudaSuccess) return STAIR_ERR_CUDA;
            if (cudaStreamWaitEvent(c.st, ls->join[l - 1], 0) != cudaSuccess) return STAIR_ERR_CUDA;
        }
    return STAIR_OK;
}

// Groups are listed in schedule order (wave-major).  The groups of one wave are independent (they read earlier waves and write
// disjoint arena ranges), so they are spread over up to LANES streams forked from / joined back into the caller's stream; every
// lane has its own scratch.  At B = 4096 the module kernels are launch/latency-bound (10-20 us each on a mostly idle GPU):
// running a wave's groups side by side hides that latency.
inline int run_modules(Ctx& c) {
    const int ng = c.b.n_groups;
    LaneStreams* ls = g_lanes > 1 ? lane_streams() : nullptr;
    if (ls && g_dep_sched && c.b.group_deps && ng <= MAX_SCHED_GROUPS) return run_modules_dep(c, ls);
    int gi = 0;
    while (gi < ng) {
        int gj = gi + 1;
        while (gj < ng && c.b.groups[gj].level == c.b.groups[gi].level) ++gj;
        const int nw = gj - gi;
        if (nw == 1 || !ls) {
            for (int g = gi; g < gj; ++g) STAIR_TRY(run_group(c, c.b.groups[g], g));
            gi = gj;
            continue;
        }
        const int lanes = nw < g_lanes ? nw : g_lanes;
        // longest-processing-time-first assignment of the wave's groups to lanes
        int order[64], lane_of[64];
        long long load[LANES] = {0};
        const int nwc = nw < 64 ? nw : 64;
        for (int k = 0; k < nwc; ++k) order[k] = gi + k;
        for (int a = 1; a < nwc; ++a)
            for (int b2 = a; b2 > 0 && group_cost(c.b.groups[order[b2]], c.T) > group_cost(c.b.groups[order[b2 - 1]], c.T); --b2) {
                const int t = order[b2]; order[b2] = order[b2 - 1]; order[b2 - 1] = t;
            }
        for (int k = 0; k < nwc; ++k) {
            int best = 0;
            for (int l = 1; l < lanes; ++l) if (load[l] < load[best]) best = l;
            lane_of[k] = best;
            load[best] += group_cost(c.b.groups[order[k]], c.T);
        }
        if (cudaEventRecord(ls->fork, c.st) != cudaSuccess) return STAIR_ERR_CUDA;
        for (int l = 1; l < lanes; ++l)
            if (cudaStreamWaitEvent(ls->side[l - 1], ls->fork, 0) != cudaSuccess) return STAIR_ERR_CUDA;
        for (int k = 0; k < nwc; ++k) {
            Ctx lc = c;
            const int l = lane_of[k];
            if (l > 0) { lc.st = ls->side[l - 1]; lc.ws = c.ws + static_cast<long long>(l) * c.plan.mod_bytes; }
            STAIR_TRY(run_group(lc, c.b.groups[order[k]], order[k]));
        }
        for (int g = gi + nwc; g < gj; ++g) STAIR_TRY(run_group(c, c.b.groups[g], g));      // (more than 64 groups in a wave: the rest on the main lane)
        for (int l = 1; l < lanes; ++l) {
            if (cudaEventRecord(ls->join[l - 1], ls->side[l - 1]) != cudaSuccess) return STAIR_ERR_CUDA;
            if (cudaStreamWaitEvent(c.st, ls->join[l - 1], 0) != cudaSuccess) return STAIR_ERR_CUDA;
        }
        gi = gj;
    }
    return STAIR_OK;
}

// decoder (module_net.py:135-138) + argmax (train_module.py:252)
inline int run_decoder(Ctx& c) {
    const int B = c.b.B, H = c.H, A = c.m.A;
    bf16* VP = c.at<bf16>(c.plan.vp);
    void* D0 = c.at<void>(c.plan.v01);
    for (int done = 0; done < B; done += static_cast<int>(VEC_CAP)) {
        const int n = B - done < VEC_CAP ? B - done : static_cast<int>(VEC_CAP);
        STAIR_TRY(launch_decoder_concat(c.adt, c.buf.vec, c.b.root_node + done, c.out_slot,
                                        c.act_ptr(c.buf.qfeat, static_cast<long long>(done) * H), VP, n, c.np, n, H, c.st));
        drop_next(c, STAIR_W_DEC0_W, done);
        STAIR_TRY(gemm_planes(c, VP, 2 * H, n, n, 2 * H, 2 * H, STAIR_W_DEC0_W, STAIR_W_DEC0_B, STAIR_ACT_RELU, nullptr, D0, c.adt, 2 * H));
        if (c.np == 1) {
            STAIR_TRY(gemm_planes(c, reinterpret_cast<const bf16*>(D0), 2 * H, 0, n, A, 2 * H, STAIR_W_DEC1_W, STAIR_W_DEC1_B, STAIR_ACT_NONE, nullptr,
                                  c.buf.logits + static_cast<long long>(done) * A, STAIR_F32, A));
        } else {
            STAIR_TRY(launch_stage_rows(STAIR_F32, D0, 2 * H, nullptr, 1, 1, VP, 2 * H, n, 3, n, 2 * H, c.st));
            STAIR_TRY(gemm_planes(c, VP, 2 * H, n, n, A, 2 * H, STAIR_W_DEC1_W, STAIR_W_DEC1_B, STAIR_ACT_NONE, nullptr,
                                  c.buf.logits + static_cast<long long>(done) * A, STAIR_F32, A));
        }
    }
    return launch_argmax(c.buf.logits, c.buf.answers, B, A, c.st);
}


}  // namespace ex
}  // namespace stair
