// Training step of the batched interpreter: forward with encoder history, intermediate-supervision losses
// (train_module.py:33-194, window logic :341-406) and the backward pass of every operator, encoder and the decoder.
//
// Design: module intermediates are NOT stored by the forward; the backward walks the groups in reverse schedule order and,
// per group chunk, re-runs the group's forward (ex::run_chunk, a few GEMMs) into the forward scratch and then back-propagates
// through it.  Every nn.Linear backward is two tcgen05 GEMMs (dX = dZ.W via the transposed weight copy, dW += dZ^T.X with
// fp32 accumulation in the epilogue) fed by a staging kernel (ReLU mask, row-scale, bias gradient) and two bf16 transposes.
// The LSTMs use BPTT over the saved gate / cell / state history; their weight gradients are two big GEMMs after the loop.
// Gradients are fp32; GEMM operands are bf16 (1 plane) or bf16x3 planes (strict fp32 mode), like the forward.
#include "exec_core.cuh"
#include "train_kernels.cuh"

namespace stair {
namespace {

using namespace ex;

extern int g_bwd_lanes;
int g_dw_impl = 0;      // 0 = MN-major weight-gradient GEMM (product); 1 = transposed copies + K-major GEMM (comparison)
int g_bptt_impl = 0;    // 0 = persistent fused BPTT kernel when eligible (lstm_bptt.cu; product); 1 = per-step cell kernel + recurrent GEMMs

struct Bump {
    char* base = nullptr;
    long long cap = 0, off = 0, peak = 0;
    bool dry = false, overflow = false;
    template <typename P> P* take(long long count) {
        off = align_up(off, 256);
        const long long o = off;
        off += count * static_cast<long long>(sizeof(P));
        if (off > peak) peak = off;
        if (!dry && off > cap) overflow = true;
        return reinterpret_cast<P*>(base + o);
    }
};

struct BCtx {
    Ctx& c;
    const StairTrain& tr;
    Bump ws;
    bool dry;
};

#define RUN(expr) do { if (!b.dry) { if (b.ws.overflow) return STAIR_ERR_CAPACITY; int rc__ = (expr); if (rc__ != STAIR_OK) return rc__; } } while (0)

inline float* G(BCtx& b, int id) { return id >= 0 ? b.tr.grad[id] : nullptr; }
inline float DS(BCtx& b) { return b.c.drop_p > 0.0f ? 1.0f / (1.0f - b.c.drop_p) : 1.0f; }     // scale of kept elements at a dropout site

// ---- saved encoder history ------------------------------------------------------------------------------------------------
struct EncSaved { long long gates, c, hs; int S; };          // byte offsets into StairTrain.saved
struct SavedLayout { EncSaved enc[2]; long long total; };

SavedLayout saved_layout(const StairModel& m, const StairBatch& b) {
    const long long np = m.precision == STAIR_F32 ? 3 : 1, B = b.B, h = m.H / 2;
    const long long Bp = align_up(B, 128);       // the blocked history of the fused recurrence pads the batch to whole 128-row CTAs
    SavedLayout L;
    long long o = 0;
    for (int e = 0; e < 2; ++e) {
        const long long S = e == 0 ? b.T : b.L_max;
        L.enc[e].S = static_cast<int>(S);
        L.enc[e].gates = o; o = align_up(o + S * 2 * Bp * 4 * h * 4, 1024);
        L.enc[e].c = o; o = align_up(o + S * 2 * Bp * h * 4, 1024);
        L.enc[e].hs = o; o = align_up(o + np * 2 * (S + 1) * B * h * 2, 1024);
    }
    L.total = o;
    return L;
}

// ---- staging helpers ------------------------------------------------------------------------------------------------------------
bf16* stage_act(BCtx& b, const void* A, int sdt, int M, int K, int* rc) {
    Ctx& c = b.c;
    // bf16 storage, one plane, 16-byte row pitch: the activation rows already ARE the GEMM operand
    if (c.np == 1 && sdt == STAIR_BF16 && (K % 8) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0)
        return reinterpret_cast<bf16*>(const_cast<void*>(A));
    const long long kld = align_up(K, 8);
    bf16* p = b.ws.take<bf16>(static_cast<long long>(c.np) * M * kld);
    if (!b.dry && !b.ws.overflow) *rc = launch_stage_rows(sdt, A, K, nullptr, 1, 1, p, kld, M, c.np, M, K, c.st);
    return p;
}
bf16* stage_gather(BCtx& b, const void* base, int sdt, const int* idx, int rps, int unit, int n, int K, int* rc) {
    Ctx& c = b.c;
    const long long M = static_cast<long long>(n) * rps;
    bf16* p = b.ws.take<bf16>(static_cast<long long>(c.np) * M * K);
    if (!b.dry && !b.ws.overflow) *rc = launch_stage_rows(sdt, base, K, idx, rps, unit, p, K, M, c.np, M, K, c.st);
    return p;
}

// Linear backward.  dY fp32 [M,N]; Y = post-activation output (ReLU mask) or null; rs = row scale or null;
// Xp = layer input as bf16 planes [np][x_plane_rows, K_ld].  dX (fp32 [M,K]) receives dZ.W WITHOUT the row scale.
// yscale = 1/(1-p) when Y went through Dropout after its ReLU (dropped elements are exactly 0 in Y, so Y > 0 is the joint mask).
int linear_bwd(BCtx& b, const float* dY, long long ld_dy, const void* Y, long long ld_y, const float* rs, const bf16* Xp,
               long long x_plane_rows, int M, int N, int K, int wid, int bid, float* dX, float yscale = 1.0f) {
    Ctx& c = b.c;
    if (M <= 0) return STAIR_OK;
    const long long mark = b.ws.off;
    const long long n_ld = align_up(N, 8), k_ld = align_up(K, 8), m_ld = align_up(M, 8);
    bf16* dZ = b.ws.take<bf16>(c.np * static_cast<long long>(M) * n_ld);
    bf16* dZs = rs ? b.ws.take<bf16>(c.np * static_cast<long long>(M) * n_ld) : dZ;
    RUN(launch_dz_prep(c.adt, dY, ld_dy, Y, ld_y, rs, dZ, dZs, n_ld, M, c.np, G(b, bid), M, N, Y ? yscale : 1.0f, c.st));
    float* gW = G(b, wid);
    if (gW && g_dw_impl == 0) {
        // dW[N,K] += dZs^T . X over the M rows: both operands in place as MN-major tiles (no transposed copies)
        GemmArgs a;
        a.A = dZs; a.lda = n_ld; a.a_plane_rows = M; a.W = Xp; a.ldw = k_ld; a.w_plane_rows = static_cast<int>(x_plane_rows); a.nplanes = c.np;
        a.C = gW; a.ldc = K; a.out_dtype = STAIR_F32; a.M = N; a.N = K; a.K = M; a.accumulate = 1; a.mn_major = 1;
        a.atomic_acc = g_bwd_lanes > 1;                          // groups sharing a weight may run concurrently
        RUN(launch_gemm(a, c.st));
    } else if (gW) {
        bf16* dZt = b.ws.take<bf16>(c.np * static_cast<long long>(N) * m_ld);
        bf16* Xt = b.ws.take<bf16>(c.np * static_cast<long long>(K) * m_ld);
        RUN(launch_transpose_planes(dZs, n_ld, M, dZt, m_ld, N, c.np, M, N, c.st));
        RUN(launch_transpose_planes(Xp, k_ld, x_plane_rows, Xt, m_ld, K, c.np, M, K, c.st));
        GemmArgs a;
        a.A = dZt; a.lda = m_ld; a.a_plane_rows = N; a.W = Xt; a.ldw = m_ld; a.w_plane_rows = K; a.nplanes = c.np;
        a.C = gW; a.ldc = K; a.out_dtype = STAIR_F32; a.M = N; a.N = K; a.K = M; a.accumulate = 1;
        a.atomic_acc = g_bwd_lanes > 1;
        RUN(launch_gemm(a, c.st));
    }
    if (dX) {
        if (!c.m.wt[wid]) return STAIR_ERR_ARG;
        GemmArgs a;
        a.A = dZ; a.lda = n_ld; a.a_plane_rows = M; a.W = c.m.wt[wid]; a.ldw = n_ld; a.w_plane_rows = K; a.nplanes = c.np;
        a.C = dX; a.ldc = K; a.out_dtype = STAIR_F32; a.M = M; a.N = K; a.K = N;
        RUN(launch_gemm(a, c.st));
    }
    b.ws.off = mark;
    return STAIR_OK;
}

#define STAGE_ACT(var, ptr, sdt, M, K) bf16* var; { int rc__ = STAIR_OK; var = stage_act(b, ptr, sdt, M, K, &rc__); if (rc__) return rc__; }
#define STAGE_GATHER(var, base, sdt, idx, rps, unit, n, K) bf16* var; { int rc__ = STAIR_OK; var = stage_gather(b, base, sdt, idx, rps, unit, n, K, &rc__); if (rc__) return rc__; }

// x = relu(W2 relu(W1 feat + b1) + b2) backward given dx (grad wrt S1); S0/S1 are the recomputed activations
int mlp2_bwd(BCtx& b, float* dx, const void* S0, const void* S1, const int* feat_slots, int n, int w) {
    Ctx& c = b.c;
    const int T = c.T, H = c.H, M = n * T;
    const long long mark = b.ws.off;
    float* dS0 = b.ws.take<float>(static_cast<long long>(M) * H);
    STAGE_ACT(s0p, S0, c.adt, M, H);
    STAIR_TRY(linear_bwd(b, dx, H, S1, H, nullptr, s0p, M, M, H, H, w + 2, w + 3, dS0, DS(b)));
    float* dF = b.ws.take<float>(static_cast<long long>(M) * H);
    STAGE_GATHER(fp, c.buf.vid, c.adt, feat_slots, T, T, n, H);
    STAIR_TRY(linear_bwd(b, dS0, H, S0, H, nullptr, fp, M, M, H, H, w, w + 1, dF, DS(b)));
    RUN(launch_scatter_add_rows(dF, feat_slots, T, T, b.tr.dvid, M, H, c.st));
    b.ws.off = mark;
    return STAIR_OK;
}

// Localize body backward (shared with Superlative): df/dk from the cosine map -> video_linear, keyword_linear
int localize_bwd(BCtx& b, const float* datt, long long att_base, const void* S0, const void* S1, const void* S2, const int* feat_slots,
                 const void* kw_base, float* dkw_base, const int* kw_idx, int kw_rps, int kw_unit, int K, int n) {
    Ctx& c = b.c;
    const int T = c.T, H = c.H, M = n * T;
    const long long mark = b.ws.off;
    float* df = b.ws.take<float>(static_cast<long long>(M) * H);
    float* dk = b.ws.take<float>(static_cast<long long>(n) * K * H);
    RUN(launch_cos_att_bwd(c.adt, S1, S2, K, T, H, datt, att_base, df, dk, n, c.st));
    float* dS0 = b.ws.take<float>(static_cast<long long>(M) * H);
    STAGE_ACT(s0p, S0, c.adt, M, H);
    STAIR_TRY(linear_bwd(b, df, H, nullptr, 0, nullptr, s0p, M, M, H, H, STAIR_W_LOC_V1_W, STAIR_W_LOC_V1_B, dS0));
    float* dF = b.ws.take<float>(static_cast<long long>(M) * H);
    STAGE_GATHER(fp, c.buf.vid, c.adt, feat_slots, T, T, n, H);
    STAIR_TRY(linear_bwd(b, dS0, H, S0, H, nullptr, fp, M, M, H, H, STAIR_W_LOC_V0_W, STAIR_W_LOC_V0_B, dF, DS(b)));
    RUN(launch_scatter_add_rows(dF, feat_slots, T, T, b.tr.dvid, M, H, c.st));
    float* dKW = b.ws.take<float>(static_cast<long long>(n) * K * H);
    STAGE_GATHER(kp, kw_base, c.adt, kw_idx, kw_rps, kw_unit, n, H);
    STAIR_TRY(linear_bwd(b, dk, H, nullptr, 0, nullptr, kp, static_cast<long long>(n) * K, n * K, H, H, STAIR_W_LOC_K_W, STAIR_W_LOC_K_B, dKW));
    RUN(launch_scatter_add_rows(dKW, kw_idx, kw_rps, kw_unit, dkw_base, static_cast<long long>(n) * K, H, c.st));
    b.ws.off = mark;
    return STAIR_OK;
}

// backward of one group chunk; the chunk's forward has just been recomputed into the forward scratch (c.plan)
int chunk_bwd(BCtx& b, const StairGroup& g, int p, int n, int ob, int ab) {
    Ctx& c = b.c;
    const StairTrain& tr = b.tr;
    const int T = c.T, H = c.H, dt = c.adt;
    const int *a0 = c.arg0 + p, *a1 = c.arg1 + p, *a2 = c.arg2 + p;
    void* S0 = c.at<void>(c.plan.s0); void* S1 = c.at<void>(c.plan.s1); void* S2 = c.at<void>(c.plan.s2);
    bf16* VP = c.at<bf16>(c.plan.vp);
    void* V0 = c.at<void>(c.plan.v01);
    float* att = c.buf.att;
    void* vid_out = c.act_ptr(c.buf.vid, static_cast<long long>(ob) * T * H);
    void* vec_out = c.act_ptr(c.buf.vec, static_cast<long long>(ob) * H);
    float* dvid_out = tr.dvid + static_cast<long long>(ob) * T * H;
    float* dvec_out = tr.dvec + static_cast<long long>(ob) * H;
    float* datt_out = tr.datt + static_cast<long long>(ob) * T;
    const long long mark = b.ws.off;
    const int M = n * T;
    switch (g.op) {
    case STAIR_OP_WORD:
        RUN(launch_word_embed_bwd(tr.dvec, ob, c.b.q_off, c.pos_q + p, c.span_s + p, c.span_e + p, tr.dtokfeat, n, H, c.st));
        break;
    case STAIR_OP_LOCALIZE:
        STAIR_TRY(localize_bwd(b, tr.datt, ob, S0, S1, S2, a0, c.buf.vec, tr.dvec, a1, g.variant + 1, 1, g.variant + 1, n));
        break;
    case STAIR_OP_TEMPORAL: {
        const int mode = g.variant >> 1, K = (g.variant & 1) + 1;
        const float* r = att + static_cast<long long>(ab) * T;
        float* dr = tr.datt + static_cast<long long>(ab) * T;
        float* dS0 = b.ws.take<float>(static_cast<long long>(M) * H);
        RUN(launch_layernorm_bwd(dt, dvid_out, S0, c.Wf(STAIR_W_TEMP_LN_G), dS0, G(b, STAIR_W_TEMP_LN_G), G(b, STAIR_W_TEMP_LN_B), M, H, c.st));
        float* Gx = b.ws.take<float>(static_cast<long long>(M) * H);
        STAGE_GATHER(fp, c.buf.vid, dt, a0, T, T, n, H);
        STAIR_TRY(linear_bwd(b, dS0, H, S0, H, r, fp, M, M, H, H, STAIR_W_TEMP_D_W, STAIR_W_TEMP_D_B, Gx, DS(b)));
        RUN(launch_rowscale_bwd(dt, Gx, c.buf.vid, a0, T, T, r, dr, M, H, c.st));
        RUN(launch_scatter_add_rows(Gx, a0, T, T, tr.dvid, M, H, c.st));
        const float* params[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        float* dparams[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        if (mode > 0) for (int j = 0; j < 6; ++j) { params[j] = c.Wf(STAIR_W_TEMP_REL_BEFORE + 6 * (mode - 1) + j); dparams[j] = G(b, STAIR_W_TEMP_REL_BEFORE + 6 * (mode - 1) + j); }
        RUN(launch_temporal_relate_bwd(att, a1, K, mode, c.m.conv_k, params, dparams, dr, tr.datt, n, T, c.st));
        break;
    }
    case STAIR_OP_FILTER: {
        const int w = STAIR_W_FILT_REPR + 4 * g.variant;
        float* dAgg = b.ws.take<float>(static_cast<long long>(n) * H);
        STAGE_ACT(aggp, S2, dt, n, H);
        STAIR_TRY(linear_bwd(b, dvec_out, H, vec_out, H, nullptr, aggp, n, n, H, H, STAIR_W_FILT_D_W, STAIR_W_FILT_D_B, dAgg));
        float* dx = b.ws.take<float>(static_cast<long long>(M) * H);
        RUN(launch_bcast_T(dAgg, dx, n, T, H, c.st));
        STAIR_TRY(mlp2_bwd(b, dx, S0, S1, a0, n, w));
        break;
    }
    case STAIR_OP_FILTERFRAME: {
        const int w = STAIR_W_FF_REPR + 4 * g.variant;
        const float* gate = g.variant == 0 ? c.at<float>(c.plan.a0) : nullptr;
        if (g.head && (tr.n_ff > 0 || tr.ext_dhead_ff) && tr.dhead_ff && ab >= 0) {      // criterion_filterframe (or an external seed): through pretrain_head = Linear(H, O) into d(vid_out)
            const int O = c.m.O;
            float* dHd = tr.dhead_ff + static_cast<long long>(ab) * T * O;
            float* dOut2 = b.ws.take<float>(static_cast<long long>(M) * H);
            STAGE_ACT(vop, vid_out, dt, M, H);
            STAIR_TRY(linear_bwd(b, dHd, O, nullptr, 0, nullptr, vop, M, M, O, H, STAIR_W_FF_HEAD_W, STAIR_W_FF_HEAD_B, dOut2));
            RUN(launch_add_inplace(dvid_out, dOut2, static_cast<long long>(M) * H, c.st));
        }
        float* Gx = b.ws.take<float>(static_cast<long long>(M) * H);
        STAGE_ACT(xp, S1, dt, M, H);
        STAIR_TRY(linear_bwd(b, dvid_out, H, vid_out, H, gate, xp, M, M, H, H, STAIR_W_FF_D_W, STAIR_W_FF_D_B, Gx, DS(b)));
        if (gate) {
            float* da = b.ws.take<float>(M);
            if (!b.dry && cudaMemsetAsync(da, 0, sizeof(float) * M, c.st) != cudaSuccess) return STAIR_ERR_CUDA;
            RUN(launch_rowscale_bwd(dt, Gx, S1, nullptr, 1, 1, gate, da, M, H, c.st));
            RUN(launch_ff_attn_bwd(dt, S1, c.buf.vec, a1, c.Wf(STAIR_W_FF_ATT_W), gate, da, Gx, tr.dvec, G(b, STAIR_W_FF_ATT_W), G(b, STAIR_W_FF_ATT_B), n, T, H, c.st));
        }
        STAIR_TRY(mlp2_bwd(b, Gx, S0, S1, a0, n, w));
        break;
    }
    case STAIR_OP_HASITEM: {
        float* dS0 = b.ws.take<float>(static_cast<long long>(M) * H);
        RUN(launch_rowdot_sigmoid_bwd(dt, S0, c.Wf(STAIR_W_HAS1_W), att + static_cast<long long>(ob) * T, datt_out, dS0, G(b, STAIR_W_HAS1_W), G(b, STAIR_W_HAS1_B), M, H, DS(b), c.st));
        float* dF = b.ws.take<float>(static_cast<long long>(M) * H);
        STAGE_GATHER(fp, c.buf.vid, dt, a0, T, T, n, H);
        STAIR_TRY(linear_bwd(b, dS0, H, S0, H, nullptr, fp, M, M, H, H, STAIR_W_HAS0_W, STAIR_W_HAS0_B, dF, DS(b)));
        RUN(launch_scatter_add_rows(dF, a0, T, T, tr.dvid, M, H, c.st));
        break;
    }
    case STAIR_OP_EXISTSFRAME:
        RUN(launch_existsframe_bwd(dt, c.buf.vid, a1, c.buf.vec, a0, tr.datt, ob, tr.dvid, tr.dvec, n, T, H, c.st));
        break;
    case STAIR_OP_RELATE:
        RUN(launch_relate_bwd(att, ob, tr.datt, a0, g.variant >= 2 ? 0 : (g.variant == 0 ? 1 : -1), tr.datt, G(b, STAIR_W_REL_BETA), n, T, c.st));
        break;
    case STAIR_OP_ATTNVIDEO:
        RUN(launch_attnvideo_bwd(dt, dvid_out, c.buf.vid, a0, att, a1, tr.datt, tr.dvid, n, T, H, c.st));
        break;
    case STAIR_OP_AND:
    case STAIR_OP_XORFRAME: {
        const int op = g.op == STAIR_OP_AND ? STAIR_BIN_MIN : STAIR_BIN_ABSDIFF;
        if (g.variant == 0) RUN(launch_binary_bwd(dt, c.buf.vec, a0, a1, dvec_out, tr.dvec, H, H, op, n, c.st));
        else RUN(launch_binary_bwd(STAIR_F32, att, a0, a1, datt_out, tr.datt, T, g.variant * T, op, n, c.st));
        break;
    }
    case STAIR_OP_CHOOSE:
        RUN(launch_choose_bwd(dt, c.buf.vec, a0, a1, a2, dvec_out, tr.dvec, n, H, c.st));
        break;
    case STAIR_OP_ARRAY2:
        RUN(launch_array2_bwd(dvec_out, a0, a1, tr.dvec, n, H, c.st));
        break;
    case STAIR_OP_COMPARE:
    case STAIR_OP_EQUALS:
    case STAIR_OP_XOR: {
        const int mode = g.op == STAIR_OP_XOR ? STAIR_CAT_XOR : STAIR_CAT_PAIR;
        const int Kc = (mode == STAIR_CAT_XOR ? 3 : 2) * H;
        const int wid = g.op == STAIR_OP_XOR ? STAIR_W_XOR_W : (g.op == STAIR_OP_EQUALS ? STAIR_W_EQUALS_W : STAIR_W_COMPARE_W);
        float* dcat = b.ws.take<float>(static_cast<long long>(n) * Kc);
        STAIR_TRY(linear_bwd(b, dvec_out, H, vec_out, H, nullptr, VP, n, n, H, Kc, wid, wid + 1, dcat));
        RUN(launch_concat_bwd(dt, c.buf.vec, a0, a1, mode, dcat, tr.dvec, n, H, c.st));
        break;
    }
    case STAIR_OP_EXISTS:
    case STAIR_OP_TOACTION: {
        const bool ex = g.op == STAIR_OP_EXISTS;
        const int mode = ex ? STAIR_CAT_EXISTS : STAIR_CAT_PAIR;
        const int Kc = (ex ? 3 : 2) * H;
        const int w0 = ex ? STAIR_W_EXISTS0_W : STAIR_W_TOACT0_W, w1 = ex ? STAIR_W_EXISTS1_W : STAIR_W_TOACT1_W;
        float* dV0 = b.ws.take<float>(static_cast<long long>(n) * H);
        STAGE_ACT(v0p, V0, dt, n, H);
        STAIR_TRY(linear_bwd(b, dvec_out, H, vec_out, H, nullptr, v0p, n, n, H, H, w1, w1 + 1, dV0, ex ? DS(b) : 1.0f));   // Exists drops after both ReLUs
        float* dcat = b.ws.take<float>(static_cast<long long>(n) * Kc);
        STAIR_TRY(linear_bwd(b, dV0, H, V0, H, nullptr, VP, n, n, H, Kc, w0, w0 + 1, dcat, DS(b)));
        RUN(launch_concat_bwd(dt, c.buf.vec, a0, a1, mode, dcat, tr.dvec, n, H, c.st));
        break;
    }
    case STAIR_OP_SUPERLATIVE: {
        const int is_min = g.variant & 1, kind = g.variant >> 1;
        const int K = kind == 0 ? 1 : (kind == 1 ? 2 : T);
        float* ats = c.at<float>(c.plan.ats);
        float* dV0 = b.ws.take<float>(static_cast<long long>(n) * H);
        STAGE_ACT(v0p, V0, dt, n, H);
        STAIR_TRY(linear_bwd(b, dvec_out, H, vec_out, H, nullptr, v0p, n, n, H, H, STAIR_W_SUP_D_W, STAIR_W_SUP_D_B, dV0));
        float* datt_s = b.ws.take<float>(static_cast<long long>(n) * K * T);
        const void* act_base = kind == 2 ? c.buf.vid : c.buf.vec;
        float* dact_base = kind == 2 ? tr.dvid : tr.dvec;
        RUN(launch_super_mix_bwd(dt, ats, K, T, H, is_min, act_base, a0, kind == 2 ? T : 1, dV0, datt_s, dact_base, n, c.st));
        STAIR_TRY(localize_bwd(b, datt_s, 0, S0, S1, S2, a1, act_base, dact_base, a0, K, kind == 2 ? T : 1, K, n));
        break;
    }
    default:
        return STAIR_ERR_LAYOUT;
    }
    b.ws.off = mark;
    return STAIR_OK;
}

// ---- encoders with history (training forward) ---------------------------------------------------------------------------------------
struct EncIO {
    const void* xin; int xin_dt; long long rows; int Kin, Kin_ld, wih, bias, whh_f, whh_r;
    void* xproj; void* out; void* qfeat; const int* q_off; int S;
};

EncIO enc_io(Ctx& c, int e) {
    EncIO io;
    const StairModel& m = c.m; const StairBatch& bt = c.b;
    if (e == 0) {
        io.xin = bt.video; io.xin_dt = bt.video_dtype; io.rows = static_cast<long long>(bt.B) * c.T; io.Kin = m.V; io.Kin_ld = m.V_ld;
        io.wih = STAIR_W_VENC_WIH; io.bias = STAIR_W_VENC_B; io.whh_f = STAIR_W_VENC_WHH_F; io.whh_r = STAIR_W_VENC_WHH_R;
        io.xproj = c.at<void>(c.plan.xv); io.out = c.buf.vid; io.qfeat = nullptr; io.q_off = nullptr; io.S = c.T;
    } else {
        io.xin = bt.question; io.xin_dt = bt.question_dtype; io.rows = bt.n_tok; io.Kin = m.text_size; io.Kin_ld = m.text_ld;
        io.wih = STAIR_W_TENC_WIH; io.bias = STAIR_W_TENC_B; io.whh_f = STAIR_W_TENC_WHH_F; io.whh_r = STAIR_W_TENC_WHH_R;
        io.xproj = c.at<void>(c.plan.xq); io.out = c.buf.tokfeat; io.qfeat = c.buf.qfeat; io.q_off = bt.q_off; io.S = bt.L_max;
    }
    return io;
}

// true when the training forward runs the fused recurrence (history of gates / cell state in the blocked layout)
bool fused_history(const Ctx& c) {
    return lstm_fused_ok(c.m.precision, c.h) && g_lstm_impl == 0 && c.W(STAIR_W_VENC_WHHI_F) && c.W(STAIR_W_TENC_WHHI_F);
}

// true when the training forward, its history and the BPTT run the text encoder over length-sorted questions (the schedule of
// StairBatch.q_order / q_soff / tok_src; fused forward + persistent fused BPTT only — both read the history by grid row).  The forward and
// the backward of a step must see the same switches (stair_set_text_sort / stair_set_bptt_impl / stair_set_lstm_impl).
bool train_text_sorted(const Ctx& c) {
    const StairModel& m = c.m;
    return g_text_sort && c.b.q_order && c.b.q_soff && c.b.tok_src && fused_history(c) && !lstm_ws_ok(m.precision, c.h, c.b.B) && g_bptt_impl == 0 &&
           lstm_bptt_fused_ok(m.precision, c.h) && m.wt[STAIR_W_VENC_WHH_F] && m.wt[STAIR_W_VENC_WHH_R] && m.wt[STAIR_W_TENC_WHH_F] && m.wt[STAIR_W_TENC_WHH_R];
}

int run_encoders_train(Ctx& c, const StairTrain& tr) {
    const StairModel& m = c.m; const StairBatch& bt = c.b;
    const int B = bt.B, H = c.H, h = c.h;
    const SavedLayout SL = saved_layout(m, bt);
    if (tr.saved_bytes < SL.total) return STAIR_ERR_CAPACITY;
    char* sv = reinterpret_cast<char*>(tr.saved);
    float* g = c.at<float>(c.plan.g);
    // bf16 path: both recurrences run in the persistent fused kernel (csrc/lstm_fused.cu), which writes the BPTT history itself
    const bool fused = fused_history(c);
    const bool tsort = train_text_sorted(c);
    LstmHist hist;
    for (int e = 0; e < 2; ++e) {
        const EncIO io = enc_io(c, e);
        bf16* in = c.at<bf16>(e == 0 ? c.plan.xv_in : c.plan.xq_in);
        const bool direct = e == 0 && bt.video_dtype == STAIR_BF16 && c.np == 1 && (m.V % 8) == 0;
        GemmArgs a;
        if (direct) { a.A = io.xin; a.lda = io.Kin; a.a_plane_rows = 0; }
        else {
            STAIR_TRY(launch_stage_rows(io.xin_dt, io.xin, io.Kin, (e == 1 && tsort) ? bt.tok_src : nullptr, 1, 1, in, io.Kin_ld, io.rows, c.np, io.rows,
                                        io.Kin, c.st));
            a.A = in; a.lda = io.Kin_ld; a.a_plane_rows = static_cast<int>(io.rows);
        }
        a.nplanes = c.np; a.W = c.W(io.wih); a.ldw = io.Kin_ld; a.w_plane_rows = 4 * H; a.bias = c.Wf(io.bias); a.C = io.xproj; a.ldc = 4 * H;
        a.out_dtype = c.adt; a.M = static_cast<int>(io.rows); a.N = 4 * H; a.K = io.Kin;
        STAIR_TRY(launch_gemm(a, c.st));
        const int S = io.S;
        float* gates = reinterpret_cast<float*>(sv + SL.enc[e].gates);
        float* cs = reinterpret_cast<float*>(sv + SL.enc[e].c);
        bf16* hs = reinterpret_cast<bf16*>(sv + SL.enc[e].hs);
        const long long hs_dir = static_cast<long long>(S + 1) * B * h;      // [np][2][S+1][B][h]
        const long long hs_plane = 2 * hs_dir;
        if (cudaMemsetAsync(hs, 0, sizeof(bf16) * c.np * hs_plane, c.st) != cudaSuccess) return STAIR_ERR_CUDA;
        if (fused) {
            hist.gates[e] = gates; hist.c[e] = cs; hist.hs[e] = hs; hist.hs_dir[e] = hs_dir;
            // finished questions of the ragged text batch are not written by the fused kernel: their gate history must read as 0
            // nowhere (the BPTT cell skips inactive rows) but their cell history is read as c_prev only while active — nothing to clear
            continue;
        }
        for (int s = 0; s < S; ++s) {
            if (s > 0)
                for (int d = 0; d < 2; ++d) {
                    GemmArgs r;
                    r.A = hs + d * hs_dir + static_cast<long long>(s) * B * h; r.lda = h; r.a_plane_rows = static_cast<int>(hs_plane / h);
                    r.nplanes = c.np; r.W = c.W(d == 0 ? io.whh_f : io.whh_r); r.ldw = h; r.w_plane_rows = 4 * h;
                    r.C = g + static_cast<long long>(d) * B * 4 * h; r.ldc = 4 * h; r.out_dtype = STAIR_F32; r.M = B; r.N = 4 * h; r.K = h;
                    STAIR_TRY(launch_gemm(r, c.st));
                }
            const float* c_prev = s > 0 ? cs + static_cast<long long>(s - 1) * 2 * B * h : cs;
            STAIR_TRY(launch_lstm_cell_train(c.adt, io.xproj, g, c_prev, cs + static_cast<long long>(s) * 2 * B * h,
                                             gates + static_cast<long long>(s) * 2 * B * 4 * h, hs + static_cast<long long>(s + 1) * B * h,
                                             hs + static_cast<long long>(s) * B * h, hs_plane, hs_dir, c.np, c.adt, io.out, io.qfeat, io.q_off, B,
                                             c.T, h, s, c.st));
        }
    }
    if (fused && lstm_ws_ok(m.precision, h, B)) {
        // weight-stationary cluster recurrence (lstm_ws.cu), one launch per encoder, writing the same BPTT history
        float* cs = c.at<float>(c.plan.c);
        char* hx = c.at<char>(c.plan.hx);
        const long long nblk128 = (B + 127) / 128;
        STAIR_TRY(launch_lstm_ws(c.at<void>(c.plan.xv), c.buf.vid, nullptr, nullptr, c.T, c.W(STAIR_W_VENC_WHHI_F), c.W(STAIR_W_VENC_WHHI_R), cs, hx,
                                 B, h, err_flag_ptr(), c.st, reinterpret_cast<bf16*>(hist.gates[0]), hist.hs[0]));
        return launch_lstm_ws(c.at<void>(c.plan.xq), c.buf.tokfeat, c.buf.qfeat, bt.q_off, bt.L_max, c.W(STAIR_W_TENC_WHHI_F), c.W(STAIR_W_TENC_WHHI_R),
                              cs + 2 * nblk128 * 128 * h, hx + lstm_ws_hx_bytes(B), B, h, err_flag_ptr(), c.st, reinterpret_cast<bf16*>(hist.gates[1]),
                              hist.hs[1]);
    }
    if (fused)
        return launch_lstm_fused(c.at<void>(c.plan.xv), c.buf.vid, c.T, c.W(STAIR_W_VENC_WHHI_F), c.W(STAIR_W_VENC_WHHI_R),
                                 c.at<void>(c.plan.xq), c.buf.tokfeat, c.buf.qfeat, bt.q_off, bt.L_max, c.W(STAIR_W_TENC_WHHI_F),
                                 c.W(STAIR_W_TENC_WHHI_R), c.at<float>(c.plan.c), B, h, 1, 1, err_flag_ptr(), c.st, &hist, tsort ? bt.q_order : nullptr,
                                 tsort ? bt.q_soff : nullptr);
    return STAIR_OK;
}

// BPTT of one encoder (e = 0 video, 1 text) on the stream of b.c; `dir_lane` = index of the side stream that runs the reverse direction's
// recurrent GEMMs next to the forward direction's
int encoder_bwd(BCtx& b, int e, int dir_lane, bf16* dxb_done = nullptr) {
    Ctx& c = b.c;
    const StairTrain& tr = b.tr;
    const StairModel& m = c.m; const StairBatch& bt = c.b;
    const int B = bt.B, H = c.H, h = c.h;
    const SavedLayout SL = saved_layout(m, bt);
    char* sv = reinterpret_cast<char*>(tr.saved);
    const bool blocked = fused_history(c);
    {
        const EncIO io = enc_io(c, e);
        const int S = io.S;
        const long long mark = b.ws.off;
        const float* gates = reinterpret_cast<const float*>(sv + SL.enc[e].gates);
        const float* cs = reinterpret_cast<const float*>(sv + SL.enc[e].c);
        const bf16* hs = reinterpret_cast<const bf16*>(sv + SL.enc[e].hs);
        const long long hs_dir = static_cast<long long>(S + 1) * B * h, hs_plane = 2 * hs_dir;
        // blocked (= fused bf16 forward): the gate gradients are kept ONCE, as bf16 rows in token order [rows][8h] — operand of both weight
        // gradients (dW_ih = dG^T . x, dW_hh = dG^T . h_prev with h_prev in token order, written by the fused forward) and source of the
        // bias gradient; only the current step's slice [2][B][4h] is also written for the recurrent GEMM.  Otherwise (step-wise forward,
        // fp32-strict planes): fp32 dxproj + the direction-major plane history.
        // dxb_done: the persistent fused BPTT kernel (encoders_bwd) already produced the token-order gate gradients of this encoder
        float* dxproj = blocked ? nullptr : b.ws.take<float>(io.rows * 8 * h);
        bf16* dxb = dxb_done ? dxb_done : blocked ? b.ws.take<bf16>(io.rows * 8 * h) : nullptr;
        const int S_loop = dxb_done ? 0 : S;
        const long long dg_dir = blocked ? static_cast<long long>(B) * 4 * h : static_cast<long long>(S) * B * 4 * h, dg_plane = 2 * dg_dir;
        bf16* dg_hist = dxb_done ? nullptr : b.ws.take<bf16>(static_cast<long long>(c.np) * dg_plane);     // gate gradients, bf16 planes [np][2][S | 1][B][4h]
        float* dh_rec = dxb_done ? nullptr : b.ws.take<float>(2LL * B * h);
        float* dc = dxb_done ? nullptr : b.ws.take<float>(2LL * B * h);
        const float* dout = e == 0 ? tr.dvid : tr.dtokfeat;
        for (int s = S_loop - 1; s >= 0; --s) {
            bf16* dg_step = blocked ? dg_hist : dg_hist + static_cast<long long>(s) * B * 4 * h;
            const float* c_prev = s > 0 ? cs + static_cast<long long>(s - 1) * 2 * B * h : cs;
            if (blocked)
                RUN(launch_lstm_cell_bwd(gates, nullptr, cs, dout, dh_rec, e == 1 ? tr.dqfeat : nullptr, dc, dg_dir, dg_step, dg_plane, c.np, nullptr, dxb,
                                         io.q_off, B, c.T, h, s, S - 1, 1, c.st));
            else
                RUN(launch_lstm_cell_bwd(gates + static_cast<long long>(s) * 2 * B * 4 * h, c_prev, cs + static_cast<long long>(s) * 2 * B * h, dout, dh_rec,
                                         e == 1 ? tr.dqfeat : nullptr, dc, dg_dir, dg_step, dg_plane, c.np, dxproj, nullptr, io.q_off, B, c.T, h, s, S - 1, 0, c.st));
            if (s > 0) {
                // dh_{s-1} = dG_s . W_hh of the two directions are independent and small (B/128 x h/128 tiles each): the reverse direction
                // runs on a side stream next to the forward one instead of after it
                LaneStreams* ls = (!b.dry && g_lanes > 1) ? lane_streams() : nullptr;
                if (ls) {
                    if (cudaEventRecord(ls->join[dir_lane + 1], c.st) != cudaSuccess || cudaStreamWaitEvent(ls->side[dir_lane], ls->join[dir_lane + 1], 0) != cudaSuccess) return STAIR_ERR_CUDA;
                }
                for (int d = 0; d < 2; ++d) {
                    const int wid = d == 0 ? io.whh_f : io.whh_r;
                    if (!m.wt[wid]) return STAIR_ERR_ARG;
                    GemmArgs r;
                    r.A = dg_step + d * dg_dir; r.lda = 4 * h; r.a_plane_rows = static_cast<int>(dg_plane / (4 * h)); r.nplanes = c.np;
                    r.W = m.wt[wid]; r.ldw = 4 * h; r.w_plane_rows = h; r.C = dh_rec + static_cast<long long>(d) * B * h; r.ldc = h;
                    r.out_dtype = STAIR_F32; r.M = B; r.N = h; r.K = 4 * h;
                    RUN(launch_gemm(r, (ls && d == 1) ? ls->side[dir_lane] : c.st));
                }
                if (ls) {
                    if (cudaEventRecord(ls->join[dir_lane], ls->side[dir_lane]) != cudaSuccess || cudaStreamWaitEvent(c.st, ls->join[dir_lane], 0) != cudaSuccess) return STAIR_ERR_CUDA;
                }
            }
        }
        if (blocked) {
            const long long rows = io.rows;
            const bf16* hs_tok = hs;                              // [rows][2h] h_prev in token order (lstm_fused.cu HIST)
            for (int d = 0; d < 2; ++d) {                         // dW_hh[d] += dG[:, d]^T . h_prev[:, d] over all token rows
                float* gW = G(b, d == 0 ? io.whh_f : io.whh_r);
                if (!gW) continue;
                GemmArgs a;
                a.A = dxb + d * 4 * h; a.lda = 8 * h; a.a_plane_rows = static_cast<int>(rows); a.nplanes = 1;
                a.W = hs_tok + d * h; a.ldw = 2 * h; a.w_plane_rows = static_cast<int>(rows);
                a.C = gW; a.ldc = h; a.out_dtype = STAIR_F32; a.M = 4 * h; a.N = h; a.K = static_cast<int>(rows); a.accumulate = 1; a.mn_major = 1;
                RUN(launch_gemm(a, c.st));
            }
            // d(b_ih + b_hh) = column sums of the gate gradients; dW_ih = dG^T . x over the token rows (x in place when it is bf16)
            RUN(launch_colsum_bf16(dxb, rows, 8 * h, 8 * h, G(b, io.bias), c.st));
            float* gWih = G(b, io.wih);
            if (gWih) {
                const bool direct = e == 0 && bt.video_dtype == STAIR_BF16 && (m.V % 8) == 0;
                const bf16* xin_p = reinterpret_cast<const bf16*>(io.xin);
                if (!direct) {
                    bf16* in = b.ws.take<bf16>(io.rows * io.Kin_ld);
                    const int* src = (e == 1 && dxb_done && train_text_sorted(c)) ? bt.tok_src : nullptr;      // the gate gradients' row order
                    RUN(launch_stage_rows(io.xin_dt, io.xin, io.Kin, src, 1, 1, in, io.Kin_ld, io.rows, 1, io.rows, io.Kin, c.st));
                    xin_p = in;
                }
                GemmArgs a;
                a.A = dxb; a.lda = 8 * h; a.a_plane_rows = static_cast<int>(rows); a.nplanes = 1;
                a.W = xin_p; a.ldw = io.Kin_ld; a.w_plane_rows = static_cast<int>(rows);
                a.C = gWih; a.ldc = io.Kin; a.out_dtype = STAIR_F32; a.M = 8 * h; a.N = io.Kin; a.K = static_cast<int>(rows); a.accumulate = 1; a.mn_major = 1;
                RUN(launch_gemm(a, c.st));
            }
            b.ws.off = mark;
            return STAIR_OK;
        }
        // dW_hh[d] += sum_s dG[d][s]^T h[d][s-1]: one contraction over all (step, question) rows per direction, both operands in place
        for (int d = 0; d < 2; ++d) {
            float* gW = G(b, d == 0 ? io.whh_f : io.whh_r);
            if (!gW) continue;
            GemmArgs a;
            a.A = dg_hist + d * dg_dir; a.lda = 4 * h; a.a_plane_rows = static_cast<int>(dg_plane / (4 * h)); a.nplanes = c.np;
            a.W = hs + d * hs_dir; a.ldw = h; a.w_plane_rows = static_cast<int>(hs_plane / h);
            a.C = gW; a.ldc = h; a.out_dtype = STAIR_F32; a.M = 4 * h; a.N = h; a.K = S * B; a.accumulate = 1; a.mn_major = 1;
            RUN(launch_gemm(a, c.st));
        }
        // dW_ih, d(b_ih + b_hh) from the gate gradients of every frame / token
        {
            const bool direct = e == 0 && bt.video_dtype == STAIR_BF16 && c.np == 1 && (m.V % 8) == 0;
            const bf16* xin_p; long long xrows;
            if (direct) { xin_p = reinterpret_cast<const bf16*>(io.xin); xrows = io.rows; }
            else {
                bf16* in = b.ws.take<bf16>(static_cast<long long>(c.np) * io.rows * io.Kin_ld);
                RUN(launch_stage_rows(io.xin_dt, io.xin, io.Kin, nullptr, 1, 1, in, io.Kin_ld, io.rows, c.np, io.rows, io.Kin, c.st));
                xin_p = in; xrows = io.rows;
            }
            STAIR_TRY(linear_bwd(b, dxproj, 4 * H, nullptr, 0, nullptr, xin_p, xrows, static_cast<int>(io.rows), 4 * H, io.Kin, io.wih, io.bias, nullptr));
        }
        b.ws.off = mark;
    }
    return STAIR_OK;
}

// The two encoders' BPTT chains are independent (24 + 8 sequential steps of small kernels): the video encoder runs on a side stream
// next to the text encoder, each with its own slice of the backward workspace.
int encoders_bwd(BCtx& b) {
    Ctx& c = b.c;
    LaneStreams* ls = (!b.dry && g_lanes > 1) ? lane_streams() : nullptr;
    const long long mark0 = b.ws.off;
    // bf16 path: the whole time loop of both encoders and both directions is ONE persistent launch (lstm_bptt.cu); what remains per
    // encoder is the weight / bias gradients from the token-order gate gradients it leaves behind
    bf16* dxb_done[2] = {nullptr, nullptr};
    const StairModel& m = c.m;
    const bool fused_bptt = fused_history(c) && g_bptt_impl == 0 && lstm_bptt_fused_ok(m.precision, c.h) && m.wt[STAIR_W_VENC_WHH_F] &&
                            m.wt[STAIR_W_VENC_WHH_R] && m.wt[STAIR_W_TENC_WHH_F] && m.wt[STAIR_W_TENC_WHH_R];
    if (fused_bptt) {
        const int B = c.b.B, h = c.h;
        const SavedLayout SL = saved_layout(m, c.b);
        char* sv = reinterpret_cast<char*>(b.tr.saved);
        LstmBptt a;
        for (int e = 0; e < 2; ++e) {
            const EncIO io = enc_io(c, e);
            dxb_done[e] = b.ws.take<bf16>(io.rows * 8 * h);
            a.gates[e] = reinterpret_cast<const float*>(sv + SL.enc[e].gates);
            a.c[e] = reinterpret_cast<const float*>(sv + SL.enc[e].c);
            a.dout[e] = e == 0 ? b.tr.dvid : b.tr.dtokfeat;
            a.dxb[e] = dxb_done[e];
            a.dc[e] = b.ws.take<float>(2LL * align_up(B, 64) * h);
            a.whhT[2 * e] = m.wt[io.whh_f]; a.whhT[2 * e + 1] = m.wt[io.whh_r];
        }
        a.dqfeat = b.tr.dqfeat;
        const bool tsort = train_text_sorted(c);
        RUN(launch_lstm_bptt_fused(a, B, h, c.T, c.b.L_max, c.b.q_off, err_flag_ptr(), c.st, tsort ? c.b.q_order : nullptr, tsort ? c.b.q_soff : nullptr));
    }
    const long long mark = b.ws.off;
    // video (e = 0): side stream 2 (+ side stream 3 for its reverse-direction GEMMs); workspace [mark, video peak)
    Ctx cv = c;
    if (ls) {
        cv.st = ls->side[2];
        if (cudaEventRecord(ls->fork, c.st) != cudaSuccess || cudaStreamWaitEvent(cv.st, ls->fork, 0) != cudaSuccess) return STAIR_ERR_CUDA;
    }
    BCtx bv{cv, b.tr, b.ws, b.dry};
    bv.ws.peak = mark;
    const int rc_v = encoder_bwd(bv, 0, 3, dxb_done[0]);
    if (rc_v != STAIR_OK) return rc_v;
    if (bv.ws.overflow) b.ws.overflow = true;
    // text (e = 1): caller's stream (+ side stream 0), workspace after the video encoder's
    b.ws.off = bv.ws.peak;
    if (b.ws.off > b.ws.peak) b.ws.peak = b.ws.off;
    STAIR_TRY(encoder_bwd(b, 1, 0, dxb_done[1]));
    b.ws.off = mark0;
    if (ls) {
        if (cudaEventRecord(ls->join[2], ls->side[2]) != cudaSuccess || cudaStreamWaitEvent(c.st, ls->join[2], 0) != cudaSuccess) return STAIR_ERR_CUDA;
    }
    return STAIR_OK;
}

int losses(BCtx& b) {
    Ctx& c = b.c;
    const StairTrain& tr = b.tr;
    const int H = c.H;
    const int* aux_slot = c.buf.itab + c.il.aux_slot;
    RUN(launch_loss_att(c.buf.att, tr.datt, c.out_slot, aux_slot, tr.att_node, tr.att_kind, tr.att_slot, tr.att_gold, tr.att_w, tr.loss, tr.n_att, c.T, c.st));
    const float* hw[3] = {c.Wf(STAIR_W_EQUALS_HEAD_W), c.Wf(STAIR_W_XOR_HEAD_W), c.Wf(STAIR_W_EXISTS_HEAD_W)};
    const float* hb[3] = {c.Wf(STAIR_W_EQUALS_HEAD_B), c.Wf(STAIR_W_XOR_HEAD_B), c.Wf(STAIR_W_EXISTS_HEAD_B)};
    float* dhw[3] = {G(b, STAIR_W_EQUALS_HEAD_W), G(b, STAIR_W_XOR_HEAD_W), G(b, STAIR_W_EXISTS_HEAD_W)};
    float* dhb[3] = {G(b, STAIR_W_EQUALS_HEAD_B), G(b, STAIR_W_XOR_HEAD_B), G(b, STAIR_W_EXISTS_HEAD_B)};
    if (tr.n_bin > 0) {
        for (int j = 0; j < 3; ++j) if (!hw[j] || !hb[j]) return STAIR_ERR_ARG;
        RUN(launch_loss_bin(c.adt, c.buf.vec, tr.dvec, c.out_slot, tr.bin_node, nullptr, tr.bin_label, tr.bin_w, hw, hb, dhw, dhb, tr.bin_which,
                            tr.loss, tr.n_bin, H, c.st));
    }
    RUN(launch_loss_con(c.adt, c.buf.vec, tr.dvec, c.out_slot, tr.con_node, tr.con_pos, tr.con_w, tr.cls_rep, tr.n_cls, tr.loss, tr.n_con, H, c.st));
    RUN(launch_loss_dec(c.buf.logits, tr.answer, tr.dec_w, tr.dlogits, tr.loss, c.b.B, c.m.A, c.st));
    if (tr.n_ff > 0) {
        if (!tr.dhead_ff || !c.buf.head_ff || c.m.O <= 0) return STAIR_ERR_ARG;
        RUN(launch_loss_ff(c.buf.head_ff, tr.dhead_ff, aux_slot, tr.ff_node, tr.ff_gold, tr.ff_w, tr.loss, tr.n_ff, c.T, c.m.O, c.st));
    }
    return STAIR_OK;
}

// StairTrain.ext_*: the caller's gradients with respect to logits / attention maps / head outputs replace the built-in criteria (the
// gradient arenas were just zeroed).  Head outputs are back-propagated through their pretrain heads into the VEC gradient arena here;
// the FilterFrame head is handled by the FilterFrame group's backward (dhead_ff).
int external_seeds(BCtx& b) {
    Ctx& c = b.c;
    const StairTrain& tr = b.tr;
    const long long T = c.T, H = c.H;
    if (!b.dry) {
        cudaError_t e = cudaMemcpyAsync(tr.dlogits, tr.ext_dlogits, sizeof(float) * c.b.B * c.m.A, cudaMemcpyDeviceToDevice, c.st);
        if (e == cudaSuccess && tr.ext_datt) e = cudaMemcpyAsync(tr.datt, tr.ext_datt, sizeof(float) * c.buf.att_rows * T, cudaMemcpyDeviceToDevice, c.st);
        if (e == cudaSuccess && tr.ext_dhead_ff) {
            if (!tr.dhead_ff) return STAIR_ERR_ARG;
            e = cudaMemcpyAsync(tr.dhead_ff, tr.ext_dhead_ff, sizeof(float) * tr.dhead_ff_elems, cudaMemcpyDeviceToDevice, c.st);
        }
        if (e != cudaSuccess) return STAIR_ERR_CUDA;
    }
    for (int gi = 0; gi < c.b.n_groups; ++gi) {
        const StairGroup& g = c.b.groups[gi];
        if (!g.head || g.aux_base < 0) continue;
        int wslot = -1, nout = 0;
        switch (g.op) {
        case STAIR_OP_EQUALS: wslot = STAIR_W_EQUALS_HEAD_W; nout = 1; break;
        case STAIR_OP_XOR: wslot = STAIR_W_XOR_HEAD_W; nout = 2; break;
        case STAIR_OP_EXISTS: wslot = STAIR_W_EXISTS_HEAD_W; nout = 2; break;
        case STAIR_OP_FILTER: case STAIR_OP_TOACTION: case STAIR_OP_SUPERLATIVE:
            if (tr.ext_dhead_vec) RUN(launch_l2norm_bwd(c.adt, c.buf.vec, g.out_base, tr.ext_dhead_vec, g.aux_base, tr.dvec, g.count, static_cast<int>(H), c.st));
            continue;
        default: continue;
        }
        if (!tr.ext_dhead_small) continue;
        if (!c.Wf(wslot)) return STAIR_ERR_ARG;
        RUN(launch_small_head_bwd(c.adt, c.buf.vec, g.out_base, c.Wf(wslot), nout, tr.ext_dhead_small, g.aux_base, tr.dvec, G(b, wslot), G(b, wslot + 1),
                                  g.count, static_cast<int>(H), c.st));
    }
    return STAIR_OK;
}

int decoder_bwd(BCtx& b) {
    Ctx& c = b.c;
    const StairTrain& tr = b.tr;
    const int B = c.b.B, H = c.H, A = c.m.A;
    bf16* VP = c.at<bf16>(c.plan.vp);
    void* D0 = c.at<void>(c.plan.v01);
    for (int done = 0; done < B; done += static_cast<int>(VEC_CAP)) {
        const int n = B - done < VEC_CAP ? B - done : static_cast<int>(VEC_CAP);
        const long long mark = b.ws.off;
        // recompute this chunk's decoder hidden layer (module_net.py:136-138)
        RUN(launch_decoder_concat(c.adt, c.buf.vec, c.b.root_node + done, c.out_slot, c.act_ptr(c.buf.qfeat, static_cast<long long>(done) * H), VP, n, c.np, n, H, c.st));
        if (!b.dry) drop_next(c, STAIR_W_DEC0_W, done);
        RUN(gemm_planes(c, VP, 2 * H, n, n, 2 * H, 2 * H, STAIR_W_DEC0_W, STAIR_W_DEC0_B, STAIR_ACT_RELU, nullptr, D0, c.adt, 2 * H));
        float* dD0 = b.ws.take<float>(static_cast<long long>(n) * 2 * H);
        STAGE_ACT(d0p, D0, c.adt, n, 2 * H);
        STAIR_TRY(linear_bwd(b, tr.dlogits + static_cast<long long>(done) * A, A, nullptr, 0, nullptr, d0p, n, n, A, 2 * H, STAIR_W_DEC1_W, STAIR_W_DEC1_B, dD0));
        float* dcat = b.ws.take<float>(static_cast<long long>(n) * 2 * H);
        STAIR_TRY(linear_bwd(b, dD0, 2 * H, D0, 2 * H, nullptr, VP, n, n, 2 * H, 2 * H, STAIR_W_DEC0_W, STAIR_W_DEC0_B, dcat, DS(b)));
        RUN(launch_decoder_concat_bwd(dcat, c.b.root_node + done, c.out_slot, tr.dvec, tr.dqfeat + static_cast<long long>(done) * H, n, H, c.st));
        b.ws.off = mark;
    }
    return STAIR_OK;
}

// backward of every chunk of group gi on the stream / workspace of b (forward scratch of the recompute path: b.c.ws)
int group_bwd(BCtx& b, int gi) {
    Ctx& c = b.c;
    const StairGroup& g = c.b.groups[gi];
    const int cap = group_cap(c, g);
    char* const ws0 = c.ws;
    long long k = c.act_base ? chunk_base(c, gi) : 0;
    for (int done = 0; done < g.count; done += cap, ++k) {
        const int n = g.count - done < cap ? g.count - done : cap;
        const int p = g.node_off + done, ob = g.out_base + done * g.out_mult, ab = g.aux_base >= 0 ? g.aux_base + done : -1;
        if (c.act_base) c.ws = c.act_base + k * c.plan.mod_bytes;           // the chunk's intermediates were kept by the training forward
        else if (g.op != STAIR_OP_WORD) RUN(run_chunk(c, g, p, n, ob, ab));  // recompute the chunk's intermediates
        const int rc = chunk_bwd(b, g, p, n, ob, ab);
        c.ws = ws0;
        if (rc != STAIR_OK) return rc;
    }
    return STAIR_OK;
}

int g_bwd_lanes = 8;      // 1 = groups one after the other on the caller's stream; > 1 = concurrent streams.  Training step at B = 4096: wave by wave on
                          // 4 / 8 lanes 6.04 / 6.06 ms, by dependency on 4 / 6 / 8 lanes 5.91 / 5.90 / 5.85 ms (profiles/r2_train_bwd_sched.txt)

// Module backward: groups in reverse schedule order.  The groups of one wave are independent of each other in the backward pass too
// (they read their own output gradients and ADD into shared gradient arenas / parameter gradients with atomics), so with
// g_bwd_lanes > 1 they run on concurrent streams, every lane with its own slice of the backward workspace: the small staging /
// scatter / reduction kernels of one group then overlap the GEMMs of another.
int modules_bwd(BCtx& b) {
    Ctx& c = b.c;
    const int ng = c.b.n_groups;
    const int max_lanes = g_bwd_lanes < LANES ? g_bwd_lanes : LANES;
    if (max_lanes <= 1) {
        for (int gi = ng - 1; gi >= 0; --gi) STAIR_TRY(group_bwd(b, gi));
        return STAIR_OK;
    }
    // workspace of the largest chunk (host-only pass: allocations without launches)
    long long chunk_peak = 0;
    {
        BCtx d{c, b.tr, Bump(), true};
        d.ws.dry = true;
        for (int gi = 0; gi < ng; ++gi) {
            d.ws.off = 0; d.ws.peak = 0;
            STAIR_TRY(group_bwd(d, gi));
            if (d.ws.peak > chunk_peak) chunk_peak = d.ws.peak;
        }
        chunk_peak = align_up(chunk_peak + 256, 1024);
    }
    const long long mark = b.ws.off;
    char* lane_base = b.ws.take<char>(max_lanes * chunk_peak);
    if (b.dry) { b.ws.off = mark; return STAIR_OK; }
    if (b.ws.overflow) return STAIR_ERR_CAPACITY;
    LaneStreams* ls = lane_streams();
    if (!ls) return STAIR_ERR_CUDA;
    if (g_dep_sched && c.b.group_deps && ng <= MAX_SCHED_GROUPS) {
        // Dependency-driven order (the mirror image of run_modules_dep): the backward of group g may start as soon as every CONSUMER of
        // its outputs has been back-propagated (their kernels add into g's output-gradient slots), not when the whole later wave has.
        const int* deps = c.b.group_deps;
        int lane_of[MAX_SCHED_GROUPS], tail[LANES];
        long long lane_end[LANES], finish[MAX_SCHED_GROUPS];
        bool used[LANES];
        for (int l = 0; l < max_lanes; ++l) { tail[l] = -1; lane_end[l] = 0; used[l] = false; }
        if (cudaEventRecord(ls->fork, c.st) != cudaSuccess) return STAIR_ERR_CUDA;
        auto consumes = [&](int cg, int g) {                       // does group cg read an output of group g?
            const int* d = deps + static_cast<long long>(cg) * STAIR_MAX_GROUP_DEPS;
            if (d[0] == -2) return true;                           // "everything before me"
            for (int k = 0; k < STAIR_MAX_GROUP_DEPS && d[k] >= 0; ++k) if (d[k] == g) return true;
            return false;
        };
        for (int g = ng - 1; g >= 0; --g) {
            long long ready = 0;
            int latest = -1;
            for (int cg = g + 1; cg < ng; ++cg)
                if (consumes(cg, g) && finish[cg] >= ready) { ready = finish[cg]; latest = cg; }
            int lane = -1;
            if (latest >= 0 && tail[lane_of[latest]] == latest) lane = lane_of[latest];
            else {
                lane = 0;
                for (int l = 1; l < max_lanes; ++l) if (lane_end[l] < lane_end[lane]) lane = l;
            }
            Ctx lc = c;
            cudaStream_t st = lane == 0 ? c.st : ls->side[lane - 1];
            if (lane > 0) { lc.st = st; lc.ws = c.ws + static_cast<long long>(lane) * c.plan.mod_bytes; }
            if (lane > 0 && !used[lane]) {
                if (cudaStreamWaitEvent(st, ls->fork, 0) != cudaSuccess) return STAIR_ERR_CUDA;
                used[lane] = true;
            }
            for (int cg = g + 1; cg < ng; ++cg)
                if (lane_of[cg] != lane && consumes(cg, g) && cudaStreamWaitEvent(st, ls->done[cg], 0) != cudaSuccess) return STAIR_ERR_CUDA;
            BCtx bl{lc, b.tr, Bump(), false};
            bl.ws.base = lane_base + lane * chunk_peak; bl.ws.cap = chunk_peak;
            STAIR_TRY(group_bwd(bl, g));
            if (bl.ws.overflow) return STAIR_ERR_CAPACITY;
            if (cudaEventRecord(ls->done[g], st) != cudaSuccess) return STAIR_ERR_CUDA;
            lane_of[g] = lane;
            tail[lane] = g;
            const long long start = ready > lane_end[lane] ? ready : lane_end[lane];
            finish[g] = lane_end[lane] = start + 2 * group_cost(c.b.groups[g], c.T);
        }
        for (int l = 1; l < max_lanes; ++l)
            if (used[l]) {
                if (cudaEventRecord(ls->join[l - 1], ls->side[l - 1]) != cudaSuccess) return STAIR_ERR_CUDA;
                if (cudaStreamWaitEvent(c.st, ls->join[l - 1], 0) != cudaSuccess) return STAIR_ERR_CUDA;
            }
        b.ws.off = mark;
        return STAIR_OK;
    }
    int gj = ng;
    while (gj > 0) {
        int gi = gj - 1;
        while (gi > 0 && c.b.groups[gi - 1].level == c.b.groups[gj - 1].level) --gi;      // wave = groups [gi, gj)
        const int nw = gj - gi;
        const int lanes = nw < max_lanes ? nw : max_lanes;
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
        if (lanes > 1) {
            if (cudaEventRecord(ls->fork, c.st) != cudaSuccess) return STAIR_ERR_CUDA;
            for (int l = 1; l < lanes; ++l)
                if (cudaStreamWaitEvent(ls->side[l - 1], ls->fork, 0) != cudaSuccess) return STAIR_ERR_CUDA;
        }
        for (int k = 0; k < nwc; ++k) {
            const int l = lane_of[k];
            Ctx lc = c;
            if (l > 0) { lc.st = ls->side[l - 1]; lc.ws = c.ws + static_cast<long long>(l) * c.plan.mod_bytes; }
            BCtx bl{lc, b.tr, Bump(), false};
            bl.ws.base = lane_base + l * chunk_peak; bl.ws.cap = chunk_peak;
            STAIR_TRY(group_bwd(bl, order[k]));
            if (bl.ws.overflow) return STAIR_ERR_CAPACITY;
        }
        for (int g = gi + nwc; g < gj; ++g) {                       // (more than 64 groups in a wave: the rest on the main lane)
            BCtx bl{c, b.tr, Bump(), false};
            bl.ws.base = lane_base; bl.ws.cap = chunk_peak;
            STAIR_TRY(group_bwd(bl, g));
        }
        for (int l = 1; l < lanes; ++l) {
            if (cudaEventRecord(ls->join[l - 1], ls->side[l - 1]) != cudaSuccess) return STAIR_ERR_CUDA;
            if (cudaStreamWaitEvent(c.st, ls->join[l - 1], 0) != cudaSuccess) return STAIR_ERR_CUDA;
        }
        gj = gi;
    }
    b.ws.off = mark;
    return STAIR_OK;
}

// phases: STAIR_BWD_MODULES = losses + decoder + module groups (every gradient slot except the encoders' is final afterwards),
// STAIR_BWD_ENCODERS = BPTT + encoder weight gradients.  Split so that a data-parallel caller can all-reduce the module gradients
// while the encoders are still back-propagating.
int run_backward(BCtx& b, int phases = STAIR_BWD_ALL) {
    Ctx& c = b.c;
    const StairTrain& tr = b.tr;
    const StairBuffers& buf = c.buf;
    const long long T = c.T, H = c.H;
    if (!(phases & STAIR_BWD_MODULES)) return (phases & STAIR_BWD_ENCODERS) ? encoders_bwd(b) : STAIR_OK;
    if (!b.dry) {
        cudaError_t e = cudaMemsetAsync(tr.dvid, 0, sizeof(float) * buf.vid_slots * T * H, c.st);
        if (e == cudaSuccess) e = cudaMemsetAsync(tr.dvec, 0, sizeof(float) * buf.vec_rows * H, c.st);
        if (e == cudaSuccess) e = cudaMemsetAsync(tr.datt, 0, sizeof(float) * buf.att_rows * T, c.st);
        if (e == cudaSuccess) e = cudaMemsetAsync(tr.dtokfeat, 0, sizeof(float) * c.b.n_tok * H, c.st);
        if (e == cudaSuccess) e = cudaMemsetAsync(tr.dqfeat, 0, sizeof(float) * c.b.B * H, c.st);
        if (e == cudaSuccess) e = cudaMemsetAsync(tr.loss, 0, sizeof(float) * 8, c.st);
        if (e == cudaSuccess && (tr.n_ff > 0 || tr.ext_dhead_ff) && tr.dhead_ff) e = cudaMemsetAsync(tr.dhead_ff, 0, sizeof(float) * tr.dhead_ff_elems, c.st);
        if (e != cudaSuccess) return STAIR_ERR_CUDA;
    }
    if (tr.ext_dlogits) STAIR_TRY(external_seeds(b));
    else STAIR_TRY(losses(b));
    STAIR_TRY(decoder_bwd(b));
    STAIR_TRY(modules_bwd(b));
    if (phases & STAIR_BWD_ENCODERS) STAIR_TRY(encoders_bwd(b));
    return STAIR_OK;
}

// StairTrain.act_saved (optional): one scratch region per (group, chunk) so that the backward does not re-run the module forward
int bind_saved_activations(Ctx& c, const StairTrain& tr) {
    c.act_base = nullptr;
    if (!tr.act_saved) return STAIR_OK;
    const long long need = total_chunks(c) * c.plan.mod_bytes + 1024;
    if (tr.act_saved_bytes < need) return STAIR_ERR_CAPACITY;
    c.act_base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(tr.act_saved) + 1023) & ~static_cast<uintptr_t>(1023));
    return STAIR_OK;
}

int make_ctx(Ctx& c, const StairModel& m, const StairBatch& b, const StairBuffers& buf) {
    if (m.H % 16 || m.H > 1024 || b.T <= 0 || m.V_ld % 8 || m.text_ld % 8) return STAIR_ERR_UNSUPPORTED;
    if (m.conv_k == 0 && b.T != m.T_max) return STAIR_ERR_UNSUPPORTED;
    make_plan(m, b, &c.plan);
    c.ws = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(buf.workspace) + 1023) & ~static_cast<uintptr_t>(1023));
    c.T = b.T; c.H = m.H; c.h = m.H / 2; c.np = m.precision == STAIR_F32 ? 3 : 1; c.adt = m.precision; c.esz = m.precision == STAIR_F32 ? 4 : 2;
    stair_itab_layout(b.n_nodes, b.n_groups, &c.il);
    c.perm = buf.itab + c.il.perm; c.out_slot = buf.itab + c.il.out_slot;
    c.arg0 = buf.itab + c.il.arg_slot; c.arg1 = c.arg0 + b.n_nodes; c.arg2 = c.arg1 + b.n_nodes;
    c.pos_q = buf.itab + c.il.pos_q; c.span_s = buf.itab + c.il.pos_span; c.span_e = c.span_s + b.n_nodes;
    return STAIR_OK;
}

}  // namespace
}  // namespace stair

using namespace stair;
using namespace stair::ex;

extern "C" int64_t stair_train_saved_bytes(const StairModel* model, const StairBatch* batch) {
    if (!model || !batch) return -1;
    return saved_layout(*model, *batch).total;
}

extern "C" int64_t stair_train_act_bytes(const StairModel* model, const StairBatch* batch, const StairBuffers* buf) {
    if (!model || !batch || !buf) return -1;
    Ctx c{*model, *batch, *buf, nullptr};
    if (make_ctx(c, *model, *batch, *buf) != STAIR_OK) return -1;
    return total_chunks(c) * c.plan.mod_bytes + 1024;
}

extern "C" int64_t stair_train_workspace_bytes(const StairModel* model, const StairBatch* batch, const StairBuffers* buf, const StairTrain* train) {
    if (!model || !batch || !buf || !train) return -1;
    Ctx c{*model, *batch, *buf, nullptr};
    if (make_ctx(c, *model, *batch, *buf) != STAIR_OK) return -1;
    BCtx b{c, *train, Bump(), true};
    b.ws.dry = true;
    if (run_backward(b) != STAIR_OK) return -1;
    return b.ws.peak + 4096;
}

extern "C" int stair_nmn_forward_train(const StairModel* model, const StairBatch* batch, const StairBuffers* buf, const StairTrain* train, void* stream) {
    if (!model || !batch || !buf || !train) return STAIR_ERR_ARG;
    if (batch->B <= 0) return STAIR_OK;
    Ctx c{*model, *batch, *buf, reinterpret_cast<cudaStream_t>(stream)};
    STAIR_TRY(make_ctx(c, *model, *batch, *buf));
    if (buf->workspace_bytes < c.plan.total || buf->itab_ints < c.il.total) return STAIR_ERR_CAPACITY;
    c.drop_p = train->dropout_p; c.drop_seed = train->dropout_seed;
    if (!(c.drop_p >= 0.0f && c.drop_p < 1.0f)) return STAIR_ERR_ARG;
    STAIR_TRY(bind_saved_activations(c, *train));
    const long long before = g_launch_count;
    STAIR_TRY(launch_group_layouts(*batch, buf->itab, buf->status, c.st));
    STAIR_TRY(run_encoders_train(c, *train));
    STAIR_TRY(run_modules(c));
    STAIR_TRY(run_decoder(c));
    t_last_launches = g_launch_count - before;
    return STAIR_OK;
}

extern "C" int stair_nmn_backward(const StairModel* model, const StairBatch* batch, const StairBuffers* buf, const StairTrain* train, void* stream) {
    return stair_nmn_backward_phases(model, batch, buf, train, STAIR_BWD_ALL, stream);
}

extern "C" int stair_nmn_backward_phases(const StairModel* model, const StairBatch* batch, const StairBuffers* buf, const StairTrain* train, int phases,
                                         void* stream) {
    if (!model || !batch || !buf || !train || !(phases & STAIR_BWD_ALL)) return STAIR_ERR_ARG;
    if (batch->B <= 0) return STAIR_OK;
    Ctx c{*model, *batch, *buf, reinterpret_cast<cudaStream_t>(stream)};
    STAIR_TRY(make_ctx(c, *model, *batch, *buf));
    if (buf->workspace_bytes < c.plan.total) return STAIR_ERR_CAPACITY;
    c.drop_p = train->dropout_p; c.drop_seed = train->dropout_seed;     // the recomputed forward of every chunk regenerates the masks
    if (!(c.drop_p >= 0.0f && c.drop_p < 1.0f)) return STAIR_ERR_ARG;
    STAIR_TRY(bind_saved_activations(c, *train));
    BCtx b{c, *train, Bump(), false};
    b.ws.base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(train->workspace) + 255) & ~static_cast<uintptr_t>(255));
    b.ws.cap = train->workspace_bytes - 256;
    const long long before = g_launch_count;
    const int rc = run_backward(b, phases);
    t_last_launches = g_launch_count - before;
    if (rc == STAIR_OK && b.ws.overflow) return STAIR_ERR_CAPACITY;
    return rc;
}

extern "C" int stair_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr, double beta1, double beta2,
                               float eps, int step, void* stream) {
    const float bc1 = static_cast<float>(1.0 - pow(beta1, step)), bc2 = static_cast<float>(1.0 - pow(beta2, step));
    return launch_adam(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, bc1, bc2, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int stair_set_bwd_lanes(int lanes) { g_bwd_lanes = lanes < 1 ? 1 : (lanes > LANES ? LANES : lanes); return STAIR_OK; }
extern "C" int stair_set_dw_impl(int impl) { g_dw_impl = impl ? 1 : 0; return STAIR_OK; }
extern "C" int stair_set_bptt_impl(int impl) { g_bptt_impl = impl ? 1 : 0; return STAIR_OK; }
extern "C" int stair_set_loss_con_impl(int impl) { g_loss_con_impl = impl ? 1 : 0; return STAIR_OK; }

extern "C" int stair_adam_multi(const StairAdamSeg* segs, int n_segs, int total_tiles, float lr, double beta1, double beta2, float eps, void* stream) {
    if (!segs && n_segs > 0) return STAIR_ERR_ARG;
    return launch_adam_multi(segs, n_segs, total_tiles, lr, beta1, beta2, eps, reinterpret_cast<cudaStream_t>(stream));
}

// Host-side evaluation of the dropout mask (no GPU work): lets a binding / test restate and verify the counter-based mask that the
// training kernels apply on the device (stair_common.cuh make_drop / drop_keep).
extern "C" int stair_dropout_mask_host(float p, unsigned long long seed, int site, long long row0, int rows, int cols, unsigned char* keep) {
    if (!keep || rows < 0 || cols < 0 || !(p >= 0.0f && p < 1.0f)) return STAIR_ERR_ARG;
    const DropSpec d = make_drop(p, seed, site, row0);
    for (int r = 0; r < rows; ++r) {
        const uint32_t rh = drop_row_hash(d.key_lo, d.key_hi, d.row0 + r);
        for (int c = 0; c < cols; ++c) keep[static_cast<long long>(r) * cols + c] = (!d.thresh || drop_keep(rh, static_cast<uint32_t>(c), d.thresh)) ? 1 : 0;
    }
    return STAIR_OK;
}
