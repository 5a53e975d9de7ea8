// Internal launch API of the memory-bound NMN kernels (nmn_kernels.cu, lstm.cu, layout_group.cu) used by the
// executor (executor.cu).  Every function enqueues on `st` and returns a STAIR_* status; nothing allocates or
// synchronises.
//
// Arena conventions:  VID slot s -> vid + s*T*H ;  VEC row r -> vec + r*H ;  ATT row r -> att + r*T
// (fp32).  `dt` is the activation dtype of VID/VEC (STAIR_BF16 | STAIR_F32).  Index arrays (`*_idx`) are in the
// sorted (grouped) node order produced by layout_group.cu; outputs of a group are contiguous from `out_base`.
#pragma once
#include "stair_common.cuh"

namespace stair {

// ---- GEMM operand staging -------------------------------------------------------------------------------------
// dst (bf16, `nplanes` planes of `plane_rows` rows, pitch ld_dst) <- rows of src (`sdt`, pitch ld_src), optionally
// gathered: source row of destination row r is slots[r / rps] * unit + r % rps (VID slots: unit = rps = T; VEC rows:
// unit = 1).  Columns [cols, ld_dst) are zeroed.
int launch_stage_rows(int sdt, const void* src, long long ld_src, const int* slots, int rps, int unit, bf16* dst, long long ld_dst,
                      long long plane_rows, int nplanes, long long rows, int cols, cudaStream_t st);
#define STAIR_CAT_EXISTS 0   // [b | a | b*a]  with a = arg0 (keyword), b = arg1 (feat)      modules.py:158
#define STAIR_CAT_XOR 1      // [|a-b| | a | b]                                             modules.py:72
#define STAIR_CAT_PAIR 2     // [a | b]                                                     modules.py:21,37,119
int launch_concat_vec(int dt, const void* vec, const int* a_idx, const int* b_idx, int mode, bf16* dst, long long plane_rows,
                      int nplanes, int n, int H, cudaStream_t st);
// [root VEC row | question_feature]  (module_net.py:136)
int launch_decoder_concat(int dt, const void* vec, const int* root_node, const int* out_slot, const void* qfeat, bf16* dst, long long plane_rows,
                          int nplanes, int B, int H, cudaStream_t st);

// ---- module kernels ---------------------------------------------------------------------------------------------
int launch_word_embed(int dt, const void* tokfeat, const int* q_off, const int* pos_q, const int* span_s, const int* span_e,
                      void* vec, int out_base, int n, int H, cudaStream_t st);
// att[(out_base + i*K + k)*T + t] = (cos(f[i*T+t], kmat[i*K+k]) + 1) * 0.49
int launch_cos_att(int dt, const void* f, const void* kmat, int K, int T, int H, float* att, long long out_base, int n, cudaStream_t st);
int launch_existsframe(int dt, const void* vid, const int* feat_idx, const void* vec, const int* kw_idx, float* att, int out_base,
                       int n, int T, int H, cudaStream_t st);
// r[(aux_base+i)*T + t] = relate[mode](mean_k att[(att_idx[i]+k)*T + t]);  params = {w0,b0,w1,b1,w2,b2} (Linear [T,T] or conv [k])
int launch_temporal_relate(const float* att, const int* att_idx, int K, int mode, int conv_k, const float* const* params,
                           float* att_out, int aux_base, int n, int T, cudaStream_t st);
int launch_layernorm(int dt, const void* x, const float* gamma, const float* beta, void* out, long long rows, int H, cudaStream_t st);
int launch_sum_T(int dt, const void* x, void* out, int n, int T, int H, cudaStream_t st);
// a[i*T+t] = sigmoid(w[0:H].x[i*T+t] + w[H:2H].vec[kw_idx[i]] + b)       (FilterFrame.attention, modules.py:407-409)
int launch_ff_attn(int dt, const void* x, const void* vec, const int* kw_idx, const float* w, const float* b, float* a,
                   int n, int T, int H, cudaStream_t st);
int launch_attnvideo(int dt, void* vid, const int* feat_idx, const float* att, const int* att_idx, int out_base,
                     int n, int T, int H, cudaStream_t st);
int launch_relate(const float* att, const int* att_idx, const float* beta, int sign, float* att_out, int out_base, int n, int T, cudaStream_t st);
// att[(out_base+i)*T+t] = sigmoid(w.x[i*T+t] + b)                          (HasItem tail, modules.py:128-129)
int launch_rowdot_sigmoid(int dt, const void* x, const float* w, const float* b, float* att, int out_base, int n, int T, int H, cudaStream_t st);
// in-place dropout of an fp32 map: x[i] = keep(row0 + i, col 0) ? x[i] * scale : 0     (HasItem: Sigmoid -> Dropout, modules.py:129)
int launch_drop_rows(float* x, long long rows, DropSpec d, cudaStream_t st);
int launch_choose(int dt, void* vec, const int* k1, const int* k2, const int* q, int out_base, int n, int H, cudaStream_t st);
#define STAIR_BIN_MIN 0
#define STAIR_BIN_ABSDIFF 1
// len = elements per instance (H for VEC rows, K*T for ATT rows); base arrays index units of `unit` elements
int launch_binary(int dt, void* base, const int* a_idx, const int* b_idx, int out_base, int unit, int len, int op, int n, cudaStream_t st);
int launch_array2(int dt, void* vec, const int* a_idx, const int* b_idx, int out_base, int n, int H, cudaStream_t st);
// Superlative tail (modules.py:243-247): w = softmax_K(sum_t att[i][k][t]); min -> 1-w; dst[i] = sum_k w_k actions[i][k]
// actions rows: base + (act_idx[i]*act_unit + k) * H
int launch_super_mix(int dt, const float* att, int K, int T, int H, int is_min, const void* act_base, const int* act_idx, int act_unit,
                     void* dst, int n, cudaStream_t st);
int launch_small_head(int dt, const void* vec, int row_base, const float* w, const float* b, int nout, float* out, int out_base, int n, int H, cudaStream_t st);
int launch_l2norm(int dt, const void* vec, int row_base, float* out, int out_base, int n, int H, cudaStream_t st);
int launch_argmax(const float* logits, int* out, int rows, int cols, cudaStream_t st);
int launch_relate_scan(const float* att, int mode, float* out, int n, int T, cudaStream_t st);

// TMA-staged streaming versions (row_stream.cu) of the cosine maps and the HasItem tail; row_stream_ok says which shapes they cover
bool row_stream_ok(int dt, int K, int T, int H);
int launch_cos_stream(int dt, const void* f, const int* feat_idx, const void* kw, const int* kw_idx, int K, int T, int H, float* att,
                      long long out_base, int n, cudaStream_t st);
int launch_rowdot_stream(int dt, const void* x, const float* w, const float* b, float* att, long long out_base, int n, int T, int H, cudaStream_t st);
extern int g_row_stream;      // 0 (default): register-staged kernels only; 1: HasItem tail streams; 2: the cosine maps stream too (measured slower: their
                              // three reductions per 1 KB row are issue-bound, not bandwidth-bound)

// ---- LSTM cells (gate order i,f,g,o; nn.LSTM, video_nmn/module_net.py:39-47) -------------------------------------
// xproj [rows, 8h] = W_ih x + b for both directions (fwd gates | reverse gates); g [2][B, 4h] = W_hh h_prev (fp32);
// c [2][B,h] fp32; hstate bf16 [nplanes][2*B, h] (A operand of the next step; direction d at rows d*B..);
// video: out[(b*T+t)*2h + d*h + j].
int launch_lstm_cell_video(int xdt, const void* xproj, const float* g, float* c, bf16* hstate, int nplanes, int odt, void* out,
                           int B, int T, int h, int step, cudaStream_t st);
// text: ragged; direction 0 consumes token step, direction 1 token L-1-step; final h of both directions -> qfeat [B, 2h]
int launch_lstm_cell_text(int xdt, const void* xproj, const float* g, float* c, bf16* hstate, int nplanes, int odt, void* tokfeat,
                          void* qfeat, const int* q_off, int B, int h, int step, cudaStream_t st);

// fused persistent recurrence for both encoders and both directions (lstm_fused.cu; bf16 path, h in {64,128,192,256})
bool lstm_fused_ok(int precision, int h);
// training: BPTT history written by the fused kernel, index 0 = video encoder, 1 = text encoder (layouts: executor_bwd.cu saved_layout).
// hs must be zero-filled by the caller (rows of finished questions are not written).  gates[e] holds the bf16 coefficient history
// (train_kernels.cuh lstm_hist_coef_off); c[e] is unused (the running cell state stays in c_scratch).
struct LstmHist { float* gates[2]; float* c[2]; bf16* hs[2]; long long hs_dir[2]; };
int launch_lstm_fused(const void* xproj_v, void* vid_out, int T, const void* whh_v_f, const void* whh_v_r,
                      const void* xproj_t, void* tokfeat, void* qfeat, const int* q_off, int L_max, const void* whh_t_f,
                      const void* whh_t_r, float* c_scratch, int B, int h, int run_video, int run_text, int* err_flag, cudaStream_t st,
                      const LstmHist* hist = nullptr, const int* text_order = nullptr, const int* text_soff = nullptr);
// length-sorted text schedule of the inference recurrence: order [B] = question ids in descending length, soff [B+1] = token offsets in that
// order, tok_src [n_tok] = source token row of every sorted row (the staging gather of the text projection's A operand)
bool text_sort_ok(int L_max);
int launch_text_sort(const int* q_off, int B, int L_max, int* order, int* soff, int* tok_src, cudaStream_t st);

// weight-stationary cluster recurrence of ONE encoder (lstm_ws.cu; bf16 path, h = 256): W_hh resident in the shared memory of a 4-CTA
// cluster, h exchanged through `hx` (>= lstm_ws_hx_bytes(B) bytes).  q_off == nullptr: video (T = steps for every question); else text.
// coef_h / hs_h: the BPTT history of launch_lstm_fused (null = inference).
bool lstm_ws_ok(int precision, int h, int B);
long long lstm_ws_hx_bytes(int B);
int launch_lstm_ws(const void* xproj, void* out, void* final_h, const int* q_off, int steps, const void* whh_f, const void* whh_r,
                   float* c_scratch, void* hx, int B, int h, int* err_flag, cudaStream_t st, bf16* coef_h = nullptr, bf16* hs_h = nullptr);

// fused persistent BPTT of both encoders and directions (lstm_bptt.cu; bf16 path, after a fused forward with history).
// index 0 = video, 1 = text.  gates: the forward's blocked bf16 coefficient history (c unused); dout: fp32 [rows][2h] gradient of the encoder output (dout[1] is modified: dqfeat is added to each question's last-step rows); dxb: bf16
// [rows][8h] gate pre-activation gradients in token order (output); dc: scratch of 2 * ceil(B/64)*64 * h floats per encoder;
// whhT[2*e + dir]: transposed W_hh copies [h][4h] (StairModel.wt).
struct LstmBptt {
    const float* gates[2]; const float* c[2]; float* dout[2]; const float* dqfeat; bf16* dxb[2]; float* dc[2]; const void* whhT[4];
};
bool lstm_bptt_fused_ok(int precision, int h);
// text_order / text_soff: the length-sorted schedule the fused forward ran with (null = batch order): dxb[1] rows are then in schedule order
int launch_lstm_bptt_fused(const LstmBptt& a, int B, int h, int T, int L_max, const int* q_off, int* err_flag, cudaStream_t st,
                           const int* text_order = nullptr, const int* text_soff = nullptr);

// ---- layout grouping (layout_group.cu) ----------------------------------------------------------------------------
int launch_group_layouts(const StairBatch& b, int32_t* itab, int32_t* status, cudaStream_t st);

}  // namespace stair
