// Backward / loss kernels of the NMN training step (train_kernels.cu), used by executor_bwd.cu.
// Gradients are fp32 everywhere; activations (`dt`) are bf16 or fp32 like in the forward.  Every kernel ACCUMULATES (+=) into
// its gradient outputs: the gradient arenas are zeroed once at the start of the backward pass.  "atomic" outputs may be hit by
// several instances (shared video slots, shared token rows, weight gradients).
#pragma once
#include "nmn_kernels.cuh"

namespace stair {

// dz = dY * yscale * (Y > 0 if Y)  ->  bf16 planes dZ [np][M, N_ld]; dZs = rs[m] * dz planes (only if rs, else dZs may be null);
// db[n] += sum_m dz (if db).  Columns N..N_ld are zero.
int launch_dz_prep(int ydt, const float* dY, long long ld_dy, const void* Y, long long ld_y, const float* rs, bf16* dZ, bf16* dZs,
                   long long n_ld, long long plane_rows, int nplanes, float* db, int M, int N, float yscale, cudaStream_t st);
// dst [np][C, ld_dst] = transpose(src [np][R, ld_src] (first C columns)); columns R..ld_dst zeroed
int launch_transpose_planes(const bf16* src, long long ld_src, long long src_plane_rows, bf16* dst, long long ld_dst, long long dst_plane_rows,
                            int nplanes, int R, int C, cudaStream_t st);
// dr[m] += <G[m,:], X[m,:]> ; G[m,:] *= rs[m]      (X rows optionally gathered like launch_stage_rows)
int launch_rowscale_bwd(int xdt, float* G, const void* X, const int* slots, int rps, int unit, const float* rs, float* dr, int M, int K, cudaStream_t st);
// dst[(idx[r / rps] * unit + r % rps) * H + :] += src[r, :]   (atomic)
int launch_scatter_add_rows(const float* src, const int* idx, int rps, int unit, float* dst, long long rows, int H, cudaStream_t st);
int launch_layernorm_bwd(int xdt, const float* dOut, const void* X, const float* gamma, float* dX, float* dgamma, float* dbeta, long long rows, int H, cudaStream_t st);
int launch_cos_att_bwd(int dt, const void* f, const void* kmat, int K, int T, int H, const float* datt, long long att_base, float* df, float* dk, int n, cudaStream_t st);
int launch_existsframe_bwd(int dt, const void* vid, const int* feat_idx, const void* vec, const int* kw_idx, const float* datt, int att_base,
                           float* dvid, float* dvec, int n, int T, int H, cudaStream_t st);
int launch_temporal_relate_bwd(const float* att, const int* att_idx, int K, int mode, int conv_k, const float* const* params, float* const* dparams,
                               const float* dr, float* datt, int n, int T, cudaStream_t st);
int launch_bcast_T(const float* dagg, float* dx, int n, int T, int H, cudaStream_t st);
int launch_ff_attn_bwd(int dt, const void* x, const void* vec, const int* kw_idx, const float* w, const float* a, const float* da, float* dx,
                       float* dvec, float* dw, float* db, int n, int T, int H, cudaStream_t st);
int launch_attnvideo_bwd(int dt, const float* dOut, const void* vid, const int* feat_idx, const float* att, const int* att_idx, float* datt,
                         float* dvid, int n, int T, int H, cudaStream_t st);
int launch_relate_bwd(const float* att_out, int out_base, const float* datt_out, const int* att_idx, int sign, float* datt, float* dbeta, int n, int T, cudaStream_t st);
int launch_rowdot_sigmoid_bwd(int dt, const void* x, const float* w, const float* a, const float* da, float* dx, float* dw, float* db,
                              long long rows, int H, float dscale, cudaStream_t st);
int launch_choose_bwd(int dt, const void* vec, const int* k1, const int* k2, const int* q, const float* dOut, float* dvec, int n, int H, cudaStream_t st);
int launch_binary_bwd(int dt, const void* base, const int* a_idx, const int* b_idx, const float* dOut, float* dbase, int unit, int len, int op, int n, cudaStream_t st);
int launch_array2_bwd(const float* dOut, const int* a_idx, const int* b_idx, float* dvec, int n, int H, cudaStream_t st);
int launch_concat_bwd(int dt, const void* vec, const int* a_idx, const int* b_idx, int mode, const float* dcat, float* dvec, int n, int H, cudaStream_t st);
int launch_super_mix_bwd(int dt, const float* att, int K, int T, int H, int is_min, const void* act_base, const int* act_idx, int act_unit,
                         const float* dv, float* datt_s, float* dact_base, int n, cudaStream_t st);
int launch_word_embed_bwd(const float* dvec, int out_base, const int* q_off, const int* pos_q, const int* span_s, const int* span_e,
                          float* dtokfeat, int n, int H, cudaStream_t st);
int launch_decoder_concat_bwd(const float* dcat, const int* root_node, const int* out_slot, float* dvec, float* dqfeat, int B, int H, cudaStream_t st);

// ---- losses (train_module.py:83-194); every kernel adds weight * loss to loss[slot] and writes the prediction gradients ----------
// attention_score_criterion rows: pred row = att[(kind < 2 ? out_slot[node] + kind : aux_slot[node])]; slot = 0 Localize, 1 Temporal, 2 ExistsFrame
int launch_loss_att(const float* att, float* datt, const int* out_slot, const int* aux_slot, const int* node, const int* kind, const int* slot,
                    const float* gold, const float* w, float* loss, int n, int T, cudaStream_t st);
// Exists/Xor: CE over the 2-way head; Equals: MSE on the 1-way head.  Includes the head Linear backward.
int launch_loss_bin(int dt, const void* vec, float* dvec, const int* out_slot, const int* node, const int* is_mse, const int* label, const float* w,
                    const float* const* head_w /*[3]: equals, xor, exists*/, const float* const* head_b, float* const* dhead_w, float* const* dhead_b,
                    const int* which /*0 equals,1 xor,2 exists*/, float* loss, int n, int H, cudaStream_t st);
// contrastive CE of the L2-normalised module output against all class text reps of the window (train_module.py:113-132,388-406)
int launch_loss_con(int dt, const void* vec, float* dvec, const int* out_slot, const int* node, const int* pos, const float* w,
                    const float* cls_rep, int n_cls, float* loss, int n, int H, cudaStream_t st);
extern int g_loss_con_impl;      // 0 = shared-memory kernel when n_cls <= 64, H <= 512 (product); 1 = register kernel always
// criterion_filterframe: BCELoss(softmax_O(head row), gold row) per (node, frame); writes d head (not accumulated) and adds to loss[7]
int launch_loss_ff(const float* head, float* dhead, const int* aux_slot, const int* node, const float* gold, const float* w, float* loss,
                   int n, int T, int O, cudaStream_t st);
// Backward of the pretrain heads for EXTERNAL gradient seeds (StairTrain.ext_*): Linear(H, nout <= 2) heads of Equals / Xor / Exists
// (dvec[row] += W^T g, dW += g x^T, db += g) and the L2Normalize heads of Filter / ToAction / Superlative (dvec[row] += (g - y (y.g)) / |x|).
// vec rows row_base .. row_base + n, seeds dout[(out_base + i) * 2 + o] / dout[(out_base + i) * H ..].
int launch_small_head_bwd(int dt, const void* vec, int row_base, const float* w, int nout, const float* dout, int out_base, float* dvec, float* dW,
                          float* db, int n, int H, cudaStream_t st);
int launch_l2norm_bwd(int dt, const void* vec, int row_base, const float* dout, int out_base, float* dvec, int n, int H, cudaStream_t st);
int launch_add_inplace(float* dst, const float* src, long long n, cudaStream_t st);     // dst += src (n a multiple of 4, 16-byte aligned)
int launch_loss_dec(const float* logits, const int* answer, float w, float* dlogits, float* loss, int B, int A, cudaStream_t st);

// Blocked BPTT history written by the fused recurrence kernel (lstm_fused.cu, HIST).  Not the gates themselves but the six per-unit
// COEFFICIENTS the backward multiplies by, as bf16 (12 bytes per hidden unit and step instead of 24 bytes of fp32 gates + c + c_prev;
// the transcendental part of the cell derivative is computed once, in the forward, where tanh(c) is already in a register):
//   A  = o (1 - tanh(c)^2)      dc_t  = dc_carry + dh A          Bi = g i (1 - i)       d pre_i = dc_t Bi
//   Bf = c_prev f (1 - f)       d pre_f = dc_t Bf                Bg = i (1 - g^2)       d pre_g = dc_t Bg
//   Bo = tanh(c) o (1 - o)      d pre_o = dh Bo                  F  = f                 dc_carry' = dc_t F
// Layout: [step][dir][32-row block][8-unit block][coefficient][row][8 units] — the writer's warp stores 512-byte contiguous pieces, a
// reader fetches 16 bytes per (row, coefficient).  RB = 32-row blocks per (step, direction) = ceil(B / 128) * 4.
constexpr int LSTM_NCOEF = 6;
enum { LSTM_CO_A = 0, LSTM_CO_BI, LSTM_CO_BF, LSTM_CO_BG, LSTM_CO_BO, LSTM_CO_F };
__host__ __device__ static inline long long lstm_hist_rb(int B) { return (static_cast<long long>(B) + 127) / 128 * 4; }
__host__ __device__ static inline long long lstm_hist_coef_off(int step, int d, int b, int coef, int u, int B, int h) {      // in bf16 elements
    return ((((static_cast<long long>(step) * 2 + d) * lstm_hist_rb(B) + (b >> 5)) * (h >> 3) + (u >> 3)) * LSTM_NCOEF + coef) * 256 + (b & 31) * 8 + (u & 7);
}

// ---- LSTM with history (training forward) and its backward ---------------------------------------------------------------------------
// gates_out [2][B][4h] (post-activation i,f,g,o), c_prev/c_out [2][B][h]; hstate planes [np][...] with plane stride hs_plane;
// inactive (finished) questions copy their state forward.  q_off == null: video (row b*T+t); else ragged text.
int launch_lstm_cell_train(int xdt, const void* xproj, const float* g, const float* c_prev, float* c_out, float* gates_out, bf16* hstate_out,
                           const bf16* hstate_prev, long long hs_plane, long long hs_dir, int nplanes, int odt, void* out, void* qfeat, const int* q_off,
                           int B, int T, int h, int step, cudaStream_t st);
// one BPTT step: dh = dout(row) + dh_rec (+ dqfeat at a question's last step); writes the gate pre-activation gradients as bf16 planes
// into the step's slice of the direction-major history (dg_planes + d*dg_dir + b*4h, plane stride dg_plane) and as fp32 into the dxproj
// row; updates dc in place.
// blocked = 0: gates / c_prev / c_cur point at step `step`'s row-major slices ([2][B][4h], [2][B][h]; c_prev at step-1's);
// blocked = 1: gates is the BASE of the blocked bf16 coefficient history (lstm_hist_coef_off); c_prev / c_cur are ignored.
int launch_lstm_cell_bwd(const float* gates, const float* c_prev, const float* c_cur, const float* dout, const float* dh_rec, const float* dqfeat,
                         float* dc, long long dg_dir, bf16* dg_planes, long long dg_plane, int nplanes, float* dxproj, bf16* dxproj_bf16, const int* q_off,
                         int B, int T, int h, int step, int last_step, int blocked, cudaStream_t st);     // dxproj_bf16 != null: bf16 rows instead of fp32
int launch_colsum_bf16(const bf16* x, long long rows, int cols, long long ld, float* db, cudaStream_t st);      // db[c] += sum_r x[r][c]

int launch_adam_multi(const StairAdamSeg* segs, int n_segs, int total_tiles, float lr, double b1, double b2, float eps, cudaStream_t st);
int launch_adam(float* p, const float* g, float* m, float* v, long long n, float lr, double b1, double b2, float eps, float bc1, float bc2, cudaStream_t st);

}  // namespace stair
