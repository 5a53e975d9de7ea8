// Internal launch API of the memory-bound NMN kernels (nmn_kernels.cu) used by the executor (executor.cu).
// Every function enqueues on `st` and returns a STAIR_* status; nothing allocates or synchronises.
//
// Arena conventions:  VID slot s -> vid + s*T*H ;  VEC row r -> vec + r*H ;
// ATT row r -> att + r*T (fp32).  `dt` is the activation dtype of VID/VEC (STAIR_BF16 | STAIR_F32).
#pragma once
#include "stair_common.cuh"

namespace stair {

// utility
int launch_cast_f32_to_bf16(const float* src, bf16* dst, long long rows, int cols, long long ld_src, long long ld_dst, cudaStream_t st);
int launch_split3(const float* src, long long ld_src, bf16* dst, long long plane_rows, long long ld_dst, long long rows, int cols, cudaStream_t st);
int launch_split3_from_bf16(const bf16* src, long long ld_src, bf16* dst, long long plane_rows, long long ld_dst, long long rows, int cols, cudaStream_t st);

// gathers / elementwise
int launch_gather_vid(int dt, const void* vid, const int* slot, void* dst, int n, int T, int H, cudaStream_t st);
int launch_word_embed(int dt, const void* tokfeat, const int* q_off, const int* inst_q, const int* span, void* vec, const int* out_row,
                      int n, int H, cudaStream_t st);
int launch_cos_att(int dt, const void* f, const void* kmat, int K, int T, int H, float* att, const int* out_row, int n, cudaStream_t st);
int launch_temporal_relate(const float* att, const int* arg_row, const int* arg_k, int mode, int conv_mode, int ksize,
                           const float* const* params /*[6]: w0,b0,w1,b1,w2,b2*/, float* r_out, int n, int T, cudaStream_t st);
int launch_layernorm(int dt, const void* x, const float* gamma, const float* beta, void* out, long long rows, int H, cudaStream_t st);
int launch_sum_T(int dt, const void* x, void* out, int n, int T, int H, cudaStream_t st);
int launch_ff_attn(int dt, void* x, const void* vec, const int* kw_row, const float* w, const float* b, int n, int T, int H, cudaStream_t st);
int launch_attnvideo(int dt, const void* vid, const int* feat_slot, const float* att, const int* att_row, void* vid_out, const int* out_slot,
                     int n, int T, int H, cudaStream_t st);
int launch_relate(const float* att, const int* arg_row, const float* beta, int sign, float* att_out, const int* out_row, int n, int T, cudaStream_t st);
int launch_rowdot_sigmoid(int dt, const void* x, const float* w, const float* b, float* att, const int* out_row, int n, int T, int H, cudaStream_t st);
int launch_existsframe(int dt, const void* vid, const int* feat_slot, const void* vec, const int* kw_row, float* att, const int* out_row,
                       int n, int T, int H, cudaStream_t st);
// concat modes
#define STAIR_CAT_EXISTS 0   // [b | a | b*a]  with a = arg0 (keyword), b = arg1 (feat)      modules.py:158
#define STAIR_CAT_XOR 1      // [|a-b| | a | b]                                             modules.py:72
#define STAIR_CAT_PAIR 2     // [a | b]                                                     modules.py:21,37,119
int launch_concat_vec(int dt, const void* vec, const int* a_row, const int* b_row, int mode, void* dst, int n, int H, cudaStream_t st);
int launch_choose(int dt, void* vec, const int* k1, const int* k2, const int* q, const int* out_row, int n, int H, cudaStream_t st);
#define STAIR_BIN_MIN 0
#define STAIR_BIN_ABSDIFF 1
int launch_binary_vec(int dt, void* vec, const int* a_row, const int* b_row, const int* out_row, int op, int n, int H, cudaStream_t st);
int launch_binary_att(float* att, const int* a_row, const int* b_row, const int* out_row, int op, int n, int T, cudaStream_t st);
int launch_array2(int dt, void* vec, const int* a_row, const int* b_row, const int* out_row, int n, int H, cudaStream_t st);
int launch_super_weights(int dt, const float* att_scratch, int K, int T, int H, int is_min, const void* actions_base, const int* actions_idx,
                         long long actions_stride /*elements between instances' index units*/, void* dst, int n, cudaStream_t st);
int launch_small_head(int dt, const void* vec, const int* row, const float* w, const float* b, int nout, float* out, int n, int H, cudaStream_t st);
int launch_l2norm(int dt, const void* vec, const int* row, float* out, int n, int H, cudaStream_t st);
int launch_decoder_concat(int dt, const void* vec, const int* root_row, const void* qfeat, void* dst, int B, int H, cudaStream_t st);
int launch_argmax(const float* logits, int* out, int B, int A, cudaStream_t st);
int launch_relate_scan(const float* att, int mode, float* out, int n, int T, cudaStream_t st);

// LSTM cells (gate order i,f,g,o; video_nmn/module_net.py:39-47)
int launch_lstm_cell_video(int dt, const void* xproj, const float* hw_f, const float* hw_r, float* c /*[2][B][h]*/, void* out /*[B,T,2h]*/,
                           int B, int T, int h, int step, cudaStream_t st);
int launch_lstm_cell_text(int dt, const void* xproj, const float* hw_f, const float* hw_r, float* c, void* tokfeat /*[sumL,2h]*/,
                          void* hstate /*[B,2h]*/, const int* q_off, int B, int h, int step, cudaStream_t st);

}  // namespace stair
