/* stair_b200 — C ABI of the B200 (sm_100a) implementation of STAIR's video_nmn ModuleNet hot path.
 *
 * The reference (yellow-binary-tree/STAIR) is pure Python/PyTorch and has no FFI layer: its boundary for this
 * path is the nn.Module surface of video_nmn/module_net.py:11-176 and video_nmn/modules.py:7-465.  This header
 * is the boundary *below* that surface: stair_b200/nmn.py (the drop-in `VideoNMN`) binds it with ctypes, and a
 * maintainer of the reference would bind exactly these entry points (INTEGRATION.md shows the stub).
 *
 * Conventions: every function returns 0 (STAIR_OK) or a negative STAIR_ERR_* code; nothing throws, exits,
 * allocates device memory or synchronises — the caller (PyTorch's caching allocator) owns every buffer and all
 * work is enqueued on the `stream` argument (a cudaStream_t passed as void*).  Pointers are device pointers
 * unless marked HOST, and must be 16-byte aligned (32-byte: the StairBuffers arenas vid / tokfeat / qfeat, the workspace and
 * StairTrain.saved, which the recurrence kernels access with 256-bit loads / stores).  sm_100a only; there is no CPU or library fallback.
 *
 * What the library itself owns (and nothing else): per calling thread, 7 non-blocking side streams + 104 events (one fork, one join per
 * lane, one completion event per module group) on which independent module groups run concurrently in dependency order (forked from and
 * joined back into `stream`, so a call stays stream-ordered for the caller), and one pinned host int that device-side protocol time-outs report into (stair_gemm_error_flag).  stair_init() creates them
 * for the calling thread on the current device (otherwise the first forward of the thread does), stair_shutdown() synchronises the
 * device and destroys them.  One process (thread) per GPU is the intended use; a thread that changes its current device must call
 * stair_shutdown() / stair_init() around the change.
 * The stair_set_* switches below are PROCESS-GLOBAL tuning / comparison knobs read at launch time (defaults = the product path; tests and
 * the profiles/ scripts flip them around single calls): two models in one process share them, they are not part of a model's state.
 * Entry points are re-entrant given distinct buffers; stair_last_launch_count is per thread.
 */
#ifndef STAIR_B200_H
#define STAIR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STAIR_ABI_VERSION 9
#define STAIR_MAX_GROUP_DEPS 8

/* status codes */
#define STAIR_OK 0
#define STAIR_ERR_ARG (-1)
#define STAIR_ERR_CUDA (-2)
#define STAIR_ERR_CAPACITY (-3)
#define STAIR_ERR_LAYOUT (-4)
#define STAIR_ERR_UNSUPPORTED (-5)

/* element types of activations / inputs */
#define STAIR_BF16 0
#define STAIR_F32 1

#define STAIR_ACT_NONE 0
#define STAIR_ACT_RELU 1

/* Operator codes.  1..18 follow NAME_TO_MODULE order (video_nmn/modules.py:446-465); 0 is a content word
 * (phrase embedding, video_nmn/module_net.py:126-131). */
enum StairOp {
    STAIR_OP_WORD = 0, STAIR_OP_AND, STAIR_OP_ATTNVIDEO, STAIR_OP_CHOOSE, STAIR_OP_COMPARE, STAIR_OP_EQUALS,
    STAIR_OP_EXISTS, STAIR_OP_EXISTSFRAME, STAIR_OP_FILTER, STAIR_OP_FILTERFRAME, STAIR_OP_HASITEM,
    STAIR_OP_LOCALIZE, STAIR_OP_RELATE, STAIR_OP_SUPERLATIVE, STAIR_OP_TEMPORAL, STAIR_OP_TOACTION, STAIR_OP_XOR,
    STAIR_OP_XORFRAME, STAIR_OP_ARRAY2, STAIR_OP_COUNT
};

/* Weight table slots (StairModel.w[]).  *_W matrices are bf16 [nplanes][N, K_ld] (nplanes = 1, or 3 bf16 planes
 * w = w0+w1+w2 in strict fp32 mode), row-major with K contiguous exactly like nn.Linear.weight; everything else
 * is fp32.  Names follow the reference state_dict keys. */
enum StairWeight {
    /* video_encoder / text_encoder: nn.LSTM(bidirectional), module_net.py:39-47.  WIH = [fwd ; reverse] rows (2*4h),
     * B = b_ih + b_hh for both directions (2*4h). */
    STAIR_W_VENC_WIH = 0, STAIR_W_VENC_B, STAIR_W_VENC_WHH_F, STAIR_W_VENC_WHH_R,
    STAIR_W_TENC_WIH, STAIR_W_TENC_B, STAIR_W_TENC_WHH_F, STAIR_W_TENC_WHH_R,
    /* decoder.{0,3}, module_net.py:49-53 */
    STAIR_W_DEC0_W, STAIR_W_DEC0_B, STAIR_W_DEC1_W, STAIR_W_DEC1_B,
    /* Localize.video_linear.{0,3}, keyword_linear.0 (modules.py:185-192); shared by Superlative (module_net.py:31-32) */
    STAIR_W_LOC_V0_W, STAIR_W_LOC_V0_B, STAIR_W_LOC_V1_W, STAIR_W_LOC_V1_B, STAIR_W_LOC_K_W, STAIR_W_LOC_K_B,
    /* Temporal.dense.0, layer_norm, relate.{before,after,between}.{0,2,4} (modules.py:255-283) — relate params fp32 */
    STAIR_W_TEMP_D_W, STAIR_W_TEMP_D_B, STAIR_W_TEMP_LN_G, STAIR_W_TEMP_LN_B,
    STAIR_W_TEMP_REL_BEFORE,               /* 6 consecutive slots per mode: w0,b0,w1,b1,w2,b2 */
    STAIR_W_TEMP_REL_AFTER = STAIR_W_TEMP_REL_BEFORE + 6,
    STAIR_W_TEMP_REL_BETWEEN = STAIR_W_TEMP_REL_AFTER + 6,
    /* Filter.param.{representation,actions,objects,relations}.{0,3}, dense.0 (modules.py:346-358):
     * 4 consecutive slots per keyword kind: w0,b0,w1,b1 */
    STAIR_W_FILT_REPR = STAIR_W_TEMP_REL_BETWEEN + 6,
    STAIR_W_FILT_ACTIONS = STAIR_W_FILT_REPR + 4,
    STAIR_W_FILT_OBJECTS = STAIR_W_FILT_ACTIONS + 4,
    STAIR_W_FILT_RELATIONS = STAIR_W_FILT_OBJECTS + 4,
    STAIR_W_FILT_D_W = STAIR_W_FILT_RELATIONS + 4, STAIR_W_FILT_D_B,
    /* FilterFrame.param.{representation,relations,actions}, attention.0, dense.0, pretrain_head (modules.py:384-396) */
    STAIR_W_FF_REPR, STAIR_W_FF_RELATIONS = STAIR_W_FF_REPR + 4, STAIR_W_FF_ACTIONS = STAIR_W_FF_RELATIONS + 4,
    STAIR_W_FF_ATT_W = STAIR_W_FF_ACTIONS + 4, STAIR_W_FF_ATT_B, STAIR_W_FF_D_W, STAIR_W_FF_D_B,
    STAIR_W_FF_HEAD_W, STAIR_W_FF_HEAD_B,
    /* HasItem.param.{0,3} (modules.py:126-129); param.3 (Linear(H,1)) is fp32 */
    STAIR_W_HAS0_W, STAIR_W_HAS0_B, STAIR_W_HAS1_W, STAIR_W_HAS1_B,
    STAIR_W_REL_BETA,                      /* Relate.beta (modules.py:420) */
    STAIR_W_SUP_D_W, STAIR_W_SUP_D_B,      /* Superlative.dense.0 (modules.py:228-230) */
    STAIR_W_COMPARE_W, STAIR_W_COMPARE_B,  /* Compare.param.0 (modules.py:18) */
    STAIR_W_EQUALS_W, STAIR_W_EQUALS_B, STAIR_W_EQUALS_HEAD_W, STAIR_W_EQUALS_HEAD_B,    /* modules.py:27-29; heads fp32 */
    STAIR_W_XOR_W, STAIR_W_XOR_B, STAIR_W_XOR_HEAD_W, STAIR_W_XOR_HEAD_B,                /* modules.py:62-64 */
    STAIR_W_EXISTS0_W, STAIR_W_EXISTS0_B, STAIR_W_EXISTS1_W, STAIR_W_EXISTS1_B,
    STAIR_W_EXISTS_HEAD_W, STAIR_W_EXISTS_HEAD_B,                                        /* modules.py:144-150 */
    STAIR_W_TOACT0_W, STAIR_W_TOACT0_B, STAIR_W_TOACT1_W, STAIR_W_TOACT1_B,              /* modules.py:105-108 */
    /* W_hh of the two encoders with rows re-ordered for the fused recurrence kernel (bf16, 1 plane, only when h % 64 == 0):
     * row c*256 + g*64 + u  <-  W_hh row g*h + c*64 + u   (chunk c of 64 hidden units, gate g in i,f,g,o) */
    STAIR_W_VENC_WHHI_F, STAIR_W_VENC_WHHI_R, STAIR_W_TENC_WHHI_F, STAIR_W_TENC_WHHI_R,
    STAIR_W_COUNT
};

/* Static model description (config dict of train_module.py:304-310 + packed weights). */
typedef struct StairModel {
    int32_t T_max;       /* config['max_video_length']; Temporal.relate is Linear(T_max,T_max) iff T_max <= 32 */
    int32_t V, V_ld;     /* config['video_size'] and the row pitch of the packed W_ih (multiple of 8) */
    int32_t H;           /* config['hidden_size'] (multiple of 16); LSTM hidden per direction is H/2 */
    int32_t text_size, text_ld;
    int32_t A;           /* config['answer_vocab_length'] */
    int32_t O;           /* config['object_types'] */
    int32_t conv_k;      /* Temporal Conv1d kernel size round(T_max/4) when T_max > 32, else 0 */
    int32_t precision;   /* STAIR_BF16: bf16 storage, fp32 accumulate.  STAIR_F32: fp32 storage, bf16x3 split GEMMs */
    const void* w[STAIR_W_COUNT];
    const void* wt[STAIR_W_COUNT];  /* training only (else NULL): transposed copies bf16 [nplanes][K, N_ld] of the *_W GEMM matrices */
} StairModel;

/* One (level, op, variant) group of module instances; groups are listed in schedule order (level-major). */
typedef struct StairGroup {
    int32_t op, variant, level;
    int32_t count;       /* instances in the batch */
    int32_t node_off;    /* first position of the group in the sorted node order (exclusive prefix of counts) */
    int32_t out_base;    /* first output index in the op's arena (VID slot / VEC row / ATT row) */
    int32_t out_mult;    /* arena units per instance (K rows for Localize, 2 for Array2, else 1) */
    int32_t aux_base;    /* first auxiliary index: Temporal -> ATT row of the stashed related_attn; modules with a
                            pretrain head -> row in the head buffer; -1 if unused */
    int32_t head;        /* 1: also compute the module's pretrain_head for every instance (res_by_step / audit) */
} StairGroup;

/* One batch of questions, compiled by the host (stair_b200/layout.py) from reference-schema `data` dicts
 * (video_nmn/dataset.py:189-233).  Node order is question-major, token order within a question. */
typedef struct StairBatch {
    int32_t B;                 /* questions */
    int32_t T;                 /* frames per question (uniform within a batch) */
    int32_t n_tok;             /* sum of question lengths */
    int32_t L_max;             /* longest question */
    int32_t n_nodes;
    int32_t n_groups;
    int32_t video_dtype;       /* STAIR_F32 | STAIR_BF16 */
    int32_t question_dtype;
    const void* video;         /* [B*T, V] (pitch V, must be a multiple of 8 elements when bf16) */
    const void* question;      /* [n_tok, text_size] packed word embeddings */
    const int32_t* q_off;      /* [B+1] token offsets */
    const int32_t* node_gid;   /* [n_nodes] group id of every node */
    const int32_t* node_q;     /* [n_nodes] question index */
    const int32_t* node_arg;   /* [3][n_nodes] child node index | -1 unused | -2 'video' (encoded frames of own question) */
    const int32_t* node_span;  /* [2][n_nodes] word span of content-word nodes; start = -1 => whole question */
    const int32_t* root_node;  /* [B] node index of token 0 */
    const StairGroup* groups;  /* HOST [n_groups] */
    const int32_t* group_tab;  /* device [4][n_groups]: node_off, out_base, out_mult, aux_base (same as `groups`) */
    const int32_t* group_deps; /* HOST [n_groups][STAIR_MAX_GROUP_DEPS] or NULL: the groups whose outputs group g reads (-1 = unused entry; a first
                                * entry of -2 = "every earlier group").  With it the module phase is scheduled by data dependency (a group starts as
                                * soon as its producers are done) instead of wave by wave. */
    /* Optional length-sorted schedule of the text recurrence — inference and, with the fused forward + persistent BPTT, the training
     * step, whose forward and backward calls must then see the same three arrays (all three or none; stair_b200.layout.collate fills them):
     * q_order [B] = the question ids in descending question length (a permutation of 0..B-1), q_soff [B+1] = token offsets in THAT order,
     * tok_src [n_tok] = the token row (batch order) of every sorted row: tok_src[q_soff[p] + s] = q_off[q_order[p]] + s.  Scheduling data
     * only: the outputs keep the batch's order and do not depend on it (training: gradients equal up to fp32 summation order).  NULL:
     * inference sorts on the device itself (two small kernels in front of the text staging) unless stair_set_text_sort(0); training
     * keeps batch order. */
    const int32_t* q_order; const int32_t* q_soff; const int32_t* tok_src;
} StairBatch;

/* Caller-owned output / scratch buffers. */
typedef struct StairBuffers {
    void* vid;  int64_t vid_slots;    /* [vid_slots][T][H] act dtype; slots 0..B-1 = encoded video (module_net.py:160-163) */
    void* vec;  int64_t vec_rows;     /* [vec_rows][H] act dtype */
    float* att; int64_t att_rows;     /* [att_rows][T] fp32 attention maps */
    void* tokfeat;                    /* [n_tok][H] act dtype (token_feature) */
    void* qfeat;                      /* [B][H] act dtype (question_feature) */
    float* logits;                    /* [B][A] */
    int32_t* answers;                 /* [B] argmax(logits) */
    float* head_small;                /* [head_small_rows][2] Equals/Xor/Exists pretrain heads */
    float* head_vec;                  /* [head_vec_rows][H] L2-normalised Filter/Superlative/ToAction heads */
    float* head_ff;                   /* [head_ff_rows][T][O] FilterFrame head */
    int32_t* itab; int64_t itab_ints; /* device int workspace, >= stair_itab_ints(n_nodes, n_groups) */
    void* workspace; int64_t workspace_bytes;   /* >= stair_nmn_workspace_bytes(...) */
    int32_t* status;                  /* device int[4]: [0] != 0 => layout grouping mismatch / device-side error code */
} StairBuffers;

/* Offsets (in int32 units) of the device tables the grouping step leaves in StairBuffers.itab. */
typedef struct StairItabLayout {
    int64_t perm;        /* [n_nodes] sorted position -> node */
    int64_t out_slot;    /* [n_nodes] node -> output index in its arena */
    int64_t aux_slot;    /* [n_nodes] node -> auxiliary index */
    int64_t arg_slot;    /* [3][n_nodes] sorted position -> resolved argument index */
    int64_t pos_q;       /* [n_nodes] sorted position -> question */
    int64_t pos_span;    /* [2][n_nodes] sorted position -> word span */
    int64_t group_off;   /* [n_groups+1] device-computed group offsets */
    int64_t total;
} StairItabLayout;

/* Training step state (train_module.py:33-194 losses, :341-412 window logic).  All pointers are device
 * pointers owned by the caller.  Loss rows are collated by the host (stair_b200/train.py) with the reference's inclusion rules;
 * every weight `w` already contains module_loss_weight (or decoder_loss_weight) / gradient_accumulation / #elements. */
typedef struct StairTrain {
    float* grad[STAIR_W_COUNT];   /* fp32 gradient accumulators, logical shapes of the parameters ([N,K] / vectors); NULL = frozen */
    /* attention_score_criterion rows (Localize / Temporal / ExistsFrame), train_module.py:83-90,157-191 */
    int32_t n_att; const int32_t* att_node; const int32_t* att_kind; const int32_t* att_slot; const float* att_gold; const float* att_w;
    /* Exists / Xor (CE on the 2-way head) and Equals (MSE on the 1-way head), train_module.py:92-107 */
    int32_t n_bin; const int32_t* bin_node; const int32_t* bin_which; const int32_t* bin_label; const float* bin_w;
    /* contrastive CE of Filter / ToAction / Superlative against the window's class text reps, train_module.py:113-139,166-171,388-406 */
    int32_t n_con; const int32_t* con_node; const int32_t* con_pos; const float* con_w; int32_t n_cls; const float* cls_rep;
    /* decoder CE, train_module.py:193-194,376-380 */
    const int32_t* answer; float dec_w;
    float* loss;                  /* float[8] sums: 0 Localize 1 Temporal 2 ExistsFrame 3 Exists/Xor 4 Equals 5 contrastive 6 decoder 7 FilterFrame */
    float* dvid; float* dvec; float* datt; float* dtokfeat; float* dqfeat; float* dlogits;   /* gradient arenas (same shapes as the forward arenas) */
    void* saved; int64_t saved_bytes;           /* BPTT history written by stair_nmn_forward_train (opaque: fp32 gates / cell / hidden state for the
                                                 * step-wise fp32-strict path, bf16 cell-derivative coefficients + token-order hidden states for the
                                                 * fused bf16 path); size from stair_train_saved_bytes */
    void* workspace; int64_t workspace_bytes;   /* backward scratch, >= stair_train_workspace_bytes */
    /* nn.Dropout(p) of the reference's training mode (video_nmn/args.py:31 default 0.25; sites: modules.py Linear->ReLU->Dropout
     * of Exists/Filter/FilterFrame/HasItem/Localize/Temporal/ToAction, HasItem's Sigmoid->Dropout, decoder module_net.py:49-53).
     * Counter-based masks keyed by (dropout_seed, site, global row, column); stair_nmn_backward must receive the same p and seed
     * as the stair_nmn_forward_train it follows.  0 = no dropout (eval semantics). */
    float dropout_p; uint64_t dropout_seed;
    /* optional: >= stair_train_act_bytes bytes in which stair_nmn_forward_train keeps every group chunk's module intermediates, so that
     * stair_nmn_backward reads them back instead of re-running the chunk's forward (NULL = recompute; HBM is 180 GB, a 4096-question
     * window needs ~3 GB) */
    void* act_saved; int64_t act_saved_bytes;
    /* criterion_filterframe (train_module.py:141-155; excluded from training by default, video_nmn/args.py:62): BCELoss(softmax_O(head),
     * gold / rowsum) on the [T, O] head of supervised FilterFrame nodes.  ff_gold [n_ff][T][O] is the normalised gold, ff_w the weight
     * per element (module_loss_weight / (ga T O)); dhead_ff has the shape of StairBuffers.head_ff.  loss[7] receives the sum. */
    int32_t n_ff; const int32_t* ff_node; const float* ff_gold; const float* ff_w; float* dhead_ff; int64_t dhead_ff_elems;
    /* External gradient seeds (all NULL = the built-in criteria above).  When `ext_dlogits` is set the backward skips its own losses and
     * starts from the caller's gradients with respect to what the forward exposed — the reference's `loss.backward()` on `logits` and on the
     * `res_by_step` tensors (module_net.py:107-113,140-145) with ANY torch loss (stair_b200/train.py DifferentiableNMN binds this as a
     * torch.autograd.Function):  ext_dlogits [B, A];  ext_datt [att_rows, T] (Localize / ExistsFrame maps and Temporal's stashed gate live in
     * the ATT arena);  ext_dhead_small [small rows, 2] (Linear heads of Exists / Xor / Equals);  ext_dhead_vec [hvec rows, H] (L2Normalize
     * heads of Filter / ToAction / Superlative);  ext_dhead_ff [ff rows, T, O] (FilterFrame head).  fp32, shapes of the forward buffers;
     * any of the last four may be NULL (= zero). */
    const float* ext_dlogits; const float* ext_datt; const float* ext_dhead_small; const float* ext_dhead_vec; const float* ext_dhead_ff;
} StairTrain;

/* Host evaluation (no GPU work) of the dropout mask of site `site` (a STAIR_W_* id of the Linear the Dropout follows) for the
 * elements (row0 + r, c), r < rows, c < cols: keep[r*cols + c] = 1 if kept.  Restated in oracle/nmn_oracle.py dropout_keep. */
int stair_dropout_mask_host(float p, unsigned long long seed, int site, long long row0, int rows, int cols, unsigned char* keep);
int stair_version(void);
int stair_init(void);        /* create the calling thread's lane streams / events and the pinned error word now (idempotent) */
int stair_shutdown(void);    /* cudaDeviceSynchronize, then destroy them (a later call re-creates them lazily) */
/* sizeof() of the ABI structs as compiled (0 StairModel, 1 StairGroup, 2 StairBatch, 3 StairBuffers, 4 StairItabLayout, 5 StairTrain):
 * 6 StairAdamSeg; lets a binding verify its mirror of the struct layouts. */
int64_t stair_sizeof(int which);
/* concurrency of the module phase: independent groups run on up to `lanes` (1..8, default 8) internal streams
 * forked from and joined back into the caller's stream (the call stays stream-ordered for the caller) */
int stair_set_lanes(int lanes);
/* module-phase scheduling: 1 (default) = by data dependency when StairBatch.group_deps is given (per-group events, no barrier between the
 * schedule waves); 0 = wave by wave (fork / join around every wave) */
/* 1 (default): the inference text recurrence runs over the questions in descending length (StairBatch.q_order / q_soff / tok_src, or a
 * counting sort on the device when those are NULL): the text projection's input is staged in that order, every 64-question block stops at
 * its own longest question and the longest blocks start first.  Outputs (token_feature / question_feature rows) keep the batch's order
 * and are bit-identical to 0 = batch order. */
int stair_set_text_sort(int on);
int stair_set_fuse_sum(int on);             /* 1: Filter's sum over frames runs in the epilogue of its second Linear (inference); default 0 (measured no faster) */
int stair_set_dep_sched(int on);
/* debug: timing events around every module group of the dependency-scheduled phase; read returns the number of groups (synchronises) */
int stair_debug_timeline(int on);
/* with stair_debug_timeline(1): ms from the start of the last forward to [1] video projection, [2] text projection, [3] recurrence, [4] grouping
 * join, [5] module phase, [6] decoder complete on the caller's stream (ms[0] = 0); returns the number of marks written (synchronises) */
int stair_debug_phase_marks(float* ms, int cap);
int stair_debug_timeline_read(float* t0_ms, float* t1_ms, int* lane, int* op, int* count, int* variant, int cap);
/* encoder recurrence implementation: 0 = fused persistent kernel when eligible (default), 1 = per-step GEMM + cell kernels */
int stair_set_lstm_impl(int impl);
/* bf16 recurrence at h = 256: 1 = weight-stationary cluster kernel (csrc/lstm_ws.cu: W_hh resident in the shared memory of a 4-CTA cluster,
 * 128 questions per block; bit-identical, measured no faster); 0 (default; env STAIR_LSTM_WS overrides) = the streaming kernel of
 * csrc/lstm_fused.cu (64 questions per CTA, W_hh re-read from L2 every step) */
int stair_set_lstm_ws(int on);

/* ---- dense contraction (tcgen05 + TMA): C[M,N] = act(row_scale[m] * (A[M,K] . W[N,K]^T) + bias[n]) -------------
 * Replaces every nn.Linear / LSTM projection call site (video_nmn/modules.py passim, module_net.py:39-53). */
int stair_gemm_bf16(const void* A, long long lda, int a_plane_rows, const void* W, long long ldw, int w_plane_rows,
                    int nplanes, const float* bias, const float* row_scale, void* C, long long ldc, int out_dtype,
                    int M, int N, int K, int act, int accumulate, void* stream);
/* Same contraction with A gathered from an arena of slots: row m = slot a_slots[m / slot_rows], frame m % slot_rows.
 * Used for every module that reads [T,H] frame features of arbitrary questions (TMA 3-D boxes, no staging copy). */
/* out[i, :] (bf16 [M / T, ld_sum]) = sum over the T consecutive rows of instance i of bf16(act(A . W^T + bias)): the frame sum of Filter
 * (video_nmn/modules.py:374-376) taken in the GEMM epilogue; the [M, N] product itself is not written.  T in {1,2,4,...,128}, M % T == 0. */
int stair_gemm_bf16_framesum(const void* A, long long lda, const void* W, long long ldw, const float* bias, void* sum_out, long long ld_sum,
                             int M, int N, int K, int act, int T, void* stream);
int stair_gemm_bf16_gather(const void* arena, long long ld, long long arena_slots, const int32_t* a_slots, int slot_rows,
                           const void* W, long long ldw, const float* bias, const float* row_scale, void* C, long long ldc,
                           int out_dtype, int M, int N, int K, int act, void* stream);
/* Weight-gradient contraction C[M,N] (+)= A^T . W over K rows: A bf16 [nplanes][a_plane_rows, lda] with element (k, m) at k*lda + m,
 * W bf16 [nplanes][w_plane_rows, ldw] with element (k, n); operands are consumed in place as MN-major tensor-core tiles
 * (dW = dZ^T . X of every nn.Linear backward, train_module.py:408).  C fp32; accumulate = 1 adds into C (split-K with atomics). */
int stair_gemm_bf16_tn(const void* A, long long lda, int a_plane_rows, const void* W, long long ldw, int w_plane_rows, int nplanes,
                       float* C, long long ldc, int M, int N, int K, int accumulate, void* stream);
int stair_set_bwd_lanes(int lanes);     /* module backward: groups of one schedule wave on up to `lanes` concurrent streams (1 = sequential) */
int stair_set_dw_impl(int impl);        /* weight gradients: 0 = MN-major operands in place (product); 1 = transposed copies + K-major GEMM */
int stair_set_loss_con_impl(int impl);  /* contrastive loss: 0 = shared-memory kernel for windows of <= 64 classes (product); 1 = register kernel always */
int stair_set_bptt_impl(int impl);      /* encoder BPTT (bf16 path): 0 = one persistent fused kernel for both encoders / directions (product); 1 = per-step cell kernel + recurrent GEMMs */
int stair_set_gemm_impl(int impl);      /* 0 = tcgen05 (product); 1 = SIMT debug kernel used to cross-check in tests */
int stair_get_gemm_impl(void);
int stair_set_gemm_wide_min(int half_waves); /* 128 x 256 tiles when a GEMM has more than half_waves * SMs / 2 tiles of 128 x 128 (default 1; 4 = the round-1 rule of two full waves) */
int stair_set_gemm_epi2(int on);         /* 1 (default) = two epilogue warp sets (384 threads) for GEMMs with <= 8 k-blocks; 0 = one set always */
int stair_set_gemm_split_k(int on);      /* 1 (default) = split-K with atomic accumulation for accumulating GEMMs with few output tiles */
int stair_set_gemm_pair(int mode);       /* CTA-pair (tcgen05 cta_group::2, 256 x 256 tiles per 2-CTA cluster) GEMM: 1 (default) = for K-major, non-gather, non-accumulating GEMMs with N % 256 == 0 and at least a quarter wave of tiles; 0 = never; 2 = whenever legal (tests) */
int stair_set_gemm_pair_gather(int on);     /* 1 (default): gathered-A GEMMs (frame-arena slots) may use the CTA-pair kernel */
int stair_set_gemm_pair_mn(int on);      /* 1 (default) = the CTA-pair kernel also runs the MN-major weight-gradient contractions (split-K over the pairs); 0 = K-major GEMMs only */
int stair_set_gemm_epilogue(int impl);  /* 0 = smem-staged TMA-store epilogue (product); 1 = direct per-row stores (comparison) */
int stair_gemm_debug_timeline(unsigned long long* dev_buf /* 8 x u64, or NULL to disable */);
int stair_gemm_error_flag(void);

/* ---- layout grouping (utils/program_parser.py:182-200,307-321 semantics, grouped on device) --------------------
 * Stable counting sort of the batch's nodes by group id + output-slot assignment + argument resolution. */
int64_t stair_itab_ints(int32_t n_nodes, int32_t n_groups);
int stair_itab_layout(int32_t n_nodes, int32_t n_groups, StairItabLayout* out /*HOST*/);
int stair_group_layouts(const StairBatch* batch /*HOST struct*/, int32_t* itab, int32_t* status, void* stream);

/* ---- whole forward: VideoNMN.forward (video_nmn/module_net.py:65-145) for a batch ------------------------------ */
int64_t stair_nmn_workspace_bytes(const StairModel* model /*HOST*/, const StairBatch* batch /*HOST*/);
#define STAIR_FWD_ENCODE_VIDEO 1 /* BiLSTM video encoder (module_net.py:160-163) */
#define STAIR_FWD_ENCODE_TEXT 2  /* BiLSTM text encoder (module_net.py:147-158) */
#define STAIR_FWD_GROUP 4        /* device layout grouping */
#define STAIR_FWD_MODULES 8      /* grouped module execution */
#define STAIR_FWD_DECODE 16      /* decoder + argmax (module_net.py:135-138, train_module.py:252) */
#define STAIR_FWD_ALL 31
int stair_nmn_forward(const StairModel* model /*HOST*/, const StairBatch* batch /*HOST*/, const StairBuffers* buf /*HOST*/,
                      int phases, void* stream);
/* Training: forward that keeps the encoder history, then losses + backward into StairTrain.grad (gradients ACCUMULATE; the
 * caller zeroes them).  Module intermediates are recomputed per group in the backward pass instead of being stored. */
int64_t stair_train_saved_bytes(const StairModel* model /*HOST*/, const StairBatch* batch /*HOST*/);
int64_t stair_train_act_bytes(const StairModel* model /*HOST*/, const StairBatch* batch /*HOST*/, const StairBuffers* buf /*HOST*/);
int64_t stair_train_workspace_bytes(const StairModel* model /*HOST*/, const StairBatch* batch /*HOST*/, const StairBuffers* buf /*HOST*/,
                                    const StairTrain* train /*HOST*/);
int stair_nmn_forward_train(const StairModel* model, const StairBatch* batch, const StairBuffers* buf, const StairTrain* train, void* stream);
int stair_nmn_backward(const StairModel* model, const StairBatch* batch, const StairBuffers* buf, const StairTrain* train, void* stream);
/* The same backward in two enqueue steps (train_module.py:408 is one loss.backward()): STAIR_BWD_MODULES = losses + decoder + module
 * groups — afterwards every StairTrain.grad slot except the encoders' (STAIR_W_VENC_* / STAIR_W_TENC_*) is final —, STAIR_BWD_ENCODERS =
 * BPTT + encoder weight gradients.  A data-parallel caller all-reduces the module gradients while the encoders back-propagate. */
#define STAIR_BWD_MODULES 1
#define STAIR_BWD_ENCODERS 2
#define STAIR_BWD_ALL 3
int stair_nmn_backward_phases(const StairModel* model, const StairBatch* batch, const StairBuffers* buf, const StairTrain* train, int phases,
                              void* stream);
/* ---- one operator outside the interpreter: the per-class forward(*params) of video_nmn/modules.py:7-465 for group->count instances ----
 * The caller places the operands in the arenas of `buf` (VID slots / VEC rows / ATT rows; tokfeat / qfeat / logits / answers / itab unused) and
 * passes their indices: args[k * count + i] = arena index of argument k (< 3) of instance i, in the argument order the layout compiler resolves
 * (stair_b200/layout.py Layout._resolve: keyword strings are folded into group->variant, so e.g. Temporal = {feat VID slot, attention ATT row},
 * ExistsFrame = {keyword VEC row, feat VID slot}).  Output i lands at group->out_base + i * group->out_mult of the op's arena, aux / head rows as
 * in StairGroup.  group->node_off is ignored.  Runs the same group code as stair_nmn_forward (no separate kernels).  buf->workspace needs
 * stair_op_workspace_bytes bytes. */
int64_t stair_op_workspace_bytes(const StairModel* model /*HOST*/, int T, const StairGroup* group /*HOST*/);
int stair_op_forward(const StairModel* model /*HOST*/, int T, const StairGroup* group /*HOST*/, const int32_t* args /*DEVICE [3][count]*/,
                     const StairBuffers* buf /*HOST*/, void* stream);
/* torch.optim.Adam step (weight_decay 0) on one parameter tensor; `step` counts from 1 (train_module.py:326-332,408-412) */
int stair_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr, double beta1, double beta2,
                    float eps, int step, void* stream);      /* betas are doubles: 1 - beta is formed in double like torch does */
/* Multi-tensor Adam fused with the refresh of the kernels' weight copies: ONE launch updates every listed parameter (fp32 master,
 * exp_avg, exp_avg_sq; same arithmetic as stair_adam_step) and rewrites what the forward / backward kernels read — the bf16 plane
 * copy [nplanes][rows, ld], its transposed copy [cols, ld_t] (B operand of dX = dZ.W), the gate-interleaved W_hh copy of the fused
 * recurrence, or the fp32 vector copy.  Replaces one optimizer kernel per tensor plus a re-pack of all 119 tensors per step.
 * `segs` is a DEVICE array; tiles are 64 x 64 elements (matrices) or 1024 elements (vectors); tile0 = prefix sum of tile counts. */
typedef struct StairAdamSeg {
    float* p; const float* g; float* m; float* v;            /* [rows, cols] fp32, contiguous */
    float* p2; const float* g2; float* m2; float* v2;        /* optional second parameter whose SUM with p the kernels read (LSTM b_ih + b_hh); NULL = none */
    void* packed; int64_t packed_ld; int64_t packed_plane;   /* kind 1: bf16 row r at packed + r*ld (+ plane*packed_plane); kind 0: fp32 vector */
    void* packed_t; int64_t packed_t_ld; int64_t packed_t_plane;   /* transposed bf16 copy: element (r, c) at packed_t + c*ld_t + r; NULL = none */
    void* packed_perm;                                       /* bf16 copy with rows gate-interleaved (r = g*hh + c64*64 + j -> c64*256 + g*64 + j); NULL = none */
    int32_t rows, cols, kind, nplanes, perm_hh, tile0;
    float bc1, bc2;                                          /* 1 - beta1^step, 1 - beta2^step of this parameter */
} StairAdamSeg;
int stair_adam_multi(const StairAdamSeg* segs /*DEVICE*/, int n_segs, int total_tiles, float lr, double beta1, double beta2, float eps, void* stream);
/* number of kernels stair_nmn_forward launched in its last call on this thread (bench.py's gpu_launches). */
int64_t stair_last_launch_count(void);

/* ---- raw-feature ingest, the step right before the path (video_nmn/dataset.py:134-172) ---------
 * RX / TGIF-QA: out[b,t,0:Da] = mean over the F frames of appearance[b,t,:,:] (dataset.py:150-152), out[b,t,Da:] = motion[b,t,:]
 * (dataset.py:161-172).  appearance [B,T,F,Da], motion [B,T,Dm] (NULL with Dm = 0), out [B,T,Da+Dm]; Da, Dm multiples of 8. */
int stair_ingest_pool_concat(const void* appearance, const void* motion, int in_dtype, void* out, int out_dtype, int B, int T, int F,
                             int Da, int Dm, void* stream);
/* I3D npy features: out[b,t,:] = feats[b, t*step, :] for t < T (dataset.py:138-141: every 2nd row, then [:max_video_length]) */
int stair_ingest_subsample(const void* feats, int in_dtype, void* out, int out_dtype, int B, int n_frames, int T, int D, int step, void* stream);

/* ---- host-side collate staging (NO GPU work; video_nmn/dataset.py:463-476 collate_fn / to_device) -------------
 * dst rows [dst_row[i], dst_row[i] + rows[i]) <- src[i] (rows[i] x cols elements, contiguous HOST memory), converting src_dtype -> dst_dtype
 * (fp32 -> bf16 is round-to-nearest-even, bit-identical to torch's conversion for every non-NaN value; NaN stays NaN); dst is HOST (pinned) row-major with pitch `cols`.
 * One call stages a whole batch with `threads` host threads (<= 0: all, at most 32) instead of one torch copy per question. */
int stair_host_collate_rows(const void* const* src /*HOST*/, const long long* rows /*HOST*/, const long long* dst_row /*HOST*/, int n, long long cols,
                            int src_dtype, void* dst /*HOST*/, int dst_dtype, int threads);

/* ---- Filter-audit head, the step right after the path (evaluate.py:65-117) ----------------------
 * out_idx[i, 0..k) = indices of the k phrase representations (reps fp32 [P, H]) most cosine-similar to query row i (row
 * row_idx[i] of q, or row i when row_idx is NULL; pitch ldq elements of `dtype`), out_sim the similarities, descending
 * (nn.CosineSimilarity eps 1e-8, torch.argsort(descending=True)[:k]; ties -> lowest index). */
int stair_cosine_topk(int dtype, const void* q, long long ldq, const int32_t* row_idx, const float* reps, int P, int H, int k,
                      int32_t* out_idx, float* out_sim, int n, void* stream);

/* ---- single operators (memory-bound kernels), exported for unit parity tests ----------------------------------- */
/* TemporalModule.relate_ (video_nmn/modules.py:290-308): cumsum before/after/between masks.  mode: 0 while, 1 before,
 * 2 after, 3 between (att then holds two rows per instance).  att [n][K][T] fp32 -> out [n][T]. */
int stair_relate_scan(const float* att, int mode, float* out, int n, int T, void* stream);
/* attention = (cos(f_t, k_k) + 1) * 0.49 (modules.py:205-216, :170-177): f [n*T,H], k [n*K,H] -> att [n][K][T] */
int stair_cos_attention(int dtype, const void* f, const void* k, int K, int T, int H, float* att, int n, void* stream);
/* RelateModule (modules.py:417-435): softmax_T(att +/- beta) ; sign = +1 forward, -1 backward */
int stair_relate(const float* att, const float* beta, int sign, float* out, int n, int T, void* stream);
/* torch.argmax over the last dim (first maximal index) */
int stair_argmax(const float* x, int32_t* out, int rows, int cols, void* stream);
/* L2Normalize (module_net.py:211-216, F.normalize eps 1e-12) of n rows [H] -> fp32 */
int stair_l2normalize(int dtype, const void* x, float* out, int n, int H, void* stream);
/* LayerNorm over H (eps 1e-5, biased variance) */
int stair_layernorm(int dtype, const void* x, const float* gamma, const float* beta, void* out, long long rows, int H, void* stream);
/* Filter's frame aggregation (modules.py:374): out[i] = sum_t x[i*T+t]  ([n*T,H] -> [n,H], act dtype) */
int stair_sum_frames(int dtype, const void* x, void* out, int n, int T, int H, void* stream);
/* AttnVideoModule (modules.py:330-340): vid[out_base+i][t] = att[att_idx[i]][t] * vid[feat_idx[i]][t]  (VID arena slots, ATT rows) */
int stair_attn_video(int dtype, void* vid, const int32_t* feat_idx, const float* att, const int32_t* att_idx, int out_base, int n, int T, int H,
                     void* stream);
/* ExistsFrameModule (modules.py:162-178): att[out_base+i][t] = (cos(vid[feat_idx[i]][t], vec[kw_idx[i]]) + 1) * 0.49 */
int stair_exists_frame(int dtype, const void* vid, const int32_t* feat_idx, const void* vec, const int32_t* kw_idx, float* att, int out_base,
                       int n, int T, int H, void* stream);
/* HasItem tail (modules.py:128-129): att[out_base+i][t] = sigmoid(w . x[i*T+t] + b) */
int stair_hasitem_tail(int dtype, const void* x, const float* w, const float* b, float* att, int out_base, int n, int T, int H, void* stream);
/* TMA-staged streaming variants of the row kernels: 0 off, 1 (default) HasItem tail, 2 also the cosine maps (slower; kept for measurement) */
int stair_set_row_stream(int on);
int stair_set_cos_impl(int impl);      /* cosine maps (Localize / ExistsFrame): 0 = instance-major kernel when T % 8 == 0 (product); 1 = row-major kernel */
/* fp32 -> bf16 rows, and fp32 -> three bf16 planes (x = p0 + p1 + p2) used by the strict mode */
int stair_cast_bf16(const float* src, long long ld_src, void* dst, long long ld_dst, long long rows, int cols, void* stream);
int stair_split3(const float* src, long long ld_src, void* dst, long long ld_dst, long long plane_rows, long long rows, int cols, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STAIR_B200_H */
