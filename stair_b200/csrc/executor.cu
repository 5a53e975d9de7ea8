// Batched program interpreter: VideoNMN.forward (video_nmn/module_net.py:65-145) for B questions at once.
//
// The reference walks each question's prefix program with a Python stack and calls one nn.Module per token.  Here the
// whole batch is executed level-synchronously: after the device grouping step (layout_group.cu) every
// (level, module type, variant) group is a contiguous run of instances, and a group is executed with a handful of
// launches — tcgen05 GEMMs whose A rows are TMA-gathered straight from the frame-feature arena, plus the memory-bound
// kernels of nmn_kernels.cu.  Nothing here synchronises with the host or allocates; Choose is a device-side select.
//
// Scratch is carved from the caller's workspace by `Plan` (same arithmetic in stair_nmn_workspace_bytes).
#include "exec_core.cuh"
#include <cstdlib>

namespace stair {
namespace ex {
int g_lstm_impl = 0;                     // 0 = fused persistent recurrence when eligible, 1 = per-step GEMM + cell kernels
thread_local long long t_last_launches = 0;
int g_timeline = 0;
int g_fuse_sum = 0;                      // 1 = Filter's frame sum in the epilogue of its second Linear (inference).  Bit-identical, fewer launches (81 -> 78
                                         // at RX, 169 -> 155 at I3D) and NOT faster: these K = 512 GEMMs are epilogue-bound, and the transposing sum costs the
                                         // epilogue more than the separate 4-20 us pass (RX 1.354 -> 1.365-1.377 ms, I3D 5.57-5.72 -> 5.59-5.62 ms;
                                         // profiles/r2_fused_frame_sum_ab.txt)
static int text_sort_default() {             // STAIR_TEXT_SORT=0|1 overrides the default at library load (A/B runs)
    const char* e = getenv("STAIR_TEXT_SORT");
    return (e && (e[0] == '0' || e[0] == '1')) ? e[0] - '0' : 1;
}
int g_text_sort = text_sort_default();         // 1 = inference text recurrence over length-sorted questions (bit-identical outputs; run_encoders)
int g_dep_sched = 1;         // 1 = dependency-driven module scheduling when the batch carries group_deps
int g_lanes = 8;             // round 1 (wave scheduling), B = 4096 RX: 2 / 4 / 6 / 8 lanes = 1.64 / 1.65 / 1.54 / 1.55 ms per forward; round 2
                             // (dependency scheduling): 4 / 6 / 8 lanes = 1.365 / 1.36 / 1.34 ms vs 1.41-1.43 ms wave by wave (profiles/r2_dep_sched_ab.txt)
}

}  // namespace stair

using namespace stair;
using namespace stair::ex;

extern "C" int stair_set_lstm_impl(int impl) { g_lstm_impl = impl; return STAIR_OK; }
extern "C" int stair_debug_timeline(int on) { g_timeline = on ? 1 : 0; return STAIR_OK; }
// start / end of every group of the last dependency-scheduled module phase, in ms after the phase began (device must be idle: synchronises)
extern "C" int stair_debug_timeline_read(float* t0, float* t1, int* lane, int* op, int* count, int* variant, int cap) {
    Timeline& tl = timeline_state();
    if (!tl.ok) return 0;
    if (cudaDeviceSynchronize() != cudaSuccess) return STAIR_ERR_CUDA;
    const int n = tl.n < cap ? tl.n : cap;
    for (int g = 0; g < n; ++g) {
        cudaEventElapsedTime(&t0[g], tl.origin, tl.t0[g]);
        cudaEventElapsedTime(&t1[g], tl.origin, tl.t1[g]);
        lane[g] = tl.lane[g]; op[g] = tl.op[g]; count[g] = tl.count[g]; variant[g] = tl.variant[g];
    }
    return n;
}
// ms from the start of the last forward (with stair_debug_timeline(1)) to: [0] = 0, [1] video projection done, [2] text projection done,
// [3] recurrence done, [4] grouping joined, [5] module phase done, [6] decoder done (synchronises; a mark the call did not pass is stale)
extern "C" int stair_debug_phase_marks(float* ms, int cap) {
    PhaseMarks& pm = phase_marks();
    if (!pm.ok) return 0;
    if (cudaDeviceSynchronize() != cudaSuccess) return STAIR_ERR_CUDA;
    const int n = cap < 7 ? cap : 7;
    for (int i = 0; i < n; ++i) if (cudaEventElapsedTime(&ms[i], pm.ev[0], pm.ev[i]) != cudaSuccess) ms[i] = -1.0f;
    return n;
}
extern "C" int stair_set_fuse_sum(int on) { g_fuse_sum = on ? 1 : 0; return STAIR_OK; }
extern "C" int stair_set_text_sort(int on) { g_text_sort = on ? 1 : 0; return STAIR_OK; }
extern "C" int stair_set_dep_sched(int on) { g_dep_sched = on ? 1 : 0; return STAIR_OK; }
extern "C" int stair_set_lanes(int lanes) { g_lanes = lanes < 1 ? 1 : (lanes > LANES ? LANES : lanes); return STAIR_OK; }
extern "C" int stair_version(void) { return STAIR_ABI_VERSION; }
// explicit ownership of the library's few runtime objects (see include/stair_b200.h "Conventions")
extern "C" int stair_init(void) {
    if (!lane_streams() || !err_flag_ptr()) return STAIR_ERR_CUDA;
    return STAIR_OK;
}
extern "C" int stair_shutdown(void) {
    if (cudaDeviceSynchronize() != cudaSuccess) return STAIR_ERR_CUDA;
    lane_streams_destroy();
    err_flag_free();
    return STAIR_OK;
}
extern "C" int64_t stair_sizeof(int which) {
    switch (which) {
    case 0: return sizeof(StairModel);
    case 1: return sizeof(StairGroup);
    case 2: return sizeof(StairBatch);
    case 3: return sizeof(StairBuffers);
    case 4: return sizeof(StairItabLayout);
    case 5: return sizeof(StairTrain);
    case 6: return sizeof(StairAdamSeg);
    default: return -1;
    }
}

extern "C" int64_t stair_nmn_workspace_bytes(const StairModel* model, const StairBatch* batch) {
    if (!model || !batch) return -1;
    Plan p;
    make_plan(*model, *batch, &p);
    return p.total + 4096;
}

extern "C" int64_t stair_last_launch_count(void) { return t_last_launches; }

extern "C" int stair_nmn_forward(const StairModel* model, const StairBatch* batch, const StairBuffers* buf, int phases, void* stream) {
    if (!model || !batch || !buf) return STAIR_ERR_ARG;
    const StairModel& m = *model; const StairBatch& b = *batch;
    if (b.B <= 0) return STAIR_OK;
    if (m.H % 16 || m.H > 2048 || b.T <= 0 || m.V_ld % 8 || m.text_ld % 8) return STAIR_ERR_UNSUPPORTED;
    if (m.conv_k == 0 && b.T != m.T_max) return STAIR_ERR_UNSUPPORTED;     // Linear(T_max,T_max) relate needs T == T_max (modules.py:271-277)
    Ctx c{m, b, *buf, reinterpret_cast<cudaStream_t>(stream)};
    make_plan(m, b, &c.plan);
    if (buf->workspace_bytes < c.plan.total) return STAIR_ERR_CAPACITY;
    c.ws = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(buf->workspace) + 1023) & ~static_cast<uintptr_t>(1023));
    if (c.ws + c.plan.total > reinterpret_cast<char*>(buf->workspace) + buf->workspace_bytes) return STAIR_ERR_CAPACITY;
    c.T = b.T; c.H = m.H; c.h = m.H / 2; c.np = m.precision == STAIR_F32 ? 3 : 1; c.adt = m.precision; c.esz = m.precision == STAIR_F32 ? 4 : 2;
    c.inference = true;
    stair_itab_layout(b.n_nodes, b.n_groups, &c.il);
    if (buf->itab_ints < c.il.total) return STAIR_ERR_CAPACITY;
    c.perm = buf->itab + c.il.perm; c.out_slot = buf->itab + c.il.out_slot;
    c.arg0 = buf->itab + c.il.arg_slot; c.arg1 = c.arg0 + b.n_nodes; c.arg2 = c.arg1 + b.n_nodes;
    c.pos_q = buf->itab + c.il.pos_q; c.span_s = buf->itab + c.il.pos_span; c.span_e = c.span_s + b.n_nodes;
    const long long before = g_launch_count;
    // the layout grouping (4 small integer kernels, ~60 us of latency) does not depend on the encoders: it runs on a side lane
    // while the input projections execute, and is joined before the first module group
    const bool enc = (phases & (STAIR_FWD_ENCODE_VIDEO | STAIR_FWD_ENCODE_TEXT)) != 0;
    LaneStreams* ls = ((phases & STAIR_FWD_GROUP) && enc && g_lanes > 1) ? lane_streams() : nullptr;
    phase_mark(c.st, 0);
    if (phases & STAIR_FWD_GROUP) {
        if (ls) {
            if (cudaEventRecord(ls->fork, c.st) != cudaSuccess || cudaStreamWaitEvent(ls->side[0], ls->fork, 0) != cudaSuccess) return STAIR_ERR_CUDA;
            STAIR_TRY(launch_group_layouts(b, buf->itab, buf->status, ls->side[0]));
            if (cudaEventRecord(ls->join[0], ls->side[0]) != cudaSuccess) return STAIR_ERR_CUDA;
        } else {
            STAIR_TRY(launch_group_layouts(b, buf->itab, buf->status, c.st));
        }
    }
    if (enc) STAIR_TRY(run_encoders(c, phases));
    phase_mark(c.st, 3);
    if (ls && cudaStreamWaitEvent(c.st, ls->join[0], 0) != cudaSuccess) return STAIR_ERR_CUDA;
    phase_mark(c.st, 4);
    if (phases & STAIR_FWD_MODULES) STAIR_TRY(run_modules(c));
    phase_mark(c.st, 5);
    if (phases & STAIR_FWD_DECODE) STAIR_TRY(run_decoder(c));
    phase_mark(c.st, 6);
    t_last_launches = g_launch_count - before;
    return STAIR_OK;
}

// ---- one operator group outside the interpreter ----------------------------------------------------------------------------
// The reference exposes every operator as an nn.Module with its own forward(*params) (video_nmn/modules.py:7-465); the drop-in's
// per-class forward (stair_b200/modules.py) packs the operands of n instances into small arenas and runs them through the very
// same group code the interpreter uses (run_chunk), so a per-operator parity test exercises the product kernels.
static int op_setup(const StairModel& m, int T, const StairGroup& g, StairBatch* b, Plan* plan) {
    if (m.H % 16 || m.H > 2048 || T <= 0 || g.count <= 0) return STAIR_ERR_ARG;
    if (g.op <= STAIR_OP_WORD || g.op >= STAIR_OP_COUNT) return STAIR_ERR_ARG;
    if (m.conv_k == 0 && g.op == STAIR_OP_TEMPORAL && (g.variant >> 1) > 0 && T != m.T_max) return STAIR_ERR_UNSUPPORTED;
    *b = StairBatch();
    b->T = T; b->n_groups = 1; b->groups = &g; b->n_nodes = g.count;
    make_plan(m, *b, plan);
    return STAIR_OK;
}

extern "C" int64_t stair_op_workspace_bytes(const StairModel* model, int T, const StairGroup* group) {
    if (!model || !group) return -1;
    StairBatch b; Plan plan;
    if (op_setup(*model, T, *group, &b, &plan) != STAIR_OK) return -1;
    return plan.mod_bytes + 4096;
}

extern "C" int stair_op_forward(const StairModel* model, int T, const StairGroup* group, const int32_t* args, const StairBuffers* buf, void* stream) {
    if (!model || !group || !args || !buf) return STAIR_ERR_ARG;
    const StairModel& m = *model; const StairGroup& g = *group;
    StairBatch b; Plan plan;
    STAIR_TRY(op_setup(m, T, g, &b, &plan));
    Ctx c{m, b, *buf, reinterpret_cast<cudaStream_t>(stream)};
    c.plan = plan;
    if (buf->workspace_bytes < plan.mod_bytes + 1024) return STAIR_ERR_CAPACITY;
    c.ws = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(buf->workspace) + 1023) & ~static_cast<uintptr_t>(1023));
    c.T = T; c.H = m.H; c.h = m.H / 2; c.np = m.precision == STAIR_F32 ? 3 : 1; c.adt = m.precision; c.esz = m.precision == STAIR_F32 ? 4 : 2;
    c.perm = c.out_slot = c.pos_q = c.span_s = c.span_e = nullptr;
    c.inference = true;
    c.arg0 = args; c.arg1 = args + g.count; c.arg2 = args + 2 * g.count;
    const long long before = g_launch_count;
    StairGroup g0 = g;
    g0.node_off = 0;
    const int rc = run_group(c, g0, 0);
    t_last_launches = g_launch_count - before;
    return rc;
}
