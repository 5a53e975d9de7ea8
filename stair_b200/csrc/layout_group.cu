// Device-side grouping of a batch's heterogeneous program layouts by (level, module type, variant).
//
// The reference interprets one question at a time with a Python stack (video_nmn/module_net.py:94-133); tree edges and
// levels follow utils/program_parser.py:182-200 (get_childrens_and_parents) and :307-321 (stat_module_levels).  Here the
// host compiles every question's token list to nodes carrying a dense group id (ascending = level-major schedule order)
// and the device performs a STABLE counting sort of all nodes of the batch by group id, assigns every node its output
// slot (group outputs are contiguous, so GEMM epilogues write straight into the arenas) and resolves argument edges to
// arena indices.  Stability makes the result a pure function of the batch (bit-exact against the host/oracle grouping).
//
//   k1 group_hist     : one warp per chunk of CHUNK nodes -> cnt[chunk][group]
//   k2 group_scan     : per group exclusive scan over chunks; exclusive scan over groups -> group_off; checks it against
//                       the offsets the host derived from its own histogram (status[0] = 1 on mismatch)
//   k3 group_scatter  : stable in-chunk ranking with __match_any_sync; writes perm / out_slot / aux_slot
//   k4 group_resolve  : sorted position -> resolved argument indices, question id, word span
#include "nmn_kernels.cuh"

namespace stair {

constexpr int CHUNK = 1024;

__global__ void group_hist_kernel(const int* __restrict__ gid, int n_nodes, int n_groups, int* __restrict__ cnt) {
    extern __shared__ int hist[];
    const int lane = threadIdx.x;
    for (int g = lane; g < n_groups; g += 32) hist[g] = 0;
    __syncwarp();
    const int base = blockIdx.x * CHUNK;
    for (int i = lane; i < CHUNK && base + i < n_nodes; i += 32) {
        const int g = __ldg(gid + base + i);
        if (g >= 0 && g < n_groups) atomicAdd(&hist[g], 1);
    }
    __syncwarp();
    for (int g = lane; g < n_groups; g += 32) cnt[static_cast<long long>(blockIdx.x) * n_groups + g] = hist[g];
}

// single block; thread g owns group g (loops if n_groups > blockDim)
__global__ void group_scan_kernel(int* __restrict__ cnt, int n_chunks, int n_groups, const int* __restrict__ host_off,
                                  int n_nodes, int* __restrict__ group_off, int* __restrict__ status) {
    extern __shared__ int tot[];      // [n_groups]
    for (int g = threadIdx.x; g < n_groups; g += blockDim.x) {
        int run = 0;
        for (int c = 0; c < n_chunks; ++c) {
            const int v = cnt[static_cast<long long>(c) * n_groups + g];
            cnt[static_cast<long long>(c) * n_groups + g] = run;
            run += v;
        }
        tot[g] = run;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0, bad = 0;
        for (int g = 0; g < n_groups; ++g) {
            group_off[g] = run;
            if (host_off[g] != run) bad = 1;
            run += tot[g];
        }
        group_off[n_groups] = run;
        if (run != n_nodes) bad = 1;      // some node carried an out-of-range group id
        if (bad) atomicExch(status, 1);
    }
}

__global__ void group_scatter_kernel(const int* __restrict__ gid, int n_nodes, int n_groups, const int* __restrict__ cnt,
                                     const int* __restrict__ group_off, const int* __restrict__ out_base,
                                     const int* __restrict__ out_mult, const int* __restrict__ aux_base,
                                     int* __restrict__ perm, int* __restrict__ out_slot, int* __restrict__ aux_slot) {
    extern __shared__ int run[];      // running count per group inside this chunk
    const int lane = threadIdx.x;
    for (int g = lane; g < n_groups; g += 32) run[g] = 0;
    __syncwarp();
    const int base = blockIdx.x * CHUNK;
    for (int i0 = 0; i0 < CHUNK && base + i0 < n_nodes; i0 += 32) {
        const int node = base + i0 + lane;
        const bool valid = node < n_nodes;
        int g = valid ? __ldg(gid + node) : -1;
        if (g < 0 || g >= n_groups) g = -1;
        const unsigned peers = __match_any_sync(0xffffffffu, g);
        if (g >= 0) {
            const int rank_in_tile = __popc(peers & ((1u << lane) - 1u));
            const int r = run[g] + rank_in_tile + cnt[static_cast<long long>(blockIdx.x) * n_groups + g];   // rank inside the group
            const int pos = group_off[g] + r;
            perm[pos] = node;
            out_slot[node] = out_base[g] + r * out_mult[g];
            aux_slot[node] = aux_base[g] >= 0 ? aux_base[g] + r : -1;
        }
        __syncwarp();
        if (g >= 0 && lane == (31 - __clz(peers))) run[g] += __popc(peers);     // highest peer lane updates the running count
        __syncwarp();
    }
}

__global__ void group_resolve_kernel(const int* __restrict__ perm, const int* __restrict__ out_slot, const int* __restrict__ node_arg,
                                     const int* __restrict__ node_q, const int* __restrict__ node_span, int n_nodes,
                                     int* __restrict__ arg_slot, int* __restrict__ pos_q, int* __restrict__ pos_span) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_nodes) return;
    const int node = perm[p];
    const int q = __ldg(node_q + node);
    pos_q[p] = q;
    pos_span[p] = __ldg(node_span + node);
    pos_span[n_nodes + p] = __ldg(node_span + n_nodes + node);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int a = __ldg(node_arg + static_cast<long long>(k) * n_nodes + node);
        arg_slot[static_cast<long long>(k) * n_nodes + p] = a >= 0 ? out_slot[a] : (a == -2 ? q : -1);   // -2: 'video' -> VID slot q
    }
}

static void itab_layout(int n_nodes, int n_groups, StairItabLayout* L) {
    auto up = [](int64_t x) { return (x + 3) / 4 * 4; };          // keep every table 16-byte aligned
    int64_t o = 0;
    L->perm = o; o += up(n_nodes);
    L->out_slot = o; o += up(n_nodes);
    L->aux_slot = o; o += up(n_nodes);
    L->arg_slot = o; o += up(3LL * n_nodes);
    L->pos_q = o; o += up(n_nodes);
    L->pos_span = o; o += up(2LL * n_nodes);
    L->group_off = o; o += up(n_groups + 1);
    const int64_t chunks = (n_nodes + CHUNK - 1) / CHUNK;
    o += up(chunks * n_groups);                                    // cnt
    L->total = o;
}

int launch_group_layouts(const StairBatch& b, int32_t* itab, int32_t* status, cudaStream_t st) {
    if (b.n_nodes <= 0 || b.n_groups <= 0) return STAIR_OK;
    if (b.n_groups > 8192) return STAIR_ERR_CAPACITY;
    StairItabLayout L;
    itab_layout(b.n_nodes, b.n_groups, &L);
    const int chunks = (b.n_nodes + CHUNK - 1) / CHUNK;
    int* cnt = itab + L.group_off + (b.n_groups + 1 + 3) / 4 * 4;
    const size_t sh = b.n_groups * sizeof(int);
    const int* tab = b.group_tab;
    group_hist_kernel<<<chunks, 32, sh, st>>>(b.node_gid, b.n_nodes, b.n_groups, cnt);
    STAIR_CHECK_LAUNCH();
    group_scan_kernel<<<1, 256, sh, st>>>(cnt, chunks, b.n_groups, tab, b.n_nodes, itab + L.group_off, status);
    STAIR_CHECK_LAUNCH();
    group_scatter_kernel<<<chunks, 32, sh, st>>>(b.node_gid, b.n_nodes, b.n_groups, cnt, itab + L.group_off, tab + b.n_groups,
                                                tab + 2 * b.n_groups, tab + 3 * b.n_groups, itab + L.perm, itab + L.out_slot,
                                                itab + L.aux_slot);
    STAIR_CHECK_LAUNCH();
    group_resolve_kernel<<<(b.n_nodes + 255) / 256, 256, 0, st>>>(itab + L.perm, itab + L.out_slot, b.node_arg, b.node_q, b.node_span,
                                                                  b.n_nodes, itab + L.arg_slot, itab + L.pos_q, itab + L.pos_span);
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

}  // namespace stair

using namespace stair;

extern "C" int64_t stair_itab_ints(int32_t n_nodes, int32_t n_groups) {
    StairItabLayout L;
    itab_layout(n_nodes, n_groups, &L);
    return L.total;
}
extern "C" int stair_itab_layout(int32_t n_nodes, int32_t n_groups, StairItabLayout* out) {
    if (!out) return STAIR_ERR_ARG;
    itab_layout(n_nodes, n_groups, out);
    return STAIR_OK;
}
extern "C" int stair_group_layouts(const StairBatch* batch, int32_t* itab, int32_t* status, void* stream) {
    if (!batch || !itab || !status) return STAIR_ERR_ARG;
    return launch_group_layouts(*batch, itab, status, reinterpret_cast<cudaStream_t>(stream));
}
