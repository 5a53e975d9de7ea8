// TMA-staged streaming kernels for the row-wise dot-product operators (HBM-bound):
//   cosine attention maps   Localize modules.py:205-216, ExistsFrame modules.py:170-177   att = (cos(f_t, kw) + 1) * 0.49
//   HasItem tail            modules.py:128-129                                           att = sigmoid(w . x_t + b)
// A producer thread copies tiles of 32 frame rows (plus the keyword rows of their instances) into a shared-memory ring with
// cp.async.bulk (1-D TMA, completion on an mbarrier); 8 consumer warps reduce 4 rows each from shared memory.  The bytes in
// flight are set by the ring depth (128-160 KB per SM), not by registers x occupancy: the register-staged versions of these
// kernels stopped at 27-53 % of the HBM roofline, see profiles/r1_module_kernel_roofline.txt.
#include "nmn_kernels.cuh"
#include "tc_ptx.cuh"

namespace stair {
namespace {

constexpr int RS_TILE_ROWS = 32;
constexpr int RS_CONSUMERS = 8;                   // warps
constexpr int RS_THREADS = (RS_CONSUMERS + 1) * 32;
constexpr int RS_MAX_KW_ROWS = 8;                 // keyword rows per tile: (32 / T) instances x K <= 8

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct RowStreamParams {
    const void* f;            // frame rows: contiguous [rows, H] or the VID arena when feat_idx != null
    const int* feat_idx;      // instance -> VID slot
    const void* kw;           // keyword rows: contiguous [n*K, H] or the VEC arena when kw_idx != null (cos mode)
    const int* kw_idx;        // instance -> VEC row
    const float* w;           // HasItem: weight [H]
    const float* b;           // HasItem: bias [1]
    float* att;               // output maps
    long long out_base;       // first ATT row
    int n, K, T, H, stages;
    int* err_flag;
};

// MODE 0: cosine map, MODE 1: sigmoid(w . x + b)
template <typename AT, int MODE, int CH>
__global__ void __launch_bounds__(RS_THREADS, 1) row_stream_kernel(const RowStreamParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int T = p.T, H = p.H, K = p.K, hc = H / 8;
    const int row_bytes = H * static_cast<int>(sizeof(AT));
    const int f_bytes = RS_TILE_ROWS * row_bytes;
    const int stage_bytes = f_bytes + (MODE == 0 ? RS_MAX_KW_ROWS * row_bytes : 0);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(p.stages) * stage_bytes);
    uint64_t* empty = full + p.stages;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long rows = static_cast<long long>(p.n) * T;
    const long long tiles = (rows + RS_TILE_ROWS - 1) / RS_TILE_ROWS;
    // a tile is 32 / T whole instances (T <= 32) or 32 rows of one instance (T a multiple of 32)
    const int ipt = T <= RS_TILE_ROWS ? RS_TILE_ROWS / T : 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], RS_CONSUMERS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == RS_CONSUMERS) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
                mbar_wait(&empty[stage], phase ^ 1, p.err_flag, 301);
                uint8_t* dst = smem + static_cast<size_t>(stage) * stage_bytes;
                const long long row0 = tile * RS_TILE_ROWS;
                const int nrows = static_cast<int>(rows - row0 < RS_TILE_ROWS ? rows - row0 : RS_TILE_ROWS);
                const long long inst0 = row0 / T;
                const int ninst = T <= RS_TILE_ROWS ? (nrows + T - 1) / T : 1;
                uint32_t bytes = static_cast<uint32_t>(nrows) * row_bytes;
                if (MODE == 0) bytes += static_cast<uint32_t>(ninst) * K * row_bytes;
                mbar_arrive_expect_tx(&full[stage], bytes);
                const AT* f = reinterpret_cast<const AT*>(p.f);
                if (!p.feat_idx) bulk_load(dst, f + row0 * H, static_cast<uint32_t>(nrows) * row_bytes, &full[stage]);
                else if (T <= RS_TILE_ROWS) {
                    for (int i = 0; i < ninst; ++i)
                        bulk_load(dst + static_cast<size_t>(i) * T * row_bytes, f + static_cast<long long>(__ldg(p.feat_idx + inst0 + i)) * T * H,
                                  static_cast<uint32_t>(T) * row_bytes, &full[stage]);
                } else {
                    const int t0 = static_cast<int>(row0 - inst0 * T);
                    bulk_load(dst, f + (static_cast<long long>(__ldg(p.feat_idx + inst0)) * T + t0) * H, static_cast<uint32_t>(nrows) * row_bytes, &full[stage]);
                }
                if (MODE == 0) {
                    const AT* kw = reinterpret_cast<const AT*>(p.kw);
                    uint8_t* kd = dst + f_bytes;
                    if (!p.kw_idx) bulk_load(kd, kw + inst0 * K * H, static_cast<uint32_t>(ninst) * K * row_bytes, &full[stage]);
                    else
                        for (int i = 0; i < ninst; ++i)
                            bulk_load(kd + static_cast<size_t>(i) * row_bytes, kw + static_cast<long long>(__ldg(p.kw_idx + inst0 + i)) * H,
                                      static_cast<uint32_t>(row_bytes), &full[stage]);
                }
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
        return;
    }

    // ---- consumers: warp w reduces rows w, w+8, w+16, w+24 of every tile ------------------------------------------------------
    Vec8<float> wv[CH];
    float bias = 0.f;
    if (MODE == 1) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            const int c = lane + 32 * i;
            if (c < hc) wv[i].load(p.w + c * 8);
            else {
#pragma unroll
                for (int j = 0; j < 8; ++j) wv[i].v[j] = 0.f;
            }
        }
        bias = __ldg(p.b);
    }
    int stage = 0; uint32_t phase = 0;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        mbar_wait(&full[stage], phase, p.err_flag, 302);
        const uint8_t* src = smem + static_cast<size_t>(stage) * stage_bytes;
        const long long row0 = tile * RS_TILE_ROWS;
        const int nrows = static_cast<int>(rows - row0 < RS_TILE_ROWS ? rows - row0 : RS_TILE_ROWS);
        const long long inst0 = row0 / T;
        constexpr int RPW = RS_TILE_ROWS / RS_CONSUMERS;
        const int nk = MODE == 0 ? K : 1;
        for (int k = 0; k < nk; ++k) {
            float dot[RPW], kk[RPW], ff[RPW];
#pragma unroll
            for (int rr = 0; rr < RPW; ++rr) {
                const int r = warp + rr * RS_CONSUMERS;
                float d = 0.f, q = 0.f, s2 = 0.f;
                if (r < nrows) {
                    const int li = T <= RS_TILE_ROWS ? r / T : 0;            // local instance of this row
                    const AT* fr = reinterpret_cast<const AT*>(src) + static_cast<size_t>(r) * H;
                    const AT* kr = reinterpret_cast<const AT*>(src + f_bytes) + static_cast<size_t>(li * K + k) * H;
#pragma unroll
                    for (int i = 0; i < CH; ++i) {
                        const int c = lane + 32 * i;
                        if (c < hc) {
                            Vec8<AT> x; x.load(fr + c * 8);
                            if (MODE == 0) {
                                Vec8<AT> y; y.load(kr + c * 8);
#pragma unroll
                                for (int j = 0; j < 8; ++j) { d += x.v[j] * y.v[j]; q += y.v[j] * y.v[j]; s2 += x.v[j] * x.v[j]; }
                            } else {
#pragma unroll
                                for (int j = 0; j < 8; ++j) d += x.v[j] * wv[i].v[j];
                            }
                        }
                    }
                }
                dot[rr] = d; kk[rr] = q; ff[rr] = s2;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int rr = 0; rr < RPW; ++rr) {
                    dot[rr] += __shfl_xor_sync(0xffffffffu, dot[rr], o);
                    if (MODE == 0) {
                        kk[rr] += __shfl_xor_sync(0xffffffffu, kk[rr], o);
                        ff[rr] += __shfl_xor_sync(0xffffffffu, ff[rr], o);
                    }
                }
            }
#pragma unroll
            for (int rr = 0; rr < RPW; ++rr) {
                const int r = warp + rr * RS_CONSUMERS;
                if (lane == rr && r < nrows) {
                    const long long row = row0 + r;
                    if (MODE == 0) {
                        const long long inst = row / T;
                        const int t = static_cast<int>(row - inst * T);
                        const float nf = fmaxf(sqrtf(ff[rr]), 1e-8f), nkk = fmaxf(sqrtf(kk[rr]), 1e-8f);
                        p.att[(p.out_base + inst * K + k) * T + t] = (dot[rr] / (nf * nkk) + 1.0f) * 0.49f;
                    } else {
                        p.att[p.out_base * T + row] = sigmoidf_(dot[rr] + bias);
                    }
                }
            }
        }
        (void)inst0;
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
}

template <typename AT, int MODE>
int launch_row_stream_t(const RowStreamParams& p0, cudaStream_t st) {
    RowStreamParams p = p0;
    const int row_bytes = p.H * static_cast<int>(sizeof(AT));
    const int stage_bytes = RS_TILE_ROWS * row_bytes + (MODE == 0 ? RS_MAX_KW_ROWS * row_bytes : 0);
    int stages = (200 * 1024) / stage_bytes;
    if (stages > 4) stages = 4;
    if (stages < 2) return STAIR_ERR_UNSUPPORTED;
    p.stages = stages;
    const int smem = stages * stage_bytes + 2 * stages * 8 + 128;
    const long long rows = static_cast<long long>(p.n) * p.T;
    const long long tiles = (rows + RS_TILE_ROWS - 1) / RS_TILE_ROWS;
    const int grid = static_cast<int>(tiles < 148 ? tiles : 148);
    const int hc = p.H / 8;
#define RS_LAUNCH(CHV)                                                                                                        \
    do {                                                                                                                      \
        static int configured = 0;                                                                                            \
        if (configured < smem) {                                                                                              \
            if (cudaFuncSetAttribute(row_stream_kernel<AT, MODE, CHV>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) \
                return STAIR_ERR_CUDA;                                                                                        \
            configured = smem;                                                                                                \
        }                                                                                                                     \
        row_stream_kernel<AT, MODE, CHV><<<grid, RS_THREADS, smem, st>>>(p);                                                  \
    } while (0)
    if (hc <= 32) RS_LAUNCH(1);
    else if (hc <= 64) RS_LAUNCH(2);
    else if (hc <= 128) RS_LAUNCH(4);
    else RS_LAUNCH(8);
#undef RS_LAUNCH
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

}  // namespace

// shapes the streamer covers; everything else stays on the register-staged kernels of nmn_kernels.cu
bool row_stream_ok(int dt, int K, int T, int H) {
    const int esz = dt == STAIR_BF16 ? 2 : 4;
    if (H % 8 || H > 2048 || (H * esz) % 16) return false;
    if (!(T % 8 == 0 && ((RS_TILE_ROWS % T) == 0 || (T % RS_TILE_ROWS) == 0))) return false;
    const int ipt = T <= RS_TILE_ROWS ? RS_TILE_ROWS / T : 1;
    if (ipt * K > RS_MAX_KW_ROWS) return false;
    return (RS_TILE_ROWS + RS_MAX_KW_ROWS) * H * esz * 2 <= 200 * 1024;
}

int launch_cos_stream(int dt, const void* f, const int* feat_idx, const void* kw, const int* kw_idx, int K, int T, int H, float* att,
                      long long out_base, int n, cudaStream_t st) {
    if (n <= 0) return STAIR_OK;
    RowStreamParams p{};
    p.f = f; p.feat_idx = feat_idx; p.kw = kw; p.kw_idx = kw_idx; p.att = att; p.out_base = out_base; p.n = n; p.K = K; p.T = T; p.H = H;
    p.err_flag = err_flag_ptr();
    return dt == STAIR_BF16 ? launch_row_stream_t<bf16, 0>(p, st) : launch_row_stream_t<float, 0>(p, st);
}

int launch_rowdot_stream(int dt, const void* x, const float* w, const float* b, float* att, long long out_base, int n, int T, int H, cudaStream_t st) {
    if (n <= 0) return STAIR_OK;
    RowStreamParams p{};
    p.f = x; p.w = w; p.b = b; p.att = att; p.out_base = out_base; p.n = n; p.K = 1; p.T = T; p.H = H;
    p.err_flag = err_flag_ptr();
    return dt == STAIR_BF16 ? launch_row_stream_t<bf16, 1>(p, st) : launch_row_stream_t<float, 1>(p, st);
}

}  // namespace stair
