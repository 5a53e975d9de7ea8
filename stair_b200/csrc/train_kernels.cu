// Backward and loss kernels of the NMN training step.  Formulas are the derivatives of the forward kernels in
// nmn_kernels.cu / lstm.cu (reference forward: video_nmn/modules.py, module_net.py) and of the intermediate-supervision
// losses of train_module.py:83-194.  fp32 gradients, accumulate-into semantics (see train_kernels.cuh).
#include "train_kernels.cuh"

namespace stair {

#define DISPATCH_DT(dt, AT, ...)                                   \
    do {                                                           \
        if ((dt) == STAIR_BF16) { typedef bf16 AT; __VA_ARGS__; }  \
        else { typedef float AT; __VA_ARGS__; }                    \
    } while (0)

static inline int nblocks(long long work, int per_block, int cap = 148 * 16) {
    long long b = (work + per_block - 1) / per_block;
    if (b < 1) b = 1;
    return static_cast<int>(b > cap ? cap : b);
}

__device__ __forceinline__ void split3f(float x, bf16& p0, bf16& p1, bf16& p2) {
    p0 = __float2bfloat16_rn(x);
    float r = x - __bfloat162float(p0);
    p1 = __float2bfloat16_rn(r);
    r -= __bfloat162float(p1);
    p2 = __float2bfloat16_rn(r);
}
__device__ __forceinline__ uint32_t pack_bf16_(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void store_planes1(float v, bf16* dst, long long plane_stride, int nplanes) {
    if (nplanes == 1) { *dst = __float2bfloat16_rn(v); return; }
    bf16 a, b, c;
    split3f(v, a, b, c);
    dst[0] = a; dst[plane_stride] = b; dst[2 * plane_stride] = c;
}

// ------------------------------------------------------------------------------------------------------------------
// Linear backward staging
// ------------------------------------------------------------------------------------------------------------------
// One thread = 8 consecutive columns of one row (two float4 loads of dY, one 16-byte load of Y, 16-byte bf16 stores); the
// column sums of the bias gradient are kept in registers over the block's rows, reduced through shared memory and added with
// one atomic per column and block.
template <typename YT>
__global__ void dz_prep_kernel(const float* __restrict__ dY, long long ld_dy, const YT* __restrict__ Y, long long ld_y,
                               const float* __restrict__ rs, bf16* __restrict__ dZ, bf16* __restrict__ dZs, long long n_ld,
                               long long plane_rows, int nplanes, float* __restrict__ db, int M, int N, int rows_per_block, int vec_ok, float yscale) {
    __shared__ float red[256 * 8];
    const int n8 = static_cast<int>(n_ld / 8);
    const int cols_b = n8 < 256 ? n8 : 256;                      // threads along the columns
    const int rstep = 256 / cols_b;                              // rows processed in parallel by the block
    const int rl = threadIdx.x / cols_b, cl = threadIdx.x % cols_b;
    const int m_begin = blockIdx.x * rows_per_block, m_end = min(M, m_begin + rows_per_block);
    const long long ps = plane_rows * n_ld;
    for (int c0 = 0; c0 < n8; c0 += cols_b) {
        const int ch = c0 + cl, col = ch * 8;
        float colsum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (rl < rstep && ch < n8) {
            for (int m = m_begin + rl; m < m_end; m += rstep) {
                float dz[8];
                if (vec_ok && col + 8 <= N) {
                    const float4 a = *reinterpret_cast<const float4*>(dY + static_cast<long long>(m) * ld_dy + col);
                    const float4 b = *reinterpret_cast<const float4*>(dY + static_cast<long long>(m) * ld_dy + col + 4);
                    dz[0] = a.x; dz[1] = a.y; dz[2] = a.z; dz[3] = a.w; dz[4] = b.x; dz[5] = b.y; dz[6] = b.z; dz[7] = b.w;
                    if (Y) {
                        Vec8<YT> y; y.load(Y + static_cast<long long>(m) * ld_y + col);
#pragma unroll
                        for (int j = 0; j < 8; ++j) dz[j] = y.v[j] > 0.f ? dz[j] * yscale : 0.f;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float v = 0.f;
                        if (col + j < N) {
                            v = dY[static_cast<long long>(m) * ld_dy + col + j];
                            if (Y) v = ld1<YT>(Y + static_cast<long long>(m) * ld_y + col + j) > 0.f ? v * yscale : 0.f;
                        }
                        dz[j] = v;
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) colsum[j] += dz[j];
                bf16* d = dZ + static_cast<long long>(m) * n_ld + col;
                if (nplanes == 1) {
                    *reinterpret_cast<uint4*>(d) = make_uint4(pack_bf16_(dz[0], dz[1]), pack_bf16_(dz[2], dz[3]), pack_bf16_(dz[4], dz[5]), pack_bf16_(dz[6], dz[7]));
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) store_planes1(dz[j], d + j, ps, nplanes);
                }
                if (dZs) {
                    const float r = __ldg(rs + m);
                    bf16* e = dZs + static_cast<long long>(m) * n_ld + col;
                    if (nplanes == 1) {
                        *reinterpret_cast<uint4*>(e) = make_uint4(pack_bf16_(dz[0] * r, dz[1] * r), pack_bf16_(dz[2] * r, dz[3] * r),
                                                                  pack_bf16_(dz[4] * r, dz[5] * r), pack_bf16_(dz[6] * r, dz[7] * r));
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) store_planes1(dz[j] * r, e + j, ps, nplanes);
                    }
                }
            }
        }
        if (db) {                                                // block-level column sums -> one atomic per column and block
#pragma unroll
            for (int j = 0; j < 8; ++j) red[threadIdx.x * 8 + j] = colsum[j];
            __syncthreads();
            // red[r][cl][j] is contiguous in (cl, j) = the column: consecutive threads take consecutive columns, so the shared-memory
            // reads are conflict-free and a warp's 32 atomics hit 4 sectors of db instead of 32 (same-address atomics serialise
            // in the L2 slice, so both their number — few, fat blocks — and their sector spread matter)
            const int ncols = cols_b * 8;
            for (int cc = threadIdx.x; cc < ncols; cc += 256) {
                float t = 0.f;
                for (int r = 0; r < rstep; ++r) t += red[r * ncols + cc];
                const int col_g = c0 * 8 + cc;
                if (col_g < N && t != 0.f) atomicAdd(db + col_g, t);
            }
            __syncthreads();
        }
    }
}

int launch_dz_prep(int ydt, const float* dY, long long ld_dy, const void* Y, long long ld_y, const float* rs, bf16* dZ, bf16* dZs,
                   long long n_ld, long long plane_rows, int nplanes, float* db, int M, int N, float yscale, cudaStream_t st) {
    if (M <= 0 || N <= 0) return STAIR_OK;
    if (n_ld % 8) return STAIR_ERR_ARG;
    int rpb = (M + 591) / 592;                                   // <= 4 blocks per SM: every block adds N same-address atomics
    if (rpb < 8) rpb = 8;
    const int grid = (M + rpb - 1) / rpb;
    const int yesz = ydt == STAIR_BF16 ? 2 : 4;
    const int vec_ok = ((reinterpret_cast<uintptr_t>(dY) & 15) == 0 && (ld_dy % 4) == 0 &&
                        (!Y || ((reinterpret_cast<uintptr_t>(Y) & 15) == 0 && (ld_y * yesz) % 16 == 0))) ? 1 : 0;
    DISPATCH_DT(ydt, YT, (dz_prep_kernel<YT><<<grid, 256, 0, st>>>(dY, ld_dy, reinterpret_cast<const YT*>(Y), ld_y, rs, dZ, rs ? dZs : nullptr,
                                                                   n_ld, plane_rows, nplanes, db, M, N, rpb, vec_ok, yscale)));
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

// 64 x 64 bf16 tiles through shared memory: 16-byte loads along the source rows, 16-byte stores along the destination rows.
__global__ void transpose_planes_kernel(const bf16* __restrict__ src, long long ld_src, long long src_plane, bf16* __restrict__ dst,
                                        long long ld_dst, long long dst_plane, int R, int C, int vec_ok) {
    __shared__ unsigned short tile[64][66];
    const unsigned short* s = reinterpret_cast<const unsigned short*>(src) + blockIdx.z * src_plane;
    unsigned short* d = reinterpret_cast<unsigned short*>(dst) + blockIdx.z * dst_plane;
    const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
    const int tr = threadIdx.x >> 3, tc = (threadIdx.x & 7) * 8;           // 32 rows x 8 column chunks per pass
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        const int r = r0 + tr + 32 * pass, c = c0 + tc;
        unsigned short v[8];
        if (vec_ok && r < R && c + 8 <= C) {
            const uint4 q = *reinterpret_cast<const uint4*>(s + static_cast<long long>(r) * ld_src + c);
            const unsigned short* h = reinterpret_cast<const unsigned short*>(&q);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = h[j];
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = (r < R && c + j < C) ? s[static_cast<long long>(r) * ld_src + c + j] : static_cast<unsigned short>(0);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) tile[tr + 32 * pass][tc + j] = v[j];
    }
    __syncthreads();
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        const int c = c0 + tr + 32 * pass;                       // destination row = source column
        const int r = r0 + tc;                                   // 8 consecutive destination columns = source rows
        if (c < C && r < ld_dst) {
            uint4 q;
            unsigned short* h = reinterpret_cast<unsigned short*>(&q);
#pragma unroll
            for (int j = 0; j < 8; ++j) h[j] = tile[tc + j][tr + 32 * pass];
            if (r + 8 <= ld_dst) *reinterpret_cast<uint4*>(d + static_cast<long long>(c) * ld_dst + r) = q;
            else
                for (int j = 0; j < 8 && r + j < ld_dst; ++j) d[static_cast<long long>(c) * ld_dst + r + j] = h[j];
        }
    }
}

int launch_transpose_planes(const bf16* src, long long ld_src, long long src_plane_rows, bf16* dst, long long ld_dst, long long dst_plane_rows,
                            int nplanes, int R, int C, cudaStream_t st) {
    if (R <= 0 || C <= 0) return STAIR_OK;
    if (ld_dst % 8) return STAIR_ERR_ARG;
    dim3 grid((C + 63) / 64, static_cast<unsigned>((ld_dst + 63) / 64), nplanes);
    const int vec_ok = ((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (ld_src % 8) == 0 && ((src_plane_rows * ld_src) % 8) == 0) ? 1 : 0;
    transpose_planes_kernel<<<grid, 256, 0, st>>>(src, ld_src, src_plane_rows * ld_src, dst, ld_dst, dst_plane_rows * ld_dst, R, C, vec_ok);
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

template <typename XT>
__global__ void rowscale_bwd_kernel(float* __restrict__ G, const XT* __restrict__ X, const int* __restrict__ slots, int rps, int unit,
                                    const float* __restrict__ rs, float* __restrict__ dr, int M, int K) {
    const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
    for (int m = blockIdx.x * warps + (threadIdx.x >> 5); m < M; m += gridDim.x * warps) {
        const long long xr = slots ? static_cast<long long>(__ldg(slots + m / rps)) * unit + m % rps : m;
        float* g = G + static_cast<long long>(m) * K;
        const XT* x = X + xr * K;
        float dot = 0.f;
        for (int k = lane; k < K; k += 32) dot += g[k] * ld1<XT>(x + k);
        dot = warp_sum(dot);
        const float s = __ldg(rs + m);
        for (int k = lane; k < K; k += 32) g[k] *= s;
        if (lane == 0) dr[m] += dot;
    }
}

int launch_rowscale_bwd(int xdt, float* G, const void* X, const int* slots, int rps, int unit, const float* rs, float* dr, int M, int K, cudaStream_t st) {
    if (M <= 0) return STAIR_OK;
    DISPATCH_DT(xdt, XT, (rowscale_bwd_kernel<XT><<<nblocks(M, 8), 256, 0, st>>>(G, reinterpret_cast<const XT*>(X), slots, rps < 1 ? 1 : rps, unit, rs, dr, M, K)));
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

// 4 columns per thread: one 16-byte load and one vector reduction (red.global.add.v4.f32) instead of four scalar atomics
__global__ void scatter_add_rows_v4_kernel(const float* __restrict__ src, const int* __restrict__ idx, int rps, int unit, float* __restrict__ dst,
                                           long long rows, int H4) {
    const long long total = rows * H4;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / H4;
        const int c = static_cast<int>(i % H4);
        const long long dr = static_cast<long long>(__ldg(idx + r / rps)) * unit + r % rps;
        const float4 v = reinterpret_cast<const float4*>(src)[i];
        if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) atomicAdd(reinterpret_cast<float4*>(dst + dr * H4 * 4) + c, v);
    }
}

__global__ void scatter_add_rows_kernel(const float* __restrict__ src, const int* __restrict__ idx, int rps, int unit, float* __restrict__ dst,
                                        long long rows, int H) {
    const long long total = rows * H;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / H;
        const int c = static_cast<int>(i % H);
        const long long dr = static_cast<long long>(__ldg(idx + r / rps)) * unit + r % rps;
        const float v = src[i];
        if (v != 0.f) atomicAdd(dst + dr * H + c, v);
    }
}

int launch_scatter_add_rows(const float* src, const int* idx, int rps, int unit, float* dst, long long rows, int H, cudaStream_t st) {
    if (rows <= 0) return STAIR_OK;
    if ((H & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
        scatter_add_rows_v4_kernel<<<nblocks(rows * (H / 4), 256), 256, 0, st>>>(src, idx, rps < 1 ? 1 : rps, unit, dst, rows, H / 4);
        STAIR_CHECK_LAUNCH();
        return STAIR_OK;
    }
    scatter_add_rows_kernel<<<nblocks(rows * H, 256), 256, 0, st>>>(src, idx, rps < 1 ? 1 : rps, unit, dst, rows, H);
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// LayerNorm backward (Temporal.layer_norm).  Column partial sums of dgamma / dbeta stay in registers across rows.
// ------------------------------------------------------------------------------------------------------------------
constexpr int LN_MAXC = 4;      // H <= 1024 in training

template <typename XT>
__global__ void layernorm_bwd_kernel(const float* __restrict__ dOut, const XT* __restrict__ X, const float* __restrict__ gamma,
                                     float* __restrict__ dX, float* __restrict__ dgamma, float* __restrict__ dbeta, long long rows, int H) {
    const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
    const int hc = H / 8;
    float pg[LN_MAXC][8], pb[LN_MAXC][8];
#pragma unroll
    for (int i = 0; i < LN_MAXC; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) { pg[i][j] = 0.f; pb[i][j] = 0.f; }
    for (long long row = blockIdx.x * static_cast<long long>(warps) + (threadIdx.x >> 5); row < rows; row += static_cast<long long>(gridDim.x) * warps) {
        float x[LN_MAXC][8], dy[LN_MAXC][8];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < LN_MAXC; ++i) {
            const int c = lane + 32 * i;
            if (c < hc) {
                Vec8<XT> v; v.load(X + row * H + c * 8);
                Vec8<float> d; d.load(dOut + row * H + c * 8);
#pragma unroll
                for (int j = 0; j < 8; ++j) { x[i][j] = v.v[j]; dy[i][j] = d.v[j]; s += v.v[j]; }
            }
        }
        const float mean = warp_sum(s) / H;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < LN_MAXC; ++i)
            if (lane + 32 * i < hc)
#pragma unroll
                for (int j = 0; j < 8; ++j) { const float d = x[i][j] - mean; q += d * d; }
        const float rstd = rsqrtf(warp_sum(q) / H + 1e-5f);
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < LN_MAXC; ++i) {
            const int c = lane + 32 * i;
            if (c < hc)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float xh = (x[i][j] - mean) * rstd;
                    const float dxh = dy[i][j] * __ldg(gamma + c * 8 + j);
                    pg[i][j] += dy[i][j] * xh; pb[i][j] += dy[i][j];
                    x[i][j] = xh; dy[i][j] = dxh;
                    s1 += dxh; s2 += dxh * xh;
                }
        }
        s1 = warp_sum(s1) / H; s2 = warp_sum(s2) / H;
#pragma unroll
        for (int i = 0; i < LN_MAXC; ++i) {
            const int c = lane + 32 * i;
            if (c < hc) {
                Vec8<float> o;
#pragma unroll
                for (int j = 0; j < 8; ++j) o.v[j] = rstd * (dy[i][j] - s1 - x[i][j] * s2);
                o.store(dX + row * H + c * 8);
            }
        }
    }
    // block-level reduction (8 warps -> 1) through shared memory, then one coalesced atomic per column and block: same-address
    // atomics serialise in L2, so their number per column (= blocks) and their sector spread both matter
    __shared__ float red[8][LN_MAXC * 256];
    const int wi = threadIdx.x >> 5;
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
        for (int i = 0; i < LN_MAXC; ++i) {
            const int c = lane + 32 * i;
            if (c < hc)
#pragma unroll
                for (int j = 0; j < 8; ++j) red[wi][c * 8 + j] = pass == 0 ? pg[i][j] : pb[i][j];
        }
        __syncthreads();
        float* dst = pass == 0 ? dgamma : dbeta;
        for (int c = threadIdx.x; c < H; c += blockDim.x) {
            float t = 0.f;
            for (int w2 = 0; w2 < warps; ++w2) t += red[w2][c];
            if (t != 0.f) atomicAdd(dst + c, t);
        }
        __syncthreads();
    }
}

int launch_layernorm_bwd(int xdt, const float* dOut, const void* X, const float* gamma, float* dX, float* dgamma, float* dbeta, long long rows, int H, cudaStream_t st) {
    if (rows <= 0) return STAIR_OK;
    if (H % 8 || H > 256 * LN_MAXC) return STAIR_ERR_UNSUPPORTED;
    DISPATCH_DT(xdt, XT, (layernorm_bwd_kernel<XT><<<nblocks(rows, 32, 296), 256, 0, st>>>(dOut, reinterpret_cast<const XT*>(X), gamma, dX, dgamma, dbeta, rows, H)));
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// cosine attention backward: att = (cos(f_t, k_k) + 1) * 0.49
//   df_t = sum_k g (k/(|f||k|) - cos f/|f|^2)     dk_k = sum_t g (f/(|f||k|) - cos k/|k|^2),  g = 0.49 * datt[k,t]
// ------------------------------------------------------------------------------------------------------------------
template <typename AT>
__global__ void cos_att_bwd_f_kernel(const AT* __restrict__ f, const AT* __restrict__ kmat, int K, int T, int H, const float* __restrict__ datt,
                                     long long att_base, float* __restrict__ df, long long rows) {
    const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
    for (long long row = blockIdx.x * static_cast<long long>(warps) + (threadIdx.x >> 5); row < rows; row += static_cast<long long>(gridDim.x) * warps) {
        const long long i = row / T;
        const int t = static_cast<int>(row % T);
        const AT* fr = f + row * H;
        float ff = 0.f;
        for (int c = lane; c < H; c += 32) { const float x = ld1<AT>(fr + c); ff += x * x; }
        const float nf = fmaxf(sqrtf(warp_sum(ff)), 1e-8f);
        for (int c = lane; c < H; c += 32) df[row * H + c] = 0.f;
        for (int k = 0; k < K; ++k) {
            const AT* kr = kmat + (i * K + k) * H;
            float dot = 0.f, kk = 0.f;
            for (int c = lane; c < H; c += 32) { const float x = ld1<AT>(fr + c), y = ld1<AT>(kr + c); dot += x * y; kk += y * y; }
            dot = warp_sum(dot); kk = warp_sum(kk);
            const float nk = fmaxf(sqrtf(kk), 1e-8f);
            const float cosv = dot / (nf * nk);
            const float g = 0.49f * datt[(att_base + i * K + k) * T + t];
            for (int c = lane; c < H; c += 32) {
                const float x = ld1<AT>(fr + c), y = ld1<AT>(kr + c);
                df[row * H + c] += g * (y / (nf * nk) - cosv * x / (nf * nf));
            }
        }
    }
}

template <typename AT>
__global__ void cos_att_bwd_k_kernel(const AT* __restrict__ f, const AT* __restrict__ kmat, int K, int T, int H, const float* __restrict__ datt,
                                     long long att_base, float* __restrict__ dk, long long rows /* n*K */) {
    const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
    for (long long row = blockIdx.x * static_cast<long long>(warps) + (threadIdx.x >> 5); row < rows; row += static_cast<long long>(gridDim.x) * warps) {
        const long long i = row / K;
        const AT* kr = kmat + row * H;
        float kk = 0.f;
        for (int c = lane; c < H; c += 32) { const float y = ld1<AT>(kr + c); kk += y * y; }
        const float nk = fmaxf(sqrtf(warp_sum(kk)), 1e-8f);
        for (int c = lane; c < H; c += 32) dk[row * H + c] = 0.f;
        for (int t = 0; t < T; ++t) {
            const AT* fr = f + (i * T + t) * H;
            float dot = 0.f, ff = 0.f;
            for (int c = lane; c < H; c += 32) { const float x = ld1<AT>(fr + c), y = ld1<AT>(kr + c); dot += x * y; ff += x * x; }
            dot = warp_sum(dot); ff = warp_sum(ff);
            const float nf = fmaxf(sqrtf(ff), 1e-8f);
            const float cosv = dot / (nf * nk);
            const float g = 0.49f * datt[(att_base + row) * T + t];
            for (int c = lane; c < H; c += 32) {
                const float x = ld1<AT>(fr + c), y = ld1<AT>(kr + c);
                dk[row * H + c] += g * (x / (nf * nk) - cosv * y / (nk * nk));
            }
        }
    }
}

int launch_cos_att_bwd(int dt, const void* f, const void* kmat, int K, int T, int H, const float* datt, long long att_base, float* df, float* dk, int n, cudaStream_t st) {
    if (n <= 0) return STAIR_OK;
    const long long rows_f = static_cast<long long>(n) * T, rows_k = static_cast<long long>(n) * K;
    DISPATCH_DT(dt, AT, (cos_att_bwd_f_kernel<AT><<<nblocks(rows_f, 8), 256, 0, st>>>(reinterpret_cast<const AT*>(f), reinterpret_cast<const AT*>(kmat), K, T, H, datt, att_base, df, rows_f)));
    STAIR_CHECK_LAUNCH();
    DISPATCH_DT(dt, AT, (cos_att_bwd_k_kernel<AT><<<nblocks(rows_k, 8), 256, 0, st>>>(reinterpret_cast<const AT*>(f), reinterpret_cast<const AT*>(kmat), K, T, H, datt, att_base, dk, rows_k)));
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

template <typename AT>
__global__ void existsframe_bwd_kernel(const AT* __restrict__ vid, const int* __restrict__ feat_idx, const AT* __restrict__ vec,
                                       const int* __restrict__ kw_idx, const float* __restrict__ datt, int att_base, float* __restrict__ dvid,
                                       float* __restrict__ dvec, long long rows, int T, int H) {
    const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
    for (long long row = blockIdx.x * static_cast<long long>(warps) + (threadIdx.x >> 5); row < rows; row += static_cast<long long>(gridDim.x) * warps) {
        const int i = static_cast<int>(row / T), t = static_cast<int>(row % T);
        const long long fo = (static_cast<long long>(__ldg(feat_idx + i)) * T + t) * H, ko = static_cast<long long>(__ldg(kw_idx + i)) * H;
        float dot = 0.f, ff = 0.f, kk = 0.f;
        for (int c = lane; c < H; c += 32) { const float x = ld1<AT>(vid + fo + c), y = ld1<AT>(vec + ko + c); dot += x * y; ff += x * x; kk += y * y; }
        dot = warp_sum(dot); ff = warp_sum(ff); kk = warp_sum(kk);
        const float nf = fmaxf(sqrtf(ff), 1e-8f), nk = fmaxf(sqrtf(kk), 1e-8f);
        const float cosv = dot / (nf * nk);
        const float g = 0.49f * datt[static_cast<long long>(att_base + i) * T + t];
        if (g == 0.f) continue;
        for (int c = lane; c < H; c += 32) {
            const float x = ld1<AT>(vid + fo + c), y = ld1<AT>(vec + ko + c);
            atomicAdd(dvid + fo + c, g * (y / (nf * nk) - cosv * x / (nf * nf)));
            atomicAdd(dvec + ko + c, g * (x / (nf * nk) - cosv * y / (nk * nk)));
        }
    }
}

int launch_existsframe_bwd(int dt, const void* vid, const int* feat_idx, const void* vec, const int* kw_idx, const float* datt, int att_base,
                           float* dvid, float* dvec, int n, int T, int H, cudaStream_t st) {
    if (n <= 0) return STAIR_OK;
    const long long rows = static_cast<long long>(n) * T;
    DISPATCH_DT(dt, AT, (existsframe_bwd_kernel<AT><<<nblocks(rows, 8), 256, 0, st>>>(reinterpret_cast<const AT*>(vid), feat_idx, reinterpret_cast<const AT*>(vec),
                                                                                      kw_idx, datt, att_base, dvid, dvec, rows, T, H)));
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Temporal.relate[mode] backward (3 x Linear(T,T) or 3 x Conv1d 'same'); one block per instance, one thread per frame.
// ------------------------------------------------------------------------------------------------------------------
struct RelateBwdParams { const float* p[6]; float* d[6]; };

__global__ void temporal_relate_bwd_kernel(const float* __restrict__ att, const int* __restrict__ att_idx, int K, int mode, int conv_k,
                                           RelateBwdParams rp, const float* __restrict__ dr, float* __restrict__ datt, int n, int T, int wsz) {
    extern __shared__ float sm[];
    float* act[4] = {sm, sm + T, sm + 2 * T, sm + 3 * T};       // a0 (mean), a1, a2, r
    float* dz = sm + 4 * T;                                     // gradient wrt the current layer's pre-activation
    float* da = sm + 5 * T;
    // parameter gradients of the block's instances accumulate in shared memory (thread t owns row t of every dW / element t of db,
    // so no conflicts) and are added to the global gradient once per block: per-instance atomics put n x (T*T + T) x 3 same-address
    // atomics on 216 addresses, which serialise in L2
    float* accw = sm + 6 * T;                                   // [3][wsz]
    float* accb = accw + 3 * wsz;                               // [3][T]
    const int t = threadIdx.x;
    for (int j = t; j < 3 * wsz + 3 * T; j += blockDim.x) accw[j] = 0.f;
    __syncthreads();
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const long long r0 = __ldg(att_idx + i);
        if (t < T) {
            float s = 0.f;
            for (int k = 0; k < K; ++k) s += att[(r0 + k) * T + t];
            act[0][t] = s / static_cast<float>(K);
        }
        __syncthreads();
        if (mode == 0) {
            if (t < T) { const float g = dr[static_cast<long long>(i) * T + t] / K; for (int k = 0; k < K; ++k) atomicAdd(datt + (r0 + k) * T + t, g); }
            __syncthreads();
            continue;
        }
        for (int layer = 0; layer < 3; ++layer) {
            const float* w = rp.p[2 * layer]; const float* b = rp.p[2 * layer + 1];
            if (t < T) {
                float y;
                if (conv_k == 0) {
                    y = __ldg(b + t);
                    for (int u = 0; u < T; ++u) y += __ldg(w + t * T + u) * act[layer][u];
                } else {
                    const int k = layer < 2 ? conv_k : 2 * conv_k + 1, left = (k - 1) / 2;
                    y = __ldg(b);
                    for (int j = 0; j < k; ++j) { const int u = t + j - left; if (u >= 0 && u < T) y += __ldg(w + j) * act[layer][u]; }
                }
                act[layer + 1][t] = layer < 2 ? fmaxf(y, 0.f) : sigmoidf_(y);
            }
            __syncthreads();
        }
        if (t < T) { const float r = act[3][t]; dz[t] = dr[static_cast<long long>(i) * T + t] * r * (1.f - r); }
        __syncthreads();
        for (int layer = 2; layer >= 0; --layer) {
            const float* w = rp.p[2 * layer];
            float* aw = accw + layer * wsz; float* ab = accb + layer * T;
            if (conv_k == 0) {
                if (t < T) {
                    float g = 0.f;                               // d/d act[layer][t]
                    for (int o = 0; o < T; ++o) g += __ldg(w + o * T + t) * dz[o];
                    const float d = dz[t];
                    if (d != 0.f) { for (int u = 0; u < T; ++u) aw[t * T + u] += d * act[layer][u]; ab[t] += d; }
                    da[t] = g;
                }
            } else {
                const int k = layer < 2 ? conv_k : 2 * conv_k + 1, left = (k - 1) / 2;
                if (t < T) {
                    float g = 0.f;
                    for (int j = 0; j < k; ++j) { const int o = t - j + left; if (o >= 0 && o < T) g += __ldg(w + j) * dz[o]; }
                    da[t] = g;
                }
                if (t < k) {                                     // thread j = t accumulates dw[j] over positions
                    float s = 0.f;
                    for (int o = 0; o < T; ++o) { const int u = o + t - left; if (u >= 0 && u < T) s += dz[o] * act[layer][u]; }
                    aw[t] += s;
                }
                if (t == 0) { float s = 0.f; for (int o = 0; o < T; ++o) s += dz[o]; ab[0] += s; }
            }
            __syncthreads();
            if (t < T) dz[t] = layer > 0 ? (act[layer][t] > 0.f ? da[t] : 0.f) : da[t];     // ReLU of the previous layer
            __syncthreads();
        }
        if (t < T) { const float g = dz[t] / K; for (int k = 0; k < K; ++k) atomicAdd(datt + (r0 + k) * T + t, g); }
        __syncthreads();
    }
    if (mode == 0) return;
    for (int layer = 0; layer < 3; ++layer) {
        const int nw = conv_k == 0 ? T * T : (layer < 2 ? conv_k : 2 * conv_k + 1), nb = conv_k == 0 ? T : 1;
        for (int j = t; j < nw; j += blockDim.x) { const float v = accw[layer * wsz + j]; if (v != 0.f) atomicAdd(rp.d[2 * layer] + j, v); }
        for (int j = t; j < nb; j += blockDim.x) { const float v = accb[layer * T + j]; if (v != 0.f) atomicAdd(rp.d[2 * layer + 1] + j, v); }
    }
}

int launch_temporal_relate_bwd(const float* att, const int* att_idx, int K, int mode, int conv_k, const float* const* params, float* const* dparams,
                               const float* dr, float* datt, int n, int T, cudaStream_t st) {
    if (n <= 0) return STAIR_OK;
    if (T > 1024) return STAIR_ERR_UNSUPPORTED;
    RelateBwdParams rp;
    for (int j = 0; j < 6; ++j) { rp.p[j] = params ? params[j] : nullptr; rp.d[j] = dparams ? dparams[j] : nullptr; }
    int threads = ((T + 31) / 32) * 32;
    if (conv_k && threads < 2 * conv_k + 1) threads = ((2 * conv_k + 1 + 31) / 32) * 32;
    const int wsz = conv_k == 0 ? T * T : 2 * conv_k + 1;
    if (conv_k == 0 && T > 64) return STAIR_ERR_UNSUPPORTED;      // Linear(T,T) mode exists only for T_max <= 32 (modules.py:267)
    const size_t smem = (6 * T + 3 * wsz + 3 * T) * sizeof(float);
    const int grid = n < 148 * 4 ? n : 148 * 4;
    temporal_relate_bwd_kernel<<<grid, threads, smem, st>>>(att, att_idx, K, mode, conv_k, rp, dr, datt, n, T, wsz);
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

__global__ void bcast_T_kernel(const float* __restrict__ dagg, float* __restrict__ dx, long long total, int T, int H) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long row = i / H;
        dx[i] = dagg[(row / T) * H + i % H];
    }
}
int launch_bcast_T(const float* dagg, float* dx, int n, int T, int H, cudaStream_t st) {
    if (n <= 0) return STAIR_OK;
    const long long total = static_cast<long long>(n) * T * H;
    bcast_T_kernel<<<nblocks(total, 256), 256, 0, st>>>(dagg, dx, total, T, H);
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

// FilterFrame.attention backward: a = sigmoid(w_x.x + w_k.kw + b)
constexpr int FFA_MAXC = 32;       // columns per lane: H <= 1024
template <typename AT>
__global__ void ff_attn_bwd_kernel(const AT* __restrict__ x, const AT* __restrict__ vec, const int* __restrict__ kw_idx, const float* __restrict__ w,
                                   const float* __restrict__ a, const float* __restrict__ da, float* __restrict__ dx, float* __restrict__ dvec,
                                   float* __restrict__ dw, float* __restrict__ db, long long rows, int T, int H) {
    // the gradient of the 2H attention weights is a sum over ALL rows: it is accumulated in registers across a warp's rows, reduced over
    // the block's warps in shared memory and added once per block (per-row atomics = rows x 2H same-address atomics on 2H addresses)
    __shared__ float red[8][FFA_MAXC * 32];
    const int warps = blockDim.x >> 5, lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    float gx[FFA_MAXC], gk[FFA_MAXC];
#pragma unroll
    for (int q = 0; q < FFA_MAXC; ++q) { gx[q] = 0.f; gk[q] = 0.f; }
    float gb = 0.f;
    for (long long row = blockIdx.x * static_cast<long long>(warps) + wi; row < rows; row += static_cast<long long>(gridDim.x) * warps) {
        const float av = a[row];
        const float s = da[row] * av * (1.f - av);
        if (s == 0.f) continue;
        const long long ko = static_cast<long long>(__ldg(kw_idx + row / T)) * H;
#pragma unroll
        for (int q = 0; q < FFA_MAXC; ++q) {
            const int c = lane + 32 * q;
            if (c < H) {
                dx[row * H + c] += s * __ldg(w + c);
                atomicAdd(dvec + ko + c, s * __ldg(w + H + c));
                gx[q] = fmaf(s, ld1<AT>(x + row * H + c), gx[q]);
                gk[q] = fmaf(s, ld1<AT>(vec + ko + c), gk[q]);
            }
        }
        gb += s;
    }
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
        for (int q = 0; q < FFA_MAXC; ++q) { const int c = lane + 32 * q; if (c < H) red[wi][c] = pass == 0 ? gx[q] : gk[q]; }
        __syncthreads();
        for (int c = threadIdx.x; c < H; c += blockDim.x) {
            float t = 0.f;
            for (int w2 = 0; w2 < warps; ++w2) t += red[w2][c];
            if (t != 0.f) atomicAdd(dw + pass * H + c, t);
        }
        __syncthreads();
    }
    if (lane == 0 && gb != 0.f) atomicAdd(db, gb);
}

int launch_ff_attn_bwd(int dt, const void* x, const void* vec, const int* kw_idx, const float* w, const float* a, const float* da, float* dx,
                       float* dvec, float* dw, float* db, int n, int T, int H, cudaStream_t st) {
    if (n <= 0) return STAIR_OK;
    if (H > 32 * FFA_MAXC) return STAIR_ERR_UNSUPPORTED;
    const long long rows = static_cast<long long>(n) * T;
    DISPATCH_DT(dt, AT, (ff_attn_bwd_kernel<AT><<<nblocks(rows, 8, 148 * 2), 256, 0, st>>>(reinterpret_cast<const AT*>(x), reinterpret_cast<const AT*>(vec), kw_idx, w, a, da,
                                                                                           dx, dvec, dw, db, rows, T, H)));
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

template <typename AT>
__global__ void attnvideo_bwd_kernel(const float* __restrict__ dOut, const AT* __restrict__ vid, const int* __restrict__ feat_idx,
                                     const float* __restrict__ att, const int* __restrict__ att_idx, float* __restrict__ datt,
                                     float* __restrict__ dvid, long long rows, int T, int H) {
    const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
    for (long long row = blockIdx.x * static_cast<long long>(warps) + (threadIdx.x >> 5); row < rows; row += static_cast<long long>(gridDim.x) * warps) {
        const int i = static_cast<int>(row / T), t = static_cast<int>(row % T);
        const long long fo = (static_cast<long long>(__ldg(feat_idx + i)) * T + t) * H;
        const long long ar = static_cast<long long>(__ldg(att_idx + i)) * T + t;
        const float av = att[ar];
        float dot = 0.f;
        for (int c = lane; c < H; c += 32) {
            const float g = dOut[row * H + c];
            dot += g * ld1<AT>(vid + fo + c);
            if (g != 0.f) atomicAdd(dvid + fo + c, av * g);
        }
        dot = warp_sum(dot);
        if (lane == 0) atomicAdd(datt + ar, dot);
    }
}

int launch_attnvideo_bwd(int dt, const float* dOut, const void* vid, const int* feat_idx, const float* att, const int* att_idx, float* datt,
                         float* dvid, int n, int T, int H, cudaStream_t st) {
    if (n <= 0) return STAIR_OK;
    const long long rows = static_cast<long long>(n) * T;
    DISPATCH_DT(dt, AT, (attnvideo_bwd_kernel<AT><<<nblocks(rows, 8), 256, 0, st>>>(dOut, reinterpret_cast<const AT*>(vid), feat_idx, att, att_idx, datt, dvid, rows, T, H)));
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

// Relate backward: y = softmax(x +/- beta): dx = y (dy - sum y dy); dbeta = sign dx
__global__ void relate_bwd_kernel(const float* __restrict__ att_out, int out_base, const float* __restrict__ datt_out, const int* __restrict__ att_idx,
                                  float sign, float* __restrict__ datt, float* __restrict__ dbeta, int n, int T) {
    const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
    for (int i = blockIdx.x * warps + (threadIdx.x >> 5); i < n; i += gridDim.x * warps) {
        const float* y = att_out + static_cast<long long>(out_base + i) * T;
        const float* dy = datt_out + static_cast<long long>(out_base + i) * T;
        float s = 0.f;
        for (int t = lane; t < T; t += 32) s += y[t] * dy[t];
        s = warp_sum(s);
        float* dx = datt + static_cast<long long>(__ldg(att_idx + i)) * T;
        for (int t = lane; t < T; t += 32) {
            const float g = y[t] * (dy[t] - s);
            atomicAdd(dx + t, g);
            atomicAdd(dbeta + t, sign * g);
        }
    }
}

int launch_relate_bwd(const float* att_out, int out_base, const float* datt_out, const int* att_idx, int sign, float* datt, float* dbeta, int n, int T, cudaStream_t st) {
    if (n <= 0) return STAIR_OK;
    relate_bwd_kernel<<<nblocks(n, 8), 256, 0, st>>>(att_out, out_base, datt_out, att_idx, sign > 0 ? 1.f : (sign < 0 ? -1.f : 0.f), datt, dbeta, n, T);
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

// HasItem tail backward: a = sigmoid(w.x + b)
template <typename AT>
__global__ void rowdot_sigmoid_bwd_kernel(const AT* __restrict__ x, const float* __restrict__ w, const float* __restrict__ a, const float* __restrict__ da,
                                          float* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db, long long rows, int H, float dscale) {
    const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
    for (long long row = blockIdx.x * static_cast<long long>(warps) + (threadIdx.x >> 5); row < rows; row += static_cast<long long>(gridDim.x) * warps) {
        const float av = a[row];
        float s;
        if (dscale == 1.f) s = da[row] * av * (1.f - av);
        else if (av == 0.f) s = 0.f;                              // dropped by the Sigmoid -> Dropout of HasItem (sigmoid itself is never 0)
        else { const float sg = av / dscale; s = da[row] * dscale * sg * (1.f - sg); }
        for (int c = lane; c < H; c += 32) {
            dx[row * H + c] = s * __ldg(w + c);
            if (s != 0.f) atomicAdd(dw + c, s * ld1<AT>(x + row * H + c));
        }
        if (lane == 0 && s != 0.f) atomicAdd(db, s);
    }
}

int launch_rowdot_sigmoid_bwd(int dt, const void* x, const float* w, const float* a, const float* da, float* dx, float* dw, float* db,
                              long long rows, int H, float dscale, cudaStream_t st) {
    if (rows <= 0) return STAIR_OK;
    DISPATCH_DT(dt, AT, (rowdot_sigmoid_bwd_kernel<AT><<<nblocks(rows, 8), 256, 0, st>>>(reinterpret_cast<const AT*>(x), w, a, da, dx, dw, db, rows, H, dscale)));
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

template <typename AT>
__global__ void choose_bwd_kernel(const AT* __restrict__ vec, const int* __restrict__ k1, const int* __restrict__ k2, const int* __restrict__ q,
                                  const float* __restrict__ dOut, float* __restrict__ dvec, int n, int H) {
    const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
    for (int i = blockIdx.x * warps + (threadIdx.x >> 5); i < n; i += gridDim.x * warps) {
        const long long ao = static_cast<long long>(__ldg(k1 + i)) * H, bo = static_cast<long long>(__ldg(k2 + i)) * H, co = static_cast<long long>(__ldg(q + i)) * H;
        float aa = 0, bb = 0, cc = 0, ac = 0, bc = 0;
        for (int c = lane; c < H; c += 32) {
            const float x = ld1<AT>(vec + ao + c), y = ld1<AT>(vec + bo + c), z = ld1<AT>(vec + co + c);
            aa += x * x; bb += y * y; cc += z * z; ac += x * z; bc += y * z;
        }
        aa = warp_sum(aa); bb = warp_sum(bb); cc = warp_sum(cc); ac = warp_sum(ac); bc = warp_sum(bc);
        const float nq = fmaxf(sqrtf(cc), 1e-8f);
        const bool first = ac / (fmaxf(sqrtf(aa), 1e-8f) * nq) > bc / (fmaxf(sqrtf(bb), 1e-8f) * nq);
        float* d = dvec + (first ? ao : bo);
        for (int c = lane; c < H; c += 32) atomicAdd(d + c, dOut[static_cast<long long>(i) * H + c]);
    }
}

int launch_choose_bwd(int dt, const void* vec, const int* k1, const int* k2, const int* q, const float* dOut, float* dvec, int n, int H, cudaStream_t st) {
    if (n <= 0) return STAIR_OK;
    DISPATCH_DT(dt, AT, (choose_bwd_kernel<AT><<<nblocks(n, 8), 256, 0, st>>>(reinterpret_cast<const AT*>(vec), k1, k2, q, dOut, dvec, n, H)));
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

// torch.min(a,b): grad to the smaller, split in half on ties; |a-b|: sign(a-b) * grad
template <typename AT>
__global__ void binary_bwd_kernel(const AT* __restrict__ base, const int* __restrict__ a_idx, const int* __restrict__ b_idx, const float* __restrict__ dOut,
                                  float* __restrict__ dbase, int unit, int len, int op, long long total) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int r = static_cast<int>(i / len), e = static_cast<int>(i % len);
        const long long ao = static_cast<long long>(__ldg(a_idx + r)) * unit + e, bo = static_cast<long long>(__ldg(b_idx + r)) * unit + e;
        const float x = ld1<AT>(base + ao), y = ld1<AT>(base + bo), g = dOut[i];
        float ga, gb;
        if (op == STAIR_BIN_MIN) { ga = x < y ? g : (x == y ? 0.5f * g : 0.f); gb = y < x ? g : (x == y ? 0.5f * g : 0.f); }
        else { const float s = x > y ? 1.f : (x < y ? -1.f : 0.f); ga = s * g; gb = -s * g; }
        if (ga != 0.f) atomicAdd(dbase + ao, ga);
        if (gb != 0.f) atomicAdd(dbase + bo, gb);
    }
}

int launch_binary_bwd(int dt, const void* base, const int* a_idx, const int* b_idx, const float* dOut, float* dbase, int unit, int len, int op, int n, cudaStream_t st) {
    if (n <= 0) return STAIR_OK;
    const long long total = static_cast<long long>(n) * len;
    DISPATCH_DT(dt, AT, (binary_bwd_kernel<AT><<<nblocks(total, 256), 256, 0, st>>>(reinterpret_cast<const AT*>(base), a_idx, b_idx, dOut, dbase, unit, len, op, total)));
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

__global__ void array2_bwd_kernel(const float* __restrict__ dOut, const int* __restrict__ a_idx, const int* __restrict__ b_idx, float* __restrict__ dvec, long long total, int H) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / (2 * H);
        const int rem = static_cast<int>(i % (2 * H));
        const int which = rem / H, c = rem % H;
        atomicAdd(dvec + static_cast<long long>(__ldg((which ? b_idx : a_idx) + r)) * H + c, dOut[i]);
    }
}
int launch_array2_bwd(const float* dOut, const int* a_idx, const int* b_idx, float* dvec, int n, int H, cudaStream_t st) {
    if (n <= 0) return STAIR_OK;
    const long long total = static_cast<long long>(n) * 2 * H;
    array2_bwd_kernel<<<nblocks(total, 256), 256, 0, st>>>(dOut, a_idx, b_idx, dvec, total, H);
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

template <typename AT>
__global__ void concat_bwd_kernel(const AT* __restrict__ vec, const int* __restrict__ a_idx, const int* __restrict__ b_idx, int mode,
                                  const float* __restrict__ dcat, float* __restrict__ dvec, long long total, int H) {
    const int width = (mode == STAIR_CAT_PAIR ? 2 : 3) * H;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / H;
        const int c = static_cast<int>(i % H);
        const long long ao = static_cast<long long>(__ldg(a_idx + r)) * H + c, bo = static_cast<long long>(__ldg(b_idx + r)) * H + c;
        const float* d = dcat + r * width;
        float ga, gb;
        if (mode == STAIR_CAT_EXISTS) {            // cat = [b | a | b*a]
            const float a = ld1<AT>(vec + ao), b = ld1<AT>(vec + bo);
            ga = d[H + c] + d[2 * H + c] * b; gb = d[c] + d[2 * H + c] * a;
        } else if (mode == STAIR_CAT_XOR) {        // cat = [|a-b| | a | b]
            const float a = ld1<AT>(vec + ao), b = ld1<AT>(vec + bo);
            const float s = a > b ? 1.f : (a < b ? -1.f : 0.f);
            ga = s * d[c] + d[H + c]; gb = -s * d[c] + d[2 * H + c];
        } else { ga = d[c]; gb = d[H + c]; }
        atomicAdd(dvec + ao, ga);
        atomicAdd(dvec + bo, gb);
    }
}

int launch_concat_bwd(int dt, const void* vec, const int* a_idx, const int* b_idx, int mode, const float* dcat, float* dvec, int n, int H, cudaStream_t st) {
    if (n <= 0) return STAIR_OK;
    const long long total = static_cast<long long>(n) * H;
    DISPATCH_DT(dt, AT, (concat_bwd_kernel<AT><<<nblocks(total, 256), 256, 0, st>>>(reinterpret_cast<const AT*>(vec), a_idx, b_idx, mode, dcat, dvec, total, H)));
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

// Superlative tail backward (see super_mix_kernel)
template <typename AT>
__global__ void super_mix_bwd_kernel(const float* __restrict__ att, int K, int T, int H, int is_min, const AT* __restrict__ act_base,
                                     const int* __restrict__ act_idx, int act_unit, const float* __restrict__ dv, float* __restrict__ datt_s,
                                     float* __restrict__ dact_base, int n) {
    extern __shared__ float sm[];          // p[K] softmax, dw[K]
    float* p = sm; float* dw = sm + K;
    const int i = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
    const long long ro = static_cast<long long>(__ldg(act_idx + i)) * act_unit * H;
    for (int k = warp; k < K; k += warps) {
        float s = 0.f, d = 0.f;
        for (int t = lane; t < T; t += 32) s += att[(static_cast<long long>(i) * K + k) * T + t];
        for (int c = lane; c < H; c += 32) d += dv[static_cast<long long>(i) * H + c] * ld1<AT>(act_base + ro + static_cast<long long>(k) * H + c);
        s = warp_sum(s); d = warp_sum(d);
        if (lane == 0) { p[k] = s; dw[k] = is_min ? -d : d; }
    }
    __syncthreads();
    if (warp == 0) {
        float m = -INFINITY;
        for (int k = lane; k < K; k += 32) m = fmaxf(m, p[k]);
        m = warp_max(m);
        float s = 0.f;
        for (int k = lane; k < K; k += 32) s += expf(p[k] - m);
        s = warp_sum(s);
        float dot = 0.f;
        for (int k = lane; k < K; k += 32) { p[k] = expf(p[k] - m) / s; dot += p[k] * dw[k]; }
        dot = warp_sum(dot);
        for (int k = lane; k < K; k += 32) dw[k] = p[k] * (dw[k] - dot);        // ds_k
    }
    __syncthreads();
    for (int e = threadIdx.x; e < K * T; e += blockDim.x) datt_s[static_cast<long long>(i) * K * T + e] = dw[e / T];
    for (int e = threadIdx.x; e < K * H; e += blockDim.x) {
        const int k = e / H, c = e % H;
        const float wk = is_min ? 1.f - p[k] : p[k];
        atomicAdd(dact_base + ro + static_cast<long long>(k) * H + c, wk * dv[static_cast<long long>(i) * H + c]);
    }
}

int launch_super_mix_bwd(int dt, const float* att, int K, int T, int H, int is_min, const void* act_base, const int* act_idx, int act_unit,
                         const float* dv, float* datt_s, float* dact_base, int n, cudaStream_t st) {
    if (n <= 0) return STAIR_OK;
    DISPATCH_DT(dt, AT, (super_mix_bwd_kernel<AT><<<n, 128, 2 * K * sizeof(float), st>>>(att, K, T, H, is_min, reinterpret_cast<const AT*>(act_base), act_idx, act_unit,
                                                                                        dv, datt_s, dact_base, n)));
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

__global__ void word_embed_bwd_kernel(const float* __restrict__ dvec, int out_base, const int* __restrict__ q_off, const int* __restrict__ pos_q,
                                      const int* __restrict__ span_s, const int* __restrict__ span_e, float* __restrict__ dtokfeat, long long total, int H) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int r = static_cast<int>(i / H), c = static_cast<int>(i % H);
        const int q = __ldg(pos_q + r);
        const int base = __ldg(q_off + q), L = __ldg(q_off + q + 1) - base;
        int s = __ldg(span_s + r), e = __ldg(span_e + r);
        if (s < 0) { s = 0; e = L; }
        s = min(s, L); e = min(e, L);
        if (e <= s) continue;
        const float g = dvec[static_cast<long long>(out_base + r) * H + c] / static_cast<float>(e - s);
        if (g == 0.f) continue;
        for (int t = s; t < e; ++t) atomicAdd(dtokfeat + static_cast<long long>(base + t) * H + c, g);
    }
}

int launch_word_embed_bwd(const float* dvec, int out_base, const int* q_off, const int* pos_q, const int* span_s, const int* span_e,
                          float* dtokfeat, int n, int H, cudaStream_t st) {
    if (n <= 0) return STAIR_OK;
    const long long total = static_cast<long long>(n) * H;
    word_embed_bwd_kernel<<<nblocks(total, 256), 256, 0, st>>>(dvec, out_base, q_off, pos_q, span_s, span_e, dtokfeat, total, H);
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

__global__ void decoder_concat_bwd_kernel(const float* __restrict__ dcat, const int* __restrict__ root_node, const int* __restrict__ out_slot,
                                          float* __restrict__ dvec, float* __restrict__ dqfeat, long long total, int H) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long b = i / (2 * H);
        const int c = static_cast<int>(i % (2 * H));
        const float g = dcat[i];
        if (c < H) atomicAdd(dvec + static_cast<long long>(out_slot[__ldg(root_node + b)]) * H + c, g);
        else dqfeat[b * H + (c - H)] += g;
    }
}
int launch_decoder_concat_bwd(const float* dcat, const int* root_node, const int* out_slot, float* dvec, float* dqfeat, int B, int H, cudaStream_t st) {
    if (B <= 0) return STAIR_OK;
    const long long total = static_cast<long long>(B) * 2 * H;
    decoder_concat_bwd_kernel<<<nblocks(total, 256), 256, 0, st>>>(dcat, root_node, out_slot, dvec, dqfeat, total, H);
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// losses
// ------------------------------------------------------------------------------------------------------------------
// attention_score_criterion (train_module.py:83-90): mean over elements of -(g log p + (1-g) log(1-p)); w already holds
// module_loss_weight / gradient_accumulation / (#elements of the prediction).
__global__ void loss_att_kernel(const float* __restrict__ att, float* __restrict__ datt, const int* __restrict__ out_slot, const int* __restrict__ aux_slot,
                                const int* __restrict__ node, const int* __restrict__ kind, const int* __restrict__ slot, const float* __restrict__ gold,
                                const float* __restrict__ w, float* __restrict__ loss, int n, int T) {
    const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
    for (int i = blockIdx.x * warps + (threadIdx.x >> 5); i < n; i += gridDim.x * warps) {
        const int nd = __ldg(node + i), kd = __ldg(kind + i);
        const long long row = kd < 2 ? out_slot[nd] + kd : aux_slot[nd];
        const float wi = __ldg(w + i);
        float l = 0.f;
        for (int t = lane; t < T; t += 32) {
            const float p = att[row * T + t], g = __ldg(gold + static_cast<long long>(i) * T + t);
            l += -(g * logf(p) + (1.f - g) * logf(1.f - p));
            atomicAdd(datt + row * T + t, wi * (-(g / p) + (1.f - g) / (1.f - p)));
        }
        l = warp_sum(l);
        if (lane == 0) atomicAdd(loss + __ldg(slot + i), wi * l);
    }
}

int launch_loss_att(const float* att, float* datt, const int* out_slot, const int* aux_slot, const int* node, const int* kind, const int* slot,
                    const float* gold, const float* w, float* loss, int n, int T, cudaStream_t st) {
    if (n <= 0) return STAIR_OK;
    loss_att_kernel<<<nblocks(n, 8), 256, 0, st>>>(att, datt, out_slot, aux_slot, node, kind, slot, gold, w, loss, n, T);
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

struct HeadPtrs { const float* w[3]; const float* b[3]; float* dw[3]; float* db[3]; };

// criterion_exists / criterion_equals (train_module.py:92-107) + backward of the pretrain_head Linear
template <typename AT>
__global__ void loss_bin_kernel(const AT* __restrict__ vec, float* __restrict__ dvec, const int* __restrict__ out_slot, const int* __restrict__ node,
                                const int* __restrict__ label, const float* __restrict__ w, HeadPtrs hp, const int* __restrict__ which,
                                float* __restrict__ loss, int n, int H) {
    const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
    for (int i = blockIdx.x * warps + (threadIdx.x >> 5); i < n; i += gridDim.x * warps) {
        const int wh = __ldg(which + i);                          // 0 Equals (1 logit, MSE), 1 Xor, 2 Exists (2 logits, CE)
        const long long ro = static_cast<long long>(out_slot[__ldg(node + i)]) * H;
        const float* W = hp.w[wh];
        const int nout = wh == 0 ? 1 : 2;
        float z0 = 0.f, z1 = 0.f;
        for (int c = lane; c < H; c += 32) {
            const float x = ld1<AT>(vec + ro + c);
            z0 += x * __ldg(W + c);
            if (nout == 2) z1 += x * __ldg(W + H + c);
        }
        z0 = warp_sum(z0) + __ldg(hp.b[wh]);
        z1 = nout == 2 ? warp_sum(z1) + __ldg(hp.b[wh] + 1) : 0.f;
        const float wi = __ldg(w + i);
        const int lb = __ldg(label + i);
        float l, d0, d1 = 0.f;
        if (wh == 0) { const float e = z0 - static_cast<float>(lb); l = e * e; d0 = 2.f * e; }
        else {
            const float m = fmaxf(z0, z1);
            const float e0 = expf(z0 - m), e1 = expf(z1 - m), s = e0 + e1;
            l = logf(s) + m - (lb ? z1 : z0);
            d0 = e0 / s - (lb ? 0.f : 1.f); d1 = e1 / s - (lb ? 1.f : 0.f);
        }
        d0 *= wi; d1 *= wi;
        for (int c = lane; c < H; c += 32) {
            const float x = ld1<AT>(vec + ro + c);
            float g = d0 * __ldg(W + c);
            atomicAdd(hp.dw[wh] + c, d0 * x);
            if (nout == 2) { g += d1 * __ldg(W + H + c); atomicAdd(hp.dw[wh] + H + c, d1 * x); }
            atomicAdd(dvec + ro + c, g);
        }
        if (lane == 0) {
            atomicAdd(hp.db[wh], d0);
            if (nout == 2) atomicAdd(hp.db[wh] + 1, d1);
            atomicAdd(loss + (wh == 0 ? 4 : 3), wi * l);
        }
    }
}

int launch_loss_bin(int dt, const void* vec, float* dvec, const int* out_slot, const int* node, const int* is_mse, const int* label, const float* w,
                    const float* const* head_w, const float* const* head_b, float* const* dhead_w, float* const* dhead_b, const int* which,
                    float* loss, int n, int H, cudaStream_t st) {
    (void)is_mse;
    if (n <= 0) return STAIR_OK;
    HeadPtrs hp;
    for (int j = 0; j < 3; ++j) { hp.w[j] = head_w[j]; hp.b[j] = head_b[j]; hp.dw[j] = dhead_w[j]; hp.db[j] = dhead_b[j]; }
    DISPATCH_DT(dt, AT, (loss_bin_kernel<AT><<<nblocks(n, 8), 256, 0, st>>>(reinterpret_cast<const AT*>(vec), dvec, out_slot, node, label, w, hp, which, loss, n, H)));
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

// contrastive CE (train_module.py:113-132): p = normalize(x); s_j = p . G_j over all classes of the window; loss = lse(s) - s_pos
int g_loss_con_impl = 0;           // 0 = shared-memory kernel when the window has <= 64 classes (product); 1 = register kernel always
constexpr int CON_MAXC = 32;       // columns per lane: H <= 1024
constexpr int CON_R = 4;           // rows per warp: every element of G fetched from L2/L1 serves CON_R rows (the class matrix, n_cls x H fp32,
                                   // is re-read for every row: at one row per warp the kernel is bound by that traffic)
template <typename AT, int MAXC>
__global__ void loss_con_kernel(const AT* __restrict__ vec, float* __restrict__ dvec, const int* __restrict__ out_slot, const int* __restrict__ node,
                                const int* __restrict__ pos, const float* __restrict__ w, const float* __restrict__ G, int n_cls,
                                float* __restrict__ loss, int n, int H) {
    extern __shared__ float sc[];                                  // [warps][CON_R][n_cls] class scores of the warp's current rows
    const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
    float* s = sc + static_cast<size_t>(threadIdx.x >> 5) * CON_R * n_cls;
    float lsum = 0.f;
    for (int i0 = (blockIdx.x * warps + (threadIdx.x >> 5)) * CON_R; i0 < n; i0 += gridDim.x * warps * CON_R) {
        long long ro[CON_R];
        float x[CON_R][MAXC], nrm[CON_R];
#pragma unroll
        for (int r = 0; r < CON_R; ++r) {
            const int i = min(i0 + r, n - 1);                        // tail rows recompute the last row (their results are discarded)
            ro[r] = static_cast<long long>(out_slot[__ldg(node + i)]) * H;
            float ss = 0.f;
#pragma unroll
            for (int q = 0; q < MAXC; ++q) { const int c = lane + 32 * q; x[r][q] = c < H ? ld1<AT>(vec + ro[r] + c) : 0.f; ss += x[r][q] * x[r][q]; }
            nrm[r] = fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
        }
        float m[CON_R];
#pragma unroll
        for (int r = 0; r < CON_R; ++r) m[r] = -INFINITY;
        for (int j = 0; j < n_cls; ++j) {
            const float* Gj = G + static_cast<long long>(j) * H;
            float d[CON_R];
#pragma unroll
            for (int r = 0; r < CON_R; ++r) d[r] = 0.f;
#pragma unroll
            for (int q = 0; q < MAXC; ++q) {
                const int c = lane + 32 * q;
                if (c < H) {
                    const float g = __ldg(Gj + c);
#pragma unroll
                    for (int r = 0; r < CON_R; ++r) d[r] = fmaf(x[r][q], g, d[r]);
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int r = 0; r < CON_R; ++r) d[r] += __shfl_xor_sync(0xffffffffu, d[r], o);
#pragma unroll
            for (int r = 0; r < CON_R; ++r) { const float v = d[r] / nrm[r]; if (lane == 0) s[r * n_cls + j] = v; m[r] = fmaxf(m[r], v); }
        }
        __syncwarp();
        float wi[CON_R];
        int ps[CON_R];
#pragma unroll
        for (int r = 0; r < CON_R; ++r) {
            float se = 0.f;
            for (int j = lane; j < n_cls; j += 32) se += expf(s[r * n_cls + j] - m[r]);
            se = warp_sum(se);
            const float lse = m[r] + logf(se);
            const int i = min(i0 + r, n - 1);
            ps[r] = __ldg(pos + i);
            wi[r] = i0 + r < n ? __ldg(w + i) : 0.f;
            lsum += wi[r] * (lse - s[r * n_cls + ps[r]]);
            __syncwarp();
            for (int j = lane; j < n_cls; j += 32) s[r * n_cls + j] = expf(s[r * n_cls + j] - lse);      // softmax probabilities
        }
        __syncwarp();
        float dp[CON_R][MAXC];
#pragma unroll
        for (int r = 0; r < CON_R; ++r)
#pragma unroll
            for (int q = 0; q < MAXC; ++q) { const int c = lane + 32 * q; dp[r][q] = c < H ? -__ldg(G + static_cast<long long>(ps[r]) * H + c) : 0.f; }
        for (int j = 0; j < n_cls; ++j) {
            const float* Gj = G + static_cast<long long>(j) * H;
            float pj[CON_R];
#pragma unroll
            for (int r = 0; r < CON_R; ++r) pj[r] = s[r * n_cls + j];
#pragma unroll
            for (int q = 0; q < MAXC; ++q) {
                const int c = lane + 32 * q;
                if (c < H) {
                    const float g = __ldg(Gj + c);
#pragma unroll
                    for (int r = 0; r < CON_R; ++r) dp[r][q] = fmaf(pj[r], g, dp[r][q]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < CON_R; ++r) {
            float pdp = 0.f;
#pragma unroll
            for (int q = 0; q < MAXC; ++q) pdp += dp[r][q] * x[r][q] / nrm[r];
            pdp = warp_sum(pdp);
            if (i0 + r < n) {
#pragma unroll
                for (int q = 0; q < MAXC; ++q) {
                    const int c = lane + 32 * q;
                    if (c < H) atomicAdd(dvec + ro[r] + c, wi[r] * (dp[r][q] - x[r][q] / nrm[r] * pdp) / nrm[r]);
                }
            }
        }
        __syncwarp();
    }
    if (lane == 0 && lsum != 0.f) atomicAdd(loss + 5, lsum);
}


// Shared-memory form of the contrastive CE for the common window (n_cls <= 64 classes, H <= 512, H % 32 == 0): the class matrix is staged
// once per CTA ([64][H + 1] fp32, conflict-free both by class and by column), a warp takes four rows at a time.  Scores: lane = class
// (lane, lane + 32), the normalised row is broadcast from shared memory — no shuffle reductions; gradient: lane = column.  With
// a_j = softmax_j - [j == pos]: d x_hat = sum_j a_j G_j and x_hat . d x_hat = sum_j a_j s_j (s_j = x_hat . G_j), so the projection term of
// normalize's backward needs no second sweep.  (The register form below executes 38 K warp instructions per four rows, this one ~12 K.)
constexpr int CON_S_CLS = 64, CON_S_WARPS = 8;
template <typename AT>
__global__ void __launch_bounds__(CON_S_WARPS * 32) loss_con_smem_kernel(const AT* __restrict__ vec, float* __restrict__ dvec, const int* __restrict__ out_slot,
                                                                           const int* __restrict__ node, const int* __restrict__ pos, const float* __restrict__ w,
                                                                           const float* __restrict__ G, int n_cls, float* __restrict__ loss, int n, int H) {
    extern __shared__ float sm[];
    const int GS = H + 1;
    float* sG = sm;                                                    // [64][H + 1]
    float* sx = sG + CON_S_CLS * GS + ((CON_S_CLS * GS) & 3 ? 4 - ((CON_S_CLS * GS) & 3) : 0);      // [warps][4][H] normalised rows (16-byte aligned)
    float* sa = sx + CON_S_WARPS * CON_R * H;                          // [warps][64][4] a_j of the warp's four rows
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < CON_S_CLS * H; i += blockDim.x) {
        const int j = i / H, c = i - j * H;
        sG[j * GS + c] = j < n_cls ? __ldg(G + static_cast<long long>(j) * H + c) : 0.f;
    }
    __syncthreads();
    float* x = sx + warp * CON_R * H;
    float* a = sa + warp * CON_S_CLS * CON_R;
    const int nq = H >> 5;
    float lsum = 0.f;
    for (int i0 = (blockIdx.x * CON_S_WARPS + warp) * CON_R; i0 < n; i0 += gridDim.x * CON_S_WARPS * CON_R) {
        long long ro[CON_R]; float rn[CON_R];
#pragma unroll
        for (int r = 0; r < CON_R; ++r) {
            const int i = min(i0 + r, n - 1);                          // tail rows recompute the last row (their results are discarded)
            ro[r] = static_cast<long long>(out_slot[__ldg(node + i)]) * H;
            float ss = 0.f;
            for (int q = 0; q < nq; ++q) { const float v = ld1<AT>(vec + ro[r] + lane + 32 * q); x[r * H + lane + 32 * q] = v; ss = fmaf(v, v, ss); }
            rn[r] = 1.0f / fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
        }
        __syncwarp();
        // scores of classes (lane, lane + 32) for the four rows
        float s0[CON_R], s1[CON_R];
#pragma unroll
        for (int r = 0; r < CON_R; ++r) { s0[r] = 0.f; s1[r] = 0.f; }
        const float* g0 = sG + lane * GS;
        const float* g1 = sG + (lane + 32) * GS;
#pragma unroll 2
        for (int c = 0; c < H; c += 4) {
            float4 xv[CON_R];
#pragma unroll
            for (int r = 0; r < CON_R; ++r) xv[r] = *reinterpret_cast<const float4*>(x + r * H + c);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float ga = g0[c + k], gb = g1[c + k];
#pragma unroll
                for (int r = 0; r < CON_R; ++r) {
                    const float xk = k == 0 ? xv[r].x : k == 1 ? xv[r].y : k == 2 ? xv[r].z : xv[r].w;
                    s0[r] = fmaf(xk, ga, s0[r]); s1[r] = fmaf(xk, gb, s1[r]);
                }
            }
        }
        float wi[CON_R], pdp[CON_R];
#pragma unroll
        for (int r = 0; r < CON_R; ++r) {
            const float v0 = lane < n_cls ? s0[r] * rn[r] : -INFINITY, v1 = lane + 32 < n_cls ? s1[r] * rn[r] : -INFINITY;
            const float m = warp_max(fmaxf(v0, v1));
            const float se = warp_sum(expf(v0 - m) + expf(v1 - m));
            const float lse = m + logf(se);
            const int i = min(i0 + r, n - 1);
            const int ps = __ldg(pos + i);
            wi[r] = i0 + r < n ? __ldg(w + i) : 0.f;
            const float sp = __shfl_sync(0xffffffffu, ps < 32 ? v0 : v1, ps & 31);
            lsum += wi[r] * (lse - sp);
            const float a0 = (lane < n_cls ? expf(v0 - lse) : 0.f) - (ps == lane ? 1.f : 0.f);
            const float a1 = (lane + 32 < n_cls ? expf(v1 - lse) : 0.f) - (ps == lane + 32 ? 1.f : 0.f);
            a[lane * CON_R + r] = a0; a[(lane + 32) * CON_R + r] = a1;
            pdp[r] = warp_sum((lane < n_cls ? a0 * v0 : 0.f) + (lane + 32 < n_cls ? a1 * v1 : 0.f));
        }
        __syncwarp();
        // d x_hat[c] = sum_j a_j G[j][c], lane = column (8 columns per sweep), then normalize's backward and the scatter
        for (int q0 = 0; q0 < nq; q0 += 8) {
            float dp[CON_R][8];
#pragma unroll
            for (int r = 0; r < CON_R; ++r)
#pragma unroll
                for (int q = 0; q < 8; ++q) dp[r][q] = 0.f;
#pragma unroll 2
            for (int j = 0; j < n_cls; ++j) {
                const float4 aj = *reinterpret_cast<const float4*>(a + j * CON_R);
                const float* gj = sG + j * GS + lane + 32 * q0;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (q0 + q < nq) {
                        const float g = gj[32 * q];
                        dp[0][q] = fmaf(aj.x, g, dp[0][q]); dp[1][q] = fmaf(aj.y, g, dp[1][q]);
                        dp[2][q] = fmaf(aj.z, g, dp[2][q]); dp[3][q] = fmaf(aj.w, g, dp[3][q]);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < CON_R; ++r) {
                if (i0 + r < n) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        if (q0 + q < nq) {
                            const int c = lane + 32 * (q0 + q);
                            atomicAdd(dvec + ro[r] + c, wi[r] * (dp[r][q] - x[r * H + c] * rn[r] * pdp[r]) * rn[r]);
                        }
                    }
                }
            }
        }
        __syncwarp();
    }
    if (lane == 0 && lsum != 0.f) atomicAdd(loss + 5, lsum);          // every lane holds the same sum
}

int launch_loss_con(int dt, const void* vec, float* dvec, const int* out_slot, const int* node, const int* pos, const float* w,
                    const float* cls_rep, int n_cls, float* loss, int n, int H, cudaStream_t st) {
    if (n <= 0 || n_cls <= 0) return STAIR_OK;
    if (H > 32 * CON_MAXC || n_cls > 1024) return STAIR_ERR_UNSUPPORTED;
    static_assert(CON_R == 4, "loss_con_smem_kernel keeps the a_j of a warp's rows as one float4");
    if (g_loss_con_impl == 0 && n_cls <= CON_S_CLS && H <= 512 && (H % 32) == 0) {
        const size_t smem_s = (static_cast<size_t>(CON_S_CLS) * (H + 1) + 4 + CON_S_WARPS * CON_R * H + CON_S_WARPS * CON_S_CLS * CON_R) * sizeof(float);
        const int groups = (n + CON_S_WARPS * CON_R - 1) / (CON_S_WARPS * CON_R);
        const int grid_s = groups < 148 ? groups : 148;
        static size_t configured[2] = {0, 0};
        const int di = dt == STAIR_BF16 ? 0 : 1;
        if (configured[di] < smem_s) {
            cudaError_t e = dt == STAIR_BF16 ? cudaFuncSetAttribute(loss_con_smem_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_s))
                                             : cudaFuncSetAttribute(loss_con_smem_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_s));
            if (e != cudaSuccess) return STAIR_ERR_CUDA;
            configured[di] = smem_s;
        }
        DISPATCH_DT(dt, AT, (loss_con_smem_kernel<AT><<<grid_s, CON_S_WARPS * 32, smem_s, st>>>(reinterpret_cast<const AT*>(vec), dvec, out_slot, node, pos, w, cls_rep,
                                                                                               n_cls, loss, n, H)));
        STAIR_CHECK_LAUNCH();
        return STAIR_OK;
    }
    const int warps = 4;
    const size_t smem = static_cast<size_t>(warps) * CON_R * n_cls * sizeof(float);
    const int grid = nblocks((n + CON_R - 1) / CON_R, warps);
    if (H <= 512) {
        DISPATCH_DT(dt, AT, (loss_con_kernel<AT, 16><<<grid, warps * 32, smem, st>>>(reinterpret_cast<const AT*>(vec), dvec, out_slot, node, pos, w, cls_rep,
                                                                                    n_cls, loss, n, H)));
    } else {
        DISPATCH_DT(dt, AT, (loss_con_kernel<AT, CON_MAXC><<<grid, warps * 32, smem, st>>>(reinterpret_cast<const AT*>(vec), dvec, out_slot, node, pos, w, cls_rep,
                                                                                          n_cls, loss, n, H)));
    }
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

// criterion_filterframe (train_module.py:141-155): BCELoss(softmax_O(z), gold), mean over the [T, O] elements (folded into w).
// One warp per (supervised node, frame).  BCELoss clamps log() at -100 and its backward divides by max(p (1 - p), 1e-12) (ATen).
__global__ void loss_ff_kernel(const float* __restrict__ head, float* __restrict__ dhead, const int* __restrict__ aux_slot, const int* __restrict__ node,
                               const float* __restrict__ gold, const float* __restrict__ w, float* __restrict__ loss, int n, int T, int O) {
    const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
    float lsum = 0.f;
    for (long long r = blockIdx.x * static_cast<long long>(warps) + (threadIdx.x >> 5); r < static_cast<long long>(n) * T; r += static_cast<long long>(gridDim.x) * warps) {
        const int i = static_cast<int>(r / T), t = static_cast<int>(r % T);
        const long long hr = (static_cast<long long>(aux_slot[__ldg(node + i)]) * T + t) * O;
        const float* z = head + hr;
        const float* g = gold + r * O;
        const float wi = __ldg(w + i);
        float m = -INFINITY;
        for (int j = lane; j < O; j += 32) m = fmaxf(m, z[j]);
        m = warp_max(m);
        float se = 0.f;
        for (int j = lane; j < O; j += 32) se += expf(z[j] - m);
        se = warp_sum(se);
        float l = 0.f, dot = 0.f;                                  // dot = sum_k dL/dp_k p_k
        for (int j = lane; j < O; j += 32) {
            const float p = expf(z[j] - m) / se, gj = g[j];
            l -= gj * fmaxf(logf(p), -100.f) + (1.f - gj) * fmaxf(logf(1.f - p), -100.f);
            dot += (p - gj) / fmaxf(p * (1.f - p), 1e-12f) * p;
        }
        l = warp_sum(l); dot = warp_sum(dot);
        for (int j = lane; j < O; j += 32) {
            const float p = expf(z[j] - m) / se, gj = g[j];
            dhead[hr + j] = wi * p * ((p - gj) / fmaxf(p * (1.f - p), 1e-12f) - dot);
        }
        lsum += wi * l;
    }
    if (lane == 0 && lsum != 0.f) atomicAdd(loss + 7, lsum);
}

int launch_loss_ff(const float* head, float* dhead, const int* aux_slot, const int* node, const float* gold, const float* w, float* loss,
                   int n, int T, int O, cudaStream_t st) {
    if (n <= 0) return STAIR_OK;
    loss_ff_kernel<<<nblocks(static_cast<long long>(n) * T, 8), 256, 0, st>>>(head, dhead, aux_slot, node, gold, w, loss, n, T, O);
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

// db[c] += sum_rows x[row][c] for a bf16 matrix (bias gradient of the encoder input projections from the bf16 gate gradients);
// a block owns 64 columns x a slab of rows, 4 row-groups of 64 threads, one atomic per column and block
__global__ void colsum_bf16_kernel(const bf16* __restrict__ x, long long rows, int cols, long long ld, float* __restrict__ db, long long rows_per_block) {
    __shared__ float red[4][64];
    const int cl = threadIdx.x & 63, rg = threadIdx.x >> 6;
    const int c = blockIdx.x * 64 + cl;
    const long long r0 = blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    float s = 0.f;
    if (c < cols)
        for (long long r = r0 + rg; r < r1; r += 4) s += __bfloat162float(x[r * ld + c]);
    red[rg][cl] = s;
    __syncthreads();
    if (rg == 0 && c < cols) {
        const float t = red[0][cl] + red[1][cl] + red[2][cl] + red[3][cl];
        if (t != 0.f) atomicAdd(db + c, t);
    }
}
int launch_colsum_bf16(const bf16* x, long long rows, int cols, long long ld, float* db, cudaStream_t st) {
    if (rows <= 0 || cols <= 0 || !db) return STAIR_OK;
    const int gx = (cols + 63) / 64;
    int gy = static_cast<int>((148 * 8 + gx - 1) / gx);
    if (gy > rows / 64 + 1) gy = static_cast<int>(rows / 64 + 1);
    const long long rpb = (rows + gy - 1) / gy;
    colsum_bf16_kernel<<<dim3(gx, gy), 256, 0, st>>>(x, rows, cols, ld, db, rpb);
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

__global__ void add_inplace_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n4) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        float4 a = reinterpret_cast<float4*>(dst)[i];
        const float4 b = reinterpret_cast<const float4*>(src)[i];
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        reinterpret_cast<float4*>(dst)[i] = a;
    }
}
__device__ __forceinline__ float to_float(float x) { return x; }
__device__ __forceinline__ float to_float(bf16 x) { return __bfloat162float(x); }

// ---- pretrain-head backward for external gradient seeds (StairTrain.ext_dhead_small / ext_dhead_vec) --------------------------------
template <typename AT>
__global__ void small_head_bwd_kernel(const AT* __restrict__ vec, int row_base, const float* __restrict__ w, int nout, const float* __restrict__ dout,
                                      int out_base, float* __restrict__ dvec, float* __restrict__ dW, float* __restrict__ db, int n, int H) {
    const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
    for (int i = blockIdx.x * warps + (threadIdx.x >> 5); i < n; i += gridDim.x * warps) {
        const AT* x = vec + static_cast<long long>(row_base + i) * H;
        float* dx = dvec + static_cast<long long>(row_base + i) * H;
        for (int o = 0; o < nout; ++o) {
            const float g = dout[static_cast<long long>(out_base + i) * 2 + o];
            if (g == 0.0f) continue;
            for (int c = lane; c < H; c += 32) {
                atomicAdd(dx + c, g * __ldg(w + o * H + c));
                if (dW) atomicAdd(dW + o * H + c, g * to_float(x[c]));
            }
            if (db && lane == 0) atomicAdd(db + o, g);
        }
    }
}

int launch_small_head_bwd(int dt, const void* vec, int row_base, const float* w, int nout, const float* dout, int out_base, float* dvec, float* dW,
                          float* db, int n, int H, cudaStream_t st) {
    if (n <= 0) return STAIR_OK;
    if (nout > 2) return STAIR_ERR_ARG;
    DISPATCH_DT(dt, AT, (small_head_bwd_kernel<AT><<<nblocks(n, 8, 148 * 8), 256, 0, st>>>(reinterpret_cast<const AT*>(vec), row_base, w, nout, dout,
                                                                                                    out_base, dvec, dW, db, n, H)));
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

template <typename AT>
__global__ void l2norm_bwd_kernel(const AT* __restrict__ vec, int row_base, const float* __restrict__ dout, int out_base, float* __restrict__ dvec, int n, int H) {
    const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
    for (int i = blockIdx.x * warps + (threadIdx.x >> 5); i < n; i += gridDim.x * warps) {
        const AT* x = vec + static_cast<long long>(row_base + i) * H;
        const float* g = dout + static_cast<long long>(out_base + i) * H;
        float ss = 0.f, xg = 0.f;
        for (int c = lane; c < H; c += 32) { const float xv = to_float(x[c]); ss += xv * xv; xg += xv * g[c]; }
        ss = warp_sum(ss); xg = warp_sum(xg);
        const float nrm = sqrtf(ss);
        float* dx = dvec + static_cast<long long>(row_base + i) * H;
        if (nrm > 1e-12f) {                                    // y = x / |x| :  dx = (g - y (y.g)) / |x|
            const float inv = 1.0f / nrm, k = xg * inv * inv;
            for (int c = lane; c < H; c += 32) atomicAdd(dx + c, (g[c] - to_float(x[c]) * k) * inv);
        } else {                                               // F.normalize clamps the norm at eps: y = x / eps
            for (int c = lane; c < H; c += 32) atomicAdd(dx + c, g[c] * 1e12f);
        }
    }
}

int launch_l2norm_bwd(int dt, const void* vec, int row_base, const float* dout, int out_base, float* dvec, int n, int H, cudaStream_t st) {
    if (n <= 0) return STAIR_OK;
    DISPATCH_DT(dt, AT, (l2norm_bwd_kernel<AT><<<nblocks(n, 8, 148 * 8), 256, 0, st>>>(reinterpret_cast<const AT*>(vec), row_base, dout, out_base, dvec, n, H)));
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

int launch_add_inplace(float* dst, const float* src, long long n, cudaStream_t st) {
    if (n <= 0) return STAIR_OK;
    if (n % 4) return STAIR_ERR_ARG;
    add_inplace_kernel<<<nblocks(n / 4, 256), 256, 0, st>>>(dst, src, n / 4);
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

__global__ void loss_dec_kernel(const float* __restrict__ logits, const int* __restrict__ answer, float w, float* __restrict__ dlogits,
                                float* __restrict__ loss, int B, int A) {
    const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
    for (int b = blockIdx.x * warps + (threadIdx.x >> 5); b < B; b += gridDim.x * warps) {
        const float* z = logits + static_cast<long long>(b) * A;
        float m = -INFINITY;
        for (int a = lane; a < A; a += 32) m = fmaxf(m, z[a]);
        m = warp_max(m);
        float s = 0.f;
        for (int a = lane; a < A; a += 32) s += expf(z[a] - m);
        s = warp_sum(s);
        const int y = __ldg(answer + b);
        for (int a = lane; a < A; a += 32) dlogits[static_cast<long long>(b) * A + a] = w * (expf(z[a] - m) / s - (a == y ? 1.f : 0.f));
        if (lane == 0) atomicAdd(loss + 6, w * (logf(s) + m - z[y]));
    }
}

int launch_loss_dec(const float* logits, const int* answer, float w, float* dlogits, float* loss, int B, int A, cudaStream_t st) {
    if (B <= 0) return STAIR_OK;
    loss_dec_kernel<<<nblocks(B, 8), 256, 0, st>>>(logits, answer, w, dlogits, loss, B, A);
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// LSTM with history + BPTT cell
// ------------------------------------------------------------------------------------------------------------------
template <typename XT, typename OT>
__global__ void lstm_cell_train_kernel(const XT* __restrict__ xproj, const float* __restrict__ g, const float* __restrict__ c_prev,
                                       float* __restrict__ c_out, float* __restrict__ gates_out, bf16* __restrict__ hs_out,
                                       const bf16* __restrict__ hs_prev, int nplanes, long long hs_plane, long long hs_dir, OT* __restrict__ out, OT* __restrict__ qfeat,
                                       const int* __restrict__ q_off, int B, int T, int h, int step) {
    const long long total = 2LL * B * h;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int d = static_cast<int>(i / (static_cast<long long>(B) * h));
        const long long rem = i % (static_cast<long long>(B) * h);
        const int b = static_cast<int>(rem / h), j = static_cast<int>(rem % h);
        const long long si = (static_cast<long long>(d) * B + b) * h + j;
        const long long hi = static_cast<long long>(d) * hs_dir + static_cast<long long>(b) * h + j;     // state history is direction-major
        long long row; bool active = true;
        if (q_off) {
            const int base = __ldg(q_off + b), L = __ldg(q_off + b + 1) - base;
            active = step < L;
            row = base + (d == 0 ? step : L - 1 - step);
        } else row = static_cast<long long>(b) * T + (d == 0 ? step : T - 1 - step);
        if (!active) {
            c_out[si] = c_prev[si];
            for (int p = 0; p < nplanes; ++p) hs_out[p * hs_plane + hi] = hs_prev[p * hs_plane + hi];
            for (int gt = 0; gt < 4; ++gt) gates_out[(static_cast<long long>(d) * B + b) * 4 * h + gt * h + j] = 0.f;
            continue;
        }
        const XT* x = xproj + row * 8 * h + d * 4 * h + j;
        float pi = ld1<XT>(x), pf = ld1<XT>(x + h), pg = ld1<XT>(x + 2 * h), po = ld1<XT>(x + 3 * h), cp = 0.f;
        if (step > 0) {
            const float* gr = g + (static_cast<long long>(d) * B + b) * 4 * h + j;
            pi += gr[0]; pf += gr[h]; pg += gr[2 * h]; po += gr[3 * h];
            cp = c_prev[si];
        }
        const float ig = sigmoidf_(pi), fg = sigmoidf_(pf), gg = tanhf(pg), og = sigmoidf_(po);
        const float cn = fg * cp + ig * gg;
        const float hn = og * tanhf(cn);
        c_out[si] = cn;
        float* go = gates_out + (static_cast<long long>(d) * B + b) * 4 * h + j;
        go[0] = ig; go[h] = fg; go[2 * h] = gg; go[3 * h] = og;
        st1<OT>(out + row * 2 * h + d * h + j, hn);
        if (qfeat) st1<OT>(qfeat + static_cast<long long>(b) * 2 * h + d * h + j, hn);
        store_planes1(hn, hs_out + hi, hs_plane, nplanes);
    }
}

int launch_lstm_cell_train(int xdt, const void* xproj, const float* g, const float* c_prev, float* c_out, float* gates_out, bf16* hstate_out,
                           const bf16* hstate_prev, long long hs_plane, long long hs_dir, int nplanes, int odt, void* out, void* qfeat, const int* q_off,
                           int B, int T, int h, int step, cudaStream_t st) {
    if (B <= 0) return STAIR_OK;
    if (xdt != odt) return STAIR_ERR_ARG;
    const long long total = 2LL * B * h;
    const int grid = nblocks(total, 256);
    if (xdt == STAIR_BF16)
        lstm_cell_train_kernel<bf16, bf16><<<grid, 256, 0, st>>>(reinterpret_cast<const bf16*>(xproj), g, c_prev, c_out, gates_out, hstate_out, hstate_prev,
                                                                nplanes, hs_plane, hs_dir, reinterpret_cast<bf16*>(out), reinterpret_cast<bf16*>(qfeat), q_off, B, T, h, step);
    else
        lstm_cell_train_kernel<float, float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(xproj), g, c_prev, c_out, gates_out, hstate_out, hstate_prev,
                                                                  nplanes, hs_plane, hs_dir, reinterpret_cast<float*>(out), reinterpret_cast<float*>(qfeat), q_off, B, T, h, step);
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

// One BPTT step of both directions.  dout rows = gradient of the encoder output (video: dvid slot rows; text: dtokfeat rows).
__global__ void lstm_cell_bwd_kernel(const float* __restrict__ gates, const float* __restrict__ c_prev, const float* __restrict__ c_cur,
                                     const float* __restrict__ dout, const float* __restrict__ dh_rec, const float* __restrict__ dqfeat,
                                     float* __restrict__ dc, long long dg_dir, bf16* __restrict__ dg_planes, long long dg_plane, int nplanes,
                                     float* __restrict__ dxproj, bf16* __restrict__ dxproj_bf16, const int* __restrict__ q_off, int B, int T, int h, int step,
                                     int last_step, int blocked) {
    const long long total = 2LL * B * h;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int d = static_cast<int>(i / (static_cast<long long>(B) * h));
        const long long rem = i % (static_cast<long long>(B) * h);
        const int b = static_cast<int>(rem / h), j = static_cast<int>(rem % h);
        const long long si = (static_cast<long long>(d) * B + b) * h + j;
        const long long gi = (static_cast<long long>(d) * B + b) * 4 * h + j;
        long long row; bool active = true; bool final_step = false;
        if (q_off) {
            const int base = __ldg(q_off + b), L = __ldg(q_off + b + 1) - base;
            active = step < L;
            final_step = step == L - 1;
            row = base + (d == 0 ? step : L - 1 - step);
        } else row = static_cast<long long>(b) * T + (d == 0 ? step : T - 1 - step);
        float dpi = 0.f, dpf = 0.f, dpg = 0.f, dpo = 0.f;
        if (active) {
            float dh = dout[row * 2 * h + d * h + j];
            if (step < last_step) dh += dh_rec[si];
            if (dqfeat && final_step) dh += dqfeat[static_cast<long long>(b) * 2 * h + d * h + j];
            if (blocked) {                                           // bf16 coefficient history of the fused forward (train_kernels.cuh)
                const bf16* co = reinterpret_cast<const bf16*>(gates) + lstm_hist_coef_off(step, d, b, 0, j, B, h);
                const float dct = (step < last_step ? dc[si] : 0.f) + dh * __bfloat162float(co[LSTM_CO_A * 256]);
                dpi = dct * __bfloat162float(co[LSTM_CO_BI * 256]);
                dpf = dct * __bfloat162float(co[LSTM_CO_BF * 256]);
                dpg = dct * __bfloat162float(co[LSTM_CO_BG * 256]);
                dpo = dh * __bfloat162float(co[LSTM_CO_BO * 256]);
                dc[si] = dct * __bfloat162float(co[LSTM_CO_F * 256]);
            } else {
                const float ig = gates[gi], fg = gates[gi + h], gg = gates[gi + 2 * h], og = gates[gi + 3 * h];
                const float ccur = c_cur[si], cp = step > 0 ? c_prev[si] : 0.f;
                const float tc = tanhf(ccur);
                const float dct = (step < last_step ? dc[si] : 0.f) + dh * og * (1.f - tc * tc);
                dpi = dct * gg * ig * (1.f - ig);
                dpf = dct * cp * fg * (1.f - fg);
                dpg = dct * ig * (1.f - gg * gg);
                dpo = dh * tc * og * (1.f - og);
                dc[si] = dct * fg;
            }
            if (dxproj_bf16) {                                   // bf16 path: the row-major bf16 copy IS the weight-gradient GEMM operand
                bf16* xb = dxproj_bf16 + row * 8 * h + d * 4 * h + j;
                xb[0] = __float2bfloat16_rn(dpi); xb[h] = __float2bfloat16_rn(dpf); xb[2 * h] = __float2bfloat16_rn(dpg); xb[3 * h] = __float2bfloat16_rn(dpo);
            } else {
                float* xr = dxproj + row * 8 * h + d * 4 * h + j;
                xr[0] = dpi; xr[h] = dpf; xr[2 * h] = dpg; xr[3 * h] = dpo;
            }
        } else {
            dc[si] = 0.f;
        }
        // gate gradients of this step as bf16 planes, written straight into the direction-major history [np][2][S][B][4h]: the step's
        // slice is the A operand of the recurrent GEMM (dh = dG . W_hh) and the whole history the MN-major operand of dW_hh += dG^T . h
        bf16* dgp = dg_planes + static_cast<long long>(d) * dg_dir + static_cast<long long>(b) * 4 * h + j;
        store_planes1(dpi, dgp, dg_plane, nplanes);
        store_planes1(dpf, dgp + h, dg_plane, nplanes);
        store_planes1(dpg, dgp + 2 * h, dg_plane, nplanes);
        store_planes1(dpo, dgp + 3 * h, dg_plane, nplanes);
    }
}

int launch_lstm_cell_bwd(const float* gates, const float* c_prev, const float* c_cur, const float* dout, const float* dh_rec, const float* dqfeat,
                         float* dc, long long dg_dir, bf16* dg_planes, long long dg_plane, int nplanes, float* dxproj, bf16* dxproj_bf16, const int* q_off,
                         int B, int T, int h, int step, int last_step, int blocked, cudaStream_t st) {
    if (B <= 0) return STAIR_OK;
    if (blocked && (h % 8)) return STAIR_ERR_ARG;
    lstm_cell_bwd_kernel<<<nblocks(2LL * B * h, 256), 256, 0, st>>>(gates, c_prev, c_cur, dout, dh_rec, dqfeat, dc, dg_dir, dg_planes, dg_plane, nplanes,
                                                                   dxproj, dxproj_bf16, q_off, B, T, h, step, last_step, blocked);
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

// Adam (torch.optim.Adam semantics, weight_decay 0, train_module.py:326-332)
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
                            float lr, float2 b1, float2 b2, float eps, float bc1, float bc2) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float gi = g[i];
        const float mi = b1.x * m[i] + b1.y * gi;
        const float vi = b2.x * v[i] + b2.y * gi * gi;
        m[i] = mi; v[i] = vi;
        p[i] -= (lr / bc1) * mi / (sqrtf(vi) / sqrtf(bc2) + eps);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Multi-tensor Adam + weight-copy refresh (see StairAdamSeg in include/stair_b200.h)
// ------------------------------------------------------------------------------------------------------------------
// b1 / b2 arrive as {beta, 1 - beta} pairs with 1 - beta rounded from double (torch computes 1 - beta in double: in fp32, 1 - 0.999f is
// off by 1.3e-5 relative, which the second moment inherits)
__device__ __forceinline__ float adam_update(float& p, float g, float& m, float& v, float lr, float2 b1, float2 b2, float eps, float bc1, float bc2) {
    const float mi = b1.x * m + b1.y * g;
    const float vi = b2.x * v + b2.y * g * g;
    m = mi; v = vi;
    p -= (lr / bc1) * mi / (sqrtf(vi) / sqrtf(bc2) + eps);
    return p;
}

__global__ void __launch_bounds__(256) adam_multi_kernel(const StairAdamSeg* __restrict__ segs, int n_segs, float lr, float2 b1, float2 b2, float eps) {
    __shared__ unsigned short tile[3][64][66];
    __shared__ StairAdamSeg sg;
    // segment of this block's tile: last seg with tile0 <= blockIdx.x
    if (threadIdx.x == 0) {
        int lo = 0, hi = n_segs - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (segs[mid].tile0 <= static_cast<int>(blockIdx.x)) lo = mid; else hi = mid - 1;
        }
        sg = segs[lo];
    }
    __syncthreads();
    const int t = static_cast<int>(blockIdx.x) - sg.tile0;
    if (sg.kind == 0) {                                          // fp32 vector: 1024 elements per tile
        const long long n = static_cast<long long>(sg.rows) * sg.cols;
        float* dst = reinterpret_cast<float*>(sg.packed);
        for (int k = 0; k < 4; ++k) {
            const long long i = static_cast<long long>(t) * 1024 + k * 256 + threadIdx.x;
            if (i >= n) break;
            float pv = sg.p[i], mv = sg.m[i], vv = sg.v[i];
            float val = adam_update(pv, sg.g[i], mv, vv, lr, b1, b2, eps, sg.bc1, sg.bc2);
            sg.p[i] = pv; sg.m[i] = mv; sg.v[i] = vv;
            if (sg.p2) {
                float p2 = sg.p2[i], m2 = sg.m2[i], v2 = sg.v2[i];
                val += adam_update(p2, sg.g2[i], m2, v2, lr, b1, b2, eps, sg.bc1, sg.bc2);
                sg.p2[i] = p2; sg.m2[i] = m2; sg.v2[i] = v2;
            }
            if (dst) dst[i] = val;
        }
        return;
    }
    const int tiles_c = (sg.cols + 63) >> 6;
    const int r0 = (t / tiles_c) * 64, c0 = (t % tiles_c) * 64;
    const int tr = threadIdx.x >> 3, tc = (threadIdx.x & 7) * 8;
    const bool vec = (sg.cols & 3) == 0;
    bf16* pk = reinterpret_cast<bf16*>(sg.packed);
    bf16* pp = reinterpret_cast<bf16*>(sg.packed_perm);
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        const int r = r0 + tr + 32 * pass, c = c0 + tc;
        float w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = 0.f;
        if (r < sg.rows && c < sg.cols) {
            const long long o = static_cast<long long>(r) * sg.cols + c;
            if (vec && c + 8 <= sg.cols) {
                float4 pa = *reinterpret_cast<const float4*>(sg.p + o), pb = *reinterpret_cast<const float4*>(sg.p + o + 4);
                const float4 ga = *reinterpret_cast<const float4*>(sg.g + o), gb = *reinterpret_cast<const float4*>(sg.g + o + 4);
                float4 ma = *reinterpret_cast<const float4*>(sg.m + o), mb = *reinterpret_cast<const float4*>(sg.m + o + 4);
                float4 va = *reinterpret_cast<const float4*>(sg.v + o), vb = *reinterpret_cast<const float4*>(sg.v + o + 4);
                w[0] = adam_update(pa.x, ga.x, ma.x, va.x, lr, b1, b2, eps, sg.bc1, sg.bc2);
                w[1] = adam_update(pa.y, ga.y, ma.y, va.y, lr, b1, b2, eps, sg.bc1, sg.bc2);
                w[2] = adam_update(pa.z, ga.z, ma.z, va.z, lr, b1, b2, eps, sg.bc1, sg.bc2);
                w[3] = adam_update(pa.w, ga.w, ma.w, va.w, lr, b1, b2, eps, sg.bc1, sg.bc2);
                w[4] = adam_update(pb.x, gb.x, mb.x, vb.x, lr, b1, b2, eps, sg.bc1, sg.bc2);
                w[5] = adam_update(pb.y, gb.y, mb.y, vb.y, lr, b1, b2, eps, sg.bc1, sg.bc2);
                w[6] = adam_update(pb.z, gb.z, mb.z, vb.z, lr, b1, b2, eps, sg.bc1, sg.bc2);
                w[7] = adam_update(pb.w, gb.w, mb.w, vb.w, lr, b1, b2, eps, sg.bc1, sg.bc2);
                *reinterpret_cast<float4*>(sg.p + o) = pa; *reinterpret_cast<float4*>(sg.p + o + 4) = pb;
                *reinterpret_cast<float4*>(sg.m + o) = ma; *reinterpret_cast<float4*>(sg.m + o + 4) = mb;
                *reinterpret_cast<float4*>(sg.v + o) = va; *reinterpret_cast<float4*>(sg.v + o + 4) = vb;
            } else {
                for (int j = 0; j < 8 && c + j < sg.cols; ++j) {
                    float pv = sg.p[o + j], mv = sg.m[o + j], vv = sg.v[o + j];
                    w[j] = adam_update(pv, sg.g[o + j], mv, vv, lr, b1, b2, eps, sg.bc1, sg.bc2);
                    sg.p[o + j] = pv; sg.m[o + j] = mv; sg.v[o + j] = vv;
                }
            }
            // bf16 planes of the updated weights: row-major copy (+ gate-interleaved copy), staged in smem for the transposed copy
            int rp = r;
            if (pp) { const int hh = sg.perm_hh, gi = r / hh, rem = r % hh; rp = (rem >> 6) * 256 + gi * 64 + (rem & 63); }
            for (int pl = 0; pl < sg.nplanes; ++pl) {
                unsigned short hv[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const bf16 hb = __float2bfloat16_rn(w[j]);
                    hv[j] = *reinterpret_cast<const unsigned short*>(&hb);
                    w[j] -= __bfloat162float(hb);                 // residual for the next plane (exact bf16x3 split)
                    tile[pl][tr + 32 * pass][tc + j] = hv[j];
                }
                const uint4 q = make_uint4(hv[0] | (static_cast<uint32_t>(hv[1]) << 16), hv[2] | (static_cast<uint32_t>(hv[3]) << 16),
                                           hv[4] | (static_cast<uint32_t>(hv[5]) << 16), hv[6] | (static_cast<uint32_t>(hv[7]) << 16));
                bf16* d = pk + pl * sg.packed_plane + static_cast<long long>(r) * sg.packed_ld + c;
                if (c + 8 <= sg.cols && (sg.packed_ld & 7) == 0) *reinterpret_cast<uint4*>(d) = q;
                else for (int j = 0; j < 8 && c + j < sg.cols; ++j) reinterpret_cast<unsigned short*>(d)[j] = hv[j];
                if (pp && pl == 0) {
                    bf16* e = pp + static_cast<long long>(rp) * sg.packed_ld + c;
                    if (c + 8 <= sg.cols && (sg.packed_ld & 7) == 0) *reinterpret_cast<uint4*>(e) = q;
                    else for (int j = 0; j < 8 && c + j < sg.cols; ++j) reinterpret_cast<unsigned short*>(e)[j] = hv[j];
                }
            }
        } else {
            for (int pl = 0; pl < sg.nplanes; ++pl)
#pragma unroll
                for (int j = 0; j < 8; ++j) tile[pl][tr + 32 * pass][tc + j] = 0;
        }
    }
    if (!sg.packed_t) return;                                    // block-uniform
    __syncthreads();
    unsigned short* pt = reinterpret_cast<unsigned short*>(sg.packed_t);
    for (int pl = 0; pl < sg.nplanes; ++pl)
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            const int c = c0 + tr + 32 * pass;                   // row of the transposed copy = source column
            const int r = r0 + tc;                               // 8 consecutive destination columns = source rows
            if (c >= sg.cols || r >= sg.rows) continue;
            unsigned short hv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) hv[j] = tile[pl][tc + j][tr + 32 * pass];
            unsigned short* d = pt + pl * sg.packed_t_plane + static_cast<long long>(c) * sg.packed_t_ld + r;
            if (r + 8 <= sg.rows && (sg.packed_t_ld & 7) == 0 && (reinterpret_cast<uintptr_t>(d) & 15) == 0)
                *reinterpret_cast<uint4*>(d) = make_uint4(hv[0] | (static_cast<uint32_t>(hv[1]) << 16), hv[2] | (static_cast<uint32_t>(hv[3]) << 16),
                                                          hv[4] | (static_cast<uint32_t>(hv[5]) << 16), hv[6] | (static_cast<uint32_t>(hv[7]) << 16));
            else for (int j = 0; j < 8 && r + j < sg.rows; ++j) d[j] = hv[j];
        }
}

int launch_adam_multi(const StairAdamSeg* segs, int n_segs, int total_tiles, float lr, double b1, double b2, float eps, cudaStream_t st) {
    if (n_segs <= 0 || total_tiles <= 0) return STAIR_OK;
    adam_multi_kernel<<<total_tiles, 256, 0, st>>>(segs, n_segs, lr, make_float2(static_cast<float>(b1), static_cast<float>(1.0 - b1)),
                                                   make_float2(static_cast<float>(b2), static_cast<float>(1.0 - b2)), eps);
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

int launch_adam(float* p, const float* g, float* m, float* v, long long n, float lr, double b1, double b2, float eps, float bc1, float bc2, cudaStream_t st) {
    if (n <= 0) return STAIR_OK;
    adam_kernel<<<nblocks(n, 256), 256, 0, st>>>(p, g, m, v, n, lr, make_float2(static_cast<float>(b1), static_cast<float>(1.0 - b1)),
                                                 make_float2(static_cast<float>(b2), static_cast<float>(1.0 - b2)), eps, bc1, bc2);
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

}  // namespace stair
