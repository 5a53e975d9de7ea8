// Raw-feature ingest (the step right before the hot path; reference: video_nmn/dataset.py:134-172, host/numpy there).
//   RX / TGIF-QA style:  video[b, t, 0:Da]     = mean_f appearance[b, t, f, :]      (torch.tensor(feat).mean(dim=1), :150-152)
//                        video[b, t, Da:Da+Dm] = motion[b, t, :]                    (torch.cat(..., dim=-1), :161-172)
//   I3D npy style:       video[b, t, :]        = feats[b, 2 t, :], t < T            (np.arange(0, n, 2) then [:max_video_length], :138-141)
// Both are pure HBM streams (RX: 557 KB read per question at fp32 -> 65 KB written): one thread owns 8 consecutive columns of one
// (question, clip) and keeps all F frame loads (16-byte each) in flight before summing them in frame order.
#include "nmn_kernels.cuh"

namespace stair {
namespace {

template <typename T> struct Ld8 {};
template <> struct Ld8<float> {
    static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
};
template <> struct Ld8<bf16> {
    static __device__ __forceinline__ void load(const bf16* p, float (&v)[8]) {
        const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
    }
};
template <typename T> __device__ __forceinline__ void st8(T* p, const float (&v)[8]);
template <> __device__ __forceinline__ void st8<float>(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void st8<bf16>(bf16* p, const float (&v)[8]) {
    uint4 r; __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = r;
}

// work item = 8 output columns of one (b, t) row; items [0, Da/8) pool the appearance frames, items [Da/8, (Da+Dm)/8) copy motion
template <typename IT, typename OT, int FU>
__global__ void ingest_pool_concat_kernel(const IT* __restrict__ app, const IT* __restrict__ motion, OT* __restrict__ out, long long rows, int F,
                                          int Da, int Dm, long long app_row_stride, long long motion_row_stride) {
    const int ca = Da / 8, cm = Dm / 8, cpr = ca + cm;
    const long long total = rows * cpr;
    const float inv = 1.0f / static_cast<float>(F);
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long row = i / cpr;
        const int c = static_cast<int>(i - row * cpr);
        float acc[8];
        if (c < ca) {
            const IT* src = app + row * app_row_stride + c * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = 0.f;
            int f = 0;
            for (; f + FU <= F; f += FU) {
                float v[FU][8];
#pragma unroll
                for (int u = 0; u < FU; ++u) Ld8<IT>::load(src + static_cast<long long>(f + u) * Da, v[u]);
#pragma unroll
                for (int u = 0; u < FU; ++u)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[j] += v[u][j];
            }
            for (; f < F; ++f) {
                float v[8];
                Ld8<IT>::load(src + static_cast<long long>(f) * Da, v);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] += v[j];
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] *= inv;
        } else {
            Ld8<IT>::load(motion + row * motion_row_stride + (c - ca) * 8, acc);
        }
        st8<OT>(out + row * (Da + Dm) + c * 8, acc);
    }
}

template <typename IT, typename OT>
__global__ void ingest_stride_kernel(const IT* __restrict__ src, OT* __restrict__ out, int B, int T, int D, int step, long long src_video_stride) {
    const int cpr = D / 8;
    const long long total = static_cast<long long>(B) * T * cpr;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long row = i / cpr;
        const int c = static_cast<int>(i - row * cpr);
        const long long b = row / T;
        const int t = static_cast<int>(row - b * T);
        float v[8];
        Ld8<IT>::load(src + b * src_video_stride + static_cast<long long>(t) * step * D + c * 8, v);
        st8<OT>(out + row * D + c * 8, v);
    }
}

inline int grid_for(long long items) {
    long long b = (items + 255) / 256;
    if (b < 1) b = 1;
    return static_cast<int>(b > 148LL * 32 ? 148 * 32 : b);
}

}  // namespace
}  // namespace stair

using namespace stair;

#define INGEST_DISPATCH(idt, odt, ...)                                                                   \
    do {                                                                                                 \
        if ((idt) == STAIR_F32 && (odt) == STAIR_BF16) { typedef float IT; typedef bf16 OT; __VA_ARGS__; } \
        else if ((idt) == STAIR_F32 && (odt) == STAIR_F32) { typedef float IT; typedef float OT; __VA_ARGS__; } \
        else if ((idt) == STAIR_BF16 && (odt) == STAIR_BF16) { typedef bf16 IT; typedef bf16 OT; __VA_ARGS__; } \
        else { typedef bf16 IT; typedef float OT; __VA_ARGS__; }                                         \
    } while (0)

// appearance [B, T, F, Da], motion [B, T, Dm] (or null with Dm = 0) -> out [B, T, Da + Dm]
extern "C" int stair_ingest_pool_concat(const void* appearance, const void* motion, int in_dtype, void* out, int out_dtype, int B, int T, int F,
                                        int Da, int Dm, void* stream) {
    if (B <= 0 || T <= 0) return STAIR_OK;
    if (F <= 0 || Da <= 0 || Da % 8 || Dm % 8 || Dm < 0 || (Dm > 0 && !motion)) return STAIR_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(appearance) & 15) || (reinterpret_cast<uintptr_t>(motion) & 15) || (reinterpret_cast<uintptr_t>(out) & 15))
        return STAIR_ERR_ARG;
    const long long rows = static_cast<long long>(B) * T;
    const long long items = rows * ((Da + Dm) / 8);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    INGEST_DISPATCH(in_dtype, out_dtype, (ingest_pool_concat_kernel<IT, OT, 8><<<grid_for(items), 256, 0, st>>>(
                                             reinterpret_cast<const IT*>(appearance), reinterpret_cast<const IT*>(motion), reinterpret_cast<OT*>(out),
                                             rows, F, Da, Dm, static_cast<long long>(F) * Da, Dm)));
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}

// feats [B, n_frames, D] -> out [B, T, D] with out[b, t] = feats[b, t * step]   (needs (T - 1) * step < n_frames)
extern "C" int stair_ingest_subsample(const void* feats, int in_dtype, void* out, int out_dtype, int B, int n_frames, int T, int D, int step,
                                      void* stream) {
    if (B <= 0 || T <= 0) return STAIR_OK;
    if (D <= 0 || D % 8 || step <= 0 || static_cast<long long>(T - 1) * step >= n_frames) return STAIR_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(feats) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return STAIR_ERR_ARG;
    const long long items = static_cast<long long>(B) * T * (D / 8);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    INGEST_DISPATCH(in_dtype, out_dtype, (ingest_stride_kernel<IT, OT><<<grid_for(items), 256, 0, st>>>(
                                             reinterpret_cast<const IT*>(feats), reinterpret_cast<OT*>(out), B, T, D, step,
                                             static_cast<long long>(n_frames) * D)));
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}
