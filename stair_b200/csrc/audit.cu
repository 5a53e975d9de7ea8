// Filter-audit head — the step right after the hot path (evaluate.py:65-117): cosine similarity of every
// Filter output against the phrase representations of the audit vocabulary and the top-k phrases per output:
//   gold_sims = nn.CosineSimilarity()(result.unsqueeze(0), filter_ans_reps)      (evaluate.py:107; eps 1e-8 on each norm)
//   gold_ranks = torch.argsort(gold_sims, descending=True)[:10]                 (evaluate.py:108-109)
// One block per query row: warps stride over the phrases (dot + phrase norm with 16-byte loads), similarities stay in shared
// memory, then k rounds of block-wide argmax (ties -> lowest phrase index).
#include "nmn_kernels.cuh"

namespace stair {
namespace {

template <typename QT>
__global__ void cosine_topk_kernel(const QT* __restrict__ q, long long ldq, const int* __restrict__ row_idx, const float* __restrict__ reps,
                                   int P, int H, int k, int* __restrict__ out_idx, float* __restrict__ out_sim) {
    extern __shared__ float sims[];                    // [P] + reduction scratch
    __shared__ float s_val[32];
    __shared__ int s_idx[32];
    __shared__ float s_qn;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
    const long long r = row_idx ? __ldg(row_idx + blockIdx.x) : blockIdx.x;
    const QT* qr = q + r * ldq;
    if (warp == 0) {
        float ss = 0.f;
        for (int c = lane; c < H; c += 32) { const float x = ld1<QT>(qr + c); ss += x * x; }
        ss = warp_sum(ss);
        if (lane == 0) s_qn = fmaxf(sqrtf(ss), 1e-8f);
    }
    __syncthreads();
    const float qn = s_qn;
    const int hc = H / 8;
    for (int p = warp; p < P; p += warps) {
        const float* rp = reps + static_cast<long long>(p) * H;
        float dot = 0.f, rr = 0.f;
        for (int c = lane; c < hc; c += 32) {
            Vec8<float> y; y.load(rp + c * 8);
            Vec8<QT> x; x.load(qr + c * 8);
#pragma unroll
            for (int j = 0; j < 8; ++j) { dot += x.v[j] * y.v[j]; rr += y.v[j] * y.v[j]; }
        }
        dot = warp_sum(dot); rr = warp_sum(rr);
        if (lane == 0) sims[p] = dot / (qn * fmaxf(sqrtf(rr), 1e-8f));
    }
    __syncthreads();
    for (int j = 0; j < k; ++j) {
        float best = -INFINITY; int bi = 0x7fffffff;
        for (int p = threadIdx.x; p < P; p += blockDim.x) {
            const float v = sims[p];
            if (v > best || (v == best && p < bi)) { best = v; bi = p; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) { s_val[warp] = best; s_idx[warp] = bi; }
        __syncthreads();
        if (warp == 0) {
            best = lane < warps ? s_val[lane] : -INFINITY;
            bi = lane < warps ? s_idx[lane] : 0x7fffffff;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            if (lane == 0) {
                const bool ok = bi != 0x7fffffff;
                out_idx[static_cast<long long>(blockIdx.x) * k + j] = ok ? bi : -1;
                out_sim[static_cast<long long>(blockIdx.x) * k + j] = best;
                if (ok) sims[bi] = -INFINITY;
            }
        }
        __syncthreads();
    }
}

}  // namespace
}  // namespace stair

using namespace stair;

// q: rows of `dtype` with pitch ldq (row i = q[row_idx[i]] or q[i] when row_idx is NULL); reps fp32 [P, H]; out_idx / out_sim [n, k]
extern "C" int stair_cosine_topk(int dtype, const void* q, long long ldq, const int32_t* row_idx, const float* reps, int P, int H, int k,
                                 int32_t* out_idx, float* out_sim, int n, void* stream) {
    if (n <= 0) return STAIR_OK;
    if (P <= 0 || H <= 0 || H % 8 || k <= 0 || k > P || P > 12000) return STAIR_ERR_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t smem = static_cast<size_t>(P) * sizeof(float);
    if (dtype == STAIR_BF16)
        cosine_topk_kernel<bf16><<<n, 128, smem, st>>>(reinterpret_cast<const bf16*>(q), ldq, row_idx, reps, P, H, k, out_idx, out_sim);
    else
        cosine_topk_kernel<float><<<n, 128, smem, st>>>(reinterpret_cast<const float*>(q), ldq, row_idx, reps, P, H, k, out_idx, out_sim);
    STAIR_CHECK_LAUNCH();
    return STAIR_OK;
}
