// BiLSTM cell updates of the video / text encoders (nn.LSTM, 1 layer, bidirectional; video_nmn/module_net.py:39-47,
// 147-163).  Gate order i,f,g,o; pre-activations = xproj (W_ih x + b_ih + b_hh, one big tcgen05 GEMM over all frames /
// tokens) + g (W_hh h_prev, one GEMM per direction per step).  These kernels do the pointwise part of a step for both
// directions at once, write h to the encoder output and to the bf16 state that is the A operand of the next step.
#include "nmn_kernels.cuh"
#include <type_traits>

namespace stair {

template <typename XT>
__device__ __forceinline__ void lstm_cell8(const XT* xrow, const float* grow, int h, float* c, float (&hout)[8]) {
    Vec8<XT> xi, xf, xg, xo;
    xi.load(xrow); xf.load(xrow + h); xg.load(xrow + 2 * h); xo.load(xrow + 3 * h);
    Vec8<float> gi, gf, gg, go, cv;
    if (grow) { gi.load(grow); gf.load(grow + h); gg.load(grow + 2 * h); go.load(grow + 3 * h); cv.load(c); }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float pi = xi.v[j] + (grow ? gi.v[j] : 0.f), pf = xf.v[j] + (grow ? gf.v[j] : 0.f);
        const float pg = xg.v[j] + (grow ? gg.v[j] : 0.f), po = xo.v[j] + (grow ? go.v[j] : 0.f);
        const float cn = sigmoidf_(pf) * (grow ? cv.v[j] : 0.f) + sigmoidf_(pi) * tanhf(pg);
        cv.v[j] = cn;
        hout[j] = sigmoidf_(po) * tanhf(cn);
    }
    cv.store(c);
}

__device__ __forceinline__ void store_h_state(const float (&hv)[8], bf16* dst, long long plane_stride, int nplanes) {
    uint4 r[3];
    bf16* p0 = reinterpret_cast<bf16*>(&r[0]); bf16* p1 = reinterpret_cast<bf16*>(&r[1]); bf16* p2 = reinterpret_cast<bf16*>(&r[2]);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        p0[j] = __float2bfloat16_rn(hv[j]);
        float rem = hv[j] - __bfloat162float(p0[j]);
        p1[j] = __float2bfloat16_rn(rem);
        rem -= __bfloat162float(p1[j]);
        p2[j] = __float2bfloat16_rn(rem);
    }
    for (int p = 0; p < nplanes; ++p) *reinterpret_cast<uint4*>(dst + p * plane_stride) = r[p];
}

template <typename XT, typename OT>
__global__ void lstm_cell_video_kernel(const XT* __restrict__ xproj, const float* __restrict__ g, float* __restrict__ c,
                                       bf16* __restrict__ hstate, int nplanes, OT* __restrict__ out, int B, int T, int h, int step) {
    const int hc = h / 8;
    const long long total = 2LL * B * hc;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int d = static_cast<int>(i / (static_cast<long long>(B) * hc));
        const long long rem = i % (static_cast<long long>(B) * hc);
        const int b = static_cast<int>(rem / hc), j = static_cast<int>(rem % hc) * 8;
        const int t = d == 0 ? step : T - 1 - step;
        const XT* xrow = xproj + (static_cast<long long>(b) * T + t) * 8 * h + d * 4 * h + j;
        const float* grow = step == 0 ? nullptr : g + (static_cast<long long>(d) * B + b) * 4 * h + j;
        float hv[8];
        lstm_cell8<XT>(xrow, grow, h, c + (static_cast<long long>(d) * B + b) * h + j, hv);
        Vec8<OT> o;
#pragma unroll
        for (int k = 0; k < 8; ++k) o.v[k] = hv[k];
        o.store(out + (static_cast<long long>(b) * T + t) * 2 * h + d * h + j);
        store_h_state(hv, hstate + (static_cast<long long>(d) * B + b) * h + j, 2LL * B * h, nplanes);
    }
}

template <typename XT, typename OT>
__global__ void lstm_cell_text_kernel(const XT* __restrict__ xproj, const float* __restrict__ g, float* __restrict__ c,
                                      bf16* __restrict__ hstate, int nplanes, OT* __restrict__ tokfeat, OT* __restrict__ qfeat,
                                      const int* __restrict__ q_off, int B, int h, int step) {
    const int hc = h / 8;
    const long long total = 2LL * B * hc;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int d = static_cast<int>(i / (static_cast<long long>(B) * hc));
        const long long rem = i % (static_cast<long long>(B) * hc);
        const int b = static_cast<int>(rem / hc), j = static_cast<int>(rem % hc) * 8;
        const int base = __ldg(q_off + b), L = __ldg(q_off + b + 1) - base;
        if (step >= L) continue;                    // this question is finished: state and final h stay
        const long long row = base + (d == 0 ? step : L - 1 - step);
        const XT* xrow = xproj + row * 8 * h + d * 4 * h + j;
        const float* grow = step == 0 ? nullptr : g + (static_cast<long long>(d) * B + b) * 4 * h + j;
        float hv[8];
        lstm_cell8<XT>(xrow, grow, h, c + (static_cast<long long>(d) * B + b) * h + j, hv);
        Vec8<OT> o;
#pragma unroll
        for (int k = 0; k < 8; ++k) o.v[k] = hv[k];
        o.store(tokfeat + row * 2 * h + d * h + j);
        o.store(qfeat + static_cast<long long>(b) * 2 * h + d * h + j);     // h_n: last write wins (module_net.py:157-158)
        store_h_state(hv, hstate + (static_cast<long long>(d) * B + b) * h + j, 2LL * B * h, nplanes);
    }
}

template <typename F>
static int dispatch2(int xdt, int odt, F&& f) {
    if (xdt == STAIR_BF16 && odt == STAIR_BF16) return f(static_cast<bf16*>(nullptr), static_cast<bf16*>(nullptr));
    if (xdt == STAIR_F32 && odt == STAIR_F32) return f(static_cast<float*>(nullptr), static_cast<float*>(nullptr));
    if (xdt == STAIR_F32 && odt == STAIR_BF16) return f(static_cast<float*>(nullptr), static_cast<bf16*>(nullptr));
    return f(static_cast<bf16*>(nullptr), static_cast<float*>(nullptr));
}

int launch_lstm_cell_video(int xdt, const void* xproj, const float* g, float* c, bf16* hstate, int nplanes, int odt, void* out,
                           int B, int T, int h, int step, cudaStream_t st) {
    if (B <= 0) return STAIR_OK;
    if (h % 8) return STAIR_ERR_UNSUPPORTED;
    const long long total = 2LL * B * (h / 8);
    const int grid = static_cast<int>(min((total + 127) / 128, 148LL * 16));
    return dispatch2(xdt, odt, [&](auto* xp, auto* op) {
        typedef typename std::remove_pointer<decltype(xp)>::type XT;
        typedef typename std::remove_pointer<decltype(op)>::type OT;
        lstm_cell_video_kernel<XT, OT><<<grid, 128, 0, st>>>(reinterpret_cast<const XT*>(xproj), g, c, hstate, nplanes,
                                                            reinterpret_cast<OT*>(out), B, T, h, step);
        STAIR_CHECK_LAUNCH();
        return STAIR_OK;
    });
}

int launch_lstm_cell_text(int xdt, const void* xproj, const float* g, float* c, bf16* hstate, int nplanes, int odt, void* tokfeat,
                          void* qfeat, const int* q_off, int B, int h, int step, cudaStream_t st) {
    if (B <= 0) return STAIR_OK;
    if (h % 8) return STAIR_ERR_UNSUPPORTED;
    const long long total = 2LL * B * (h / 8);
    const int grid = static_cast<int>(min((total + 127) / 128, 148LL * 16));
    return dispatch2(xdt, odt, [&](auto* xp, auto* op) {
        typedef typename std::remove_pointer<decltype(xp)>::type XT;
        typedef typename std::remove_pointer<decltype(op)>::type OT;
        lstm_cell_text_kernel<XT, OT><<<grid, 128, 0, st>>>(reinterpret_cast<const XT*>(xproj), g, c, hstate, nplanes,
                                                           reinterpret_cast<OT*>(tokfeat), reinterpret_cast<OT*>(qfeat), q_off, B, h, step);
        STAIR_CHECK_LAUNCH();
        return STAIR_OK;
    });
}

}  // namespace stair
