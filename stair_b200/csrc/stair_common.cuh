// Shared device/host helpers for the stair_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "stair_b200.h"

namespace stair {

typedef __nv_bfloat16 bf16;

static inline int cuda_status(cudaError_t e) { return e == cudaSuccess ? STAIR_OK : STAIR_ERR_CUDA; }
extern thread_local long long g_launch_count;     // kernels launched by this thread (stair_last_launch_count)
#define STAIR_CHECK_LAUNCH() do { ++::stair::g_launch_count; cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return STAIR_ERR_CUDA; } while (0)
#define STAIR_TRY(expr) do { int rc__ = (expr); if (rc__ != STAIR_OK) return rc__; } while (0)

__host__ __device__ static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- typed element access: activations are bf16 (default) or fp32 (strict mode) ----------------
template <typename T> struct Vec8;      // 8 consecutive elements, loaded with 16-byte transactions
template <> struct Vec8<float> {
    float v[8];
    __device__ __forceinline__ void load(const float* p) {
        float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    __device__ __forceinline__ void load_stream(const float* p) {      // read-once streams: do not allocate in L1
        float4 a, b;
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(p));
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p + 4));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    __device__ __forceinline__ void store(float* p) const {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
};
template <> struct Vec8<bf16> {
    float v[8];
    __device__ __forceinline__ void load(const bf16* p) {
        uint4 r = *reinterpret_cast<const uint4*>(p);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
        for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
    }
    __device__ __forceinline__ void load_stream(const bf16* p) {
        uint4 r;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
        for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
    }
    __device__ __forceinline__ void store(bf16* p) const {
        uint4 r; __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        *reinterpret_cast<uint4*>(p) = r;
    }
};

// 8 consecutive elements kept in their storage format (one 16-byte register quad for bf16): row kernels hold many of these in
// flight per lane and unpack to fp32 only when they compute, which halves the registers of the prefetched data
template <typename T> struct Raw8;
template <> struct Raw8<bf16> {
    uint4 r;
    __device__ __forceinline__ void load(const bf16* p) { r = *reinterpret_cast<const uint4*>(p); }
    // read-once streams: do not allocate in L1
    __device__ __forceinline__ void load_stream(const bf16* p) {
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    }
    __device__ __forceinline__ void zero() { r = make_uint4(0u, 0u, 0u, 0u); }
    __device__ __forceinline__ void unpack(float (&v)[8]) const {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
        for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
    }
};
template <> struct Raw8<float> {
    float4 a, b;
    __device__ __forceinline__ void load(const float* p) { a = *reinterpret_cast<const float4*>(p); b = *reinterpret_cast<const float4*>(p + 4); }
    __device__ __forceinline__ void load_stream(const float* p) {
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(p));
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p + 4));
    }
    __device__ __forceinline__ void zero() { a = make_float4(0.f, 0.f, 0.f, 0.f); b = a; }
    __device__ __forceinline__ void unpack(float (&v)[8]) const {
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
};

template <typename T> __device__ __forceinline__ float ld1(const T* p);
template <> __device__ __forceinline__ float ld1<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld1<bf16>(const bf16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void st1(T* p, float v);
template <> __device__ __forceinline__ void st1<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st1<bf16>(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// ---- dropout (training only) -----------------------------------------------------------------------------
// nn.Dropout(p) of the reference (video_nmn/modules.py: every Linear->ReLU->Dropout site, HasItem's Sigmoid->Dropout, decoder)
// as a counter-based mask: element (row, col) of dropout site `site` is kept iff hash(seed, site, row, col) >= p * 2^32 and kept
// values are scaled by 1/(1-p).  `row` is a global row id (sorted node position * rows-per-instance + r), so the mask does not
// depend on chunking and the backward pass (which re-runs the forward of a chunk) regenerates exactly the same mask.  torch's
// Philox stream cannot be reproduced batched (the reference draws per question, in interpreter order); parity is checked
// against the oracle with the same masks injected (oracle/nmn_oracle.py dropout_keep).
__host__ __device__ static inline uint32_t lowbias32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
struct DropSpec {
    uint32_t thresh = 0;      // p * 2^32 (0 = no dropout)
    uint32_t key_lo = 0, key_hi = 0;
    float scale = 1.0f;       // 1 / (1 - p)
    long long row0 = 0;       // global row id of local row 0
};
__host__ __device__ static inline uint32_t drop_row_hash(uint32_t key_lo, uint32_t key_hi, long long row) {
    return lowbias32(static_cast<uint32_t>(row) ^ key_lo) + key_hi;
}
__host__ __device__ static inline bool drop_keep(uint32_t row_hash, uint32_t col, uint32_t thresh) {
    return lowbias32(row_hash + col * 0x9E3779B1u) >= thresh;
}
static inline DropSpec make_drop(float p, unsigned long long seed, int site, long long row0) {
    DropSpec d;
    if (!(p > 0.0f)) return d;
    const double t = static_cast<double>(p) * 4294967296.0;
    d.thresh = t >= 4294967295.0 ? 0xffffffffu : static_cast<uint32_t>(t);
    d.key_lo = lowbias32(static_cast<uint32_t>(seed) ^ (static_cast<uint32_t>(site) * 0x9E3779B1u));
    d.key_hi = lowbias32(static_cast<uint32_t>(seed >> 32) + static_cast<uint32_t>(site) * 0x85EBCA77u + 1u);
    d.scale = 1.0f / (1.0f - p);
    d.row0 = row0;
    return d;
}

// ---- internal GEMM launch API (gemm_sm100.cu) ----------------------------------------------------
struct GemmArgs {
    const void* A = nullptr;        // bf16 [nplanes][a_plane_rows, lda]  (or the slot arena in gather mode)
    long long lda = 0;
    int a_plane_rows = 0;
    const int* a_slots = nullptr;   // gather mode: row m comes from slot a_slots[m / slot_rows], row m % slot_rows
    int slot_rows = 0;
    long long arena_slots = 0;
    const void* W = nullptr;        // bf16 [nplanes][w_plane_rows, ldw]
    long long ldw = 0;
    int w_plane_rows = 0;
    int nplanes = 1;
    const float* bias = nullptr;
    const float* row_scale = nullptr;
    void* C = nullptr;
    long long ldc = 0;
    int out_dtype = STAIR_BF16;
    int M = 0, N = 0, K = 0;
    int act = STAIR_ACT_NONE;
    int accumulate = 0;
    // mn_major = 1: both operands are stored "transposed": A = [nplanes][a_plane_rows (>= K), lda] with element (k, m) at k*lda + m,
    // W = [nplanes][w_plane_rows (>= K), ldw] with element (k, n) at k*ldw + n, i.e. C[M,N] = A^T . W — the weight-gradient
    // contraction dW = dZ^T . X over the rows of a layer, read in place (MN-major UMMA operands, no transposed copies).
    int mn_major = 0;
    int atomic_acc = 0;             // accumulate with atomics even without split-K (concurrent launches adding into the same C)
    DropSpec drop;                  // applied after the activation (training forward only)
    // frame-sum epilogue (Filter aggregation, modules.py:374-376, folded into the producing GEMM): instead of C, the kernel writes
    // sum_out[i][n] = sum_t bf16(act(...))[i * sum_T + t][n] (bf16 [M / sum_T, ld_sum]); sum_T divides 128 or is a multiple of 32 that divides 128
    void* sum_out = nullptr;
    long long ld_sum = 0;
    int sum_T = 0;
};
bool gemm_sum_epilogue_ok(int T);
int launch_gemm(const GemmArgs& a, cudaStream_t st);
int* err_flag_ptr();      // pinned host word that device-side protocol timeouts write their code to (stair_gemm_error_flag)
void err_flag_free();
// true when the TMA slot-gather path can serve this (slot_rows, K) without a staging copy
static inline bool gemm_gather_ok(int slot_rows) { return slot_rows >= 8 && slot_rows <= 128 && (128 % slot_rows) == 0; }

}  // namespace stair
