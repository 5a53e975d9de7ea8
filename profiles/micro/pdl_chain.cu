// How much does programmatic dependent launch (griddepcontrol) shorten a chain of small dependent kernels on one stream?
// A chain of N kernels, each reading the previous one's output (n rows of 512 bf16, like a module-phase elementwise kernel at 818 x 8 rows),
// launched (a) plainly, (b) with cudaLaunchAttributeProgrammaticStreamSerialization and griddepcontrol.launch_dependents at the top /
// griddepcontrol.wait before the first global read.  Prints us per kernel.   nvcc -arch=sm_100a -O3 -o pdl_chain pdl_chain.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

template <bool PDL>
__global__ void __launch_bounds__(256) scale_rows(const uint4* __restrict__ in, uint4* __restrict__ out, long long n16, float s) {
    __shared__ float dummy[32];
    if (threadIdx.x < 32) dummy[threadIdx.x] = s;                 // a little prologue work that does not depend on the producer
    __syncthreads();
    if (PDL) {
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    const float sc = dummy[threadIdx.x & 31];
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n16; i += gridDim.x * 256LL) {
        uint4 v = in[i];
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
        for (int k = 0; k < 4; ++k) { float2 f = __bfloat1622float2(h[k]); f.x *= sc; f.y *= sc; h[k] = __floats2bfloat162_rn(f.x, f.y); }
        out[i] = v;
    }
}

int main() {
    const long long rows = 818 * 8, H = 512, n16 = rows * H * 2 / 16;
    uint4 *a, *b;
    cudaMalloc(&a, n16 * 16); cudaMalloc(&b, n16 * 16);
    cudaMemset(a, 0, n16 * 16); cudaMemset(b, 0, n16 * 16);
    cudaStream_t st; cudaStreamCreate(&st);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int N = 40, grid = 148 * 4;
    for (int mode = 0; mode < 3; ++mode) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0, st);
            for (int i = 0; i < N; ++i) {
                uint4* src = (i & 1) ? b : a; uint4* dst = (i & 1) ? a : b;
                if (mode == 0) scale_rows<false><<<grid, 256, 0, st>>>(src, dst, n16, 1.0f);
                else {
                    cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.stream = st;
                    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                    at[0].val.programmaticStreamSerializationAllowed = 1;
                    cfg.attrs = at; cfg.numAttrs = 1;
                    if (mode == 1) cudaLaunchKernelEx(&cfg, scale_rows<true>, (const uint4*)src, dst, n16, 1.0f);
                    else cudaLaunchKernelEx(&cfg, scale_rows<false>, (const uint4*)src, dst, n16, 1.0f);      // attribute only, no griddepcontrol (implicit trigger at exit)
                }
            }
            cudaEventRecord(e1, st); cudaStreamSynchronize(st);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep == 2) printf("%s: %.2f us per kernel (chain of %d, %lld rows x 512 bf16, err %d)\n",
                                 mode == 0 ? "plain launches" : (mode == 1 ? "PDL + griddepcontrol" : "PDL attribute only"), 1e3 * ms / N, N, rows, (int)cudaGetLastError());
        }
    }
    return 0;
}
