// MUFU throughput micro-benchmark (B200): ops / clk / SM for tanh.approx.f32, ex2.approx.f32, rcp.approx.f32, tanh.approx.f16x2
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(float* out, int iters, long long* clk) {
    float x[8];
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 0.001f + i * 0.1f;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[i]));
            if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
            if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
            if (OP == 3) { unsigned u = __float_as_uint(x[i]); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(u)); x[i] = __uint_as_float(u); }
            if (OP == 4) { unsigned u = __float_as_uint(x[i]); asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(u)); x[i] = __uint_as_float(u); }
            if (OP == 5) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(x[i]));
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
int main() {
    float* out; long long* clk; cudaMalloc(&out, 148 * 1024 * 4); cudaMallocManaged(&clk, 8);
    const char* names[] = {"tanh.approx.f32", "ex2.approx.f32", "rcp.approx.f32", "tanh.approx.f16x2", "tanh.approx.bf16x2", "fma.f32"};
    for (int threads : {256, 512, 1024}) {
        for (int op = 0; op < 6; ++op) {
            const int iters = 2000;
            for (int rep = 0; rep < 2; ++rep) {
                if (op == 0) k<0><<<148, threads>>>(out, iters, clk);
                if (op == 1) k<1><<<148, threads>>>(out, iters, clk);
                if (op == 2) k<2><<<148, threads>>>(out, iters, clk);
                if (op == 3) k<3><<<148, threads>>>(out, iters, clk);
                if (op == 4) k<4><<<148, threads>>>(out, iters, clk);
                if (op == 5) k<5><<<148, threads>>>(out, iters, clk);
                cudaDeviceSynchronize();
            }
            printf("%-20s threads/SM %4d: %.2f instr-lanes / clk / SM\n", names[op], threads, (double)iters * 8 * threads / (double)*clk);
        }
    }
    return 0;
}
