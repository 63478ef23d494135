// Microbenchmark: MUFU throughput of the activation forms considered for the LSTM epilogue (B200).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mufu_bench tools/mufu_bench.cu && /tmp/mufu_bench
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdint>

template <int MODE>
__global__ void k(float* out, int iters) {
    float a[8];
    uint32_t h[8];
    for (int i = 0; i < 8; ++i) { a[i] = 0.001f * (threadIdx.x + i); h[i] = 0x3c003800u + threadIdx.x + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
            if (MODE == 1) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h[i]));
            if (MODE == 2) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(h[i]));
            if (MODE == 3) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (MODE == 4) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
            if (MODE == 5) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        }
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int vals_per_op) {
    float* out; cudaMalloc(&out, 148 * 4 * 256 * 4);
    int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 4, 256>>>(out, 100);
    cudaEventRecord(e0);
    k<MODE><<<148 * 4, 256>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = 148.0 * 4 * 256 * 8.0 * iters;          // thread-level instructions
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-22s %.3f ms  %.2f lane-ops/clk/SM (at %d MHz nominal)  -> %.2f results/clk/SM\n", name, ms,
           ops / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000, vals_per_op * ops / (ms * 1e-3) / 148 / (clk * 1e3));
    cudaFree(out);
}

int main() {
    run<0>("tanh.approx.f32", 1);
    run<1>("tanh.approx.f16x2", 2);
    run<2>("tanh.approx.bf16x2", 2);
    run<3>("ex2.approx.ftz.f32", 1);
    run<4>("ex2.approx.f16x2", 2);
    run<5>("rcp.approx.ftz.f32", 1);
    return 0;
}
