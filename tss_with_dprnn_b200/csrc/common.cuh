// Shared helpers for libdprnn_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>
#include <cstring>

#define DPRNN_STR2(x) #x
#define DPRNN_STR(x) DPRNN_STR2(x)

namespace dprnn {

// last-error slot shared by every translation unit of the library (defined in api.cu)
void set_error(const char* fmt, ...);

#define DPRNN_CHECK_ARG(cond)                                                        \
    do {                                                                             \
        if (!(cond)) {                                                               \
            ::dprnn::set_error("%s:%d: bad argument: %s", __FILE__, __LINE__, #cond); \
            return 2;                                                                \
        }                                                                            \
    } while (0)

#define DPRNN_CHECK_LAUNCH()                                                          \
    do {                                                                             \
        cudaError_t e__ = cudaGetLastError();                                        \
        if (e__ != cudaSuccess) {                                                    \
            ::dprnn::set_error("%s:%d: launch failed: %s", __FILE__, __LINE__,        \
                               cudaGetErrorString(e__));                             \
            return 1;                                                                \
        }                                                                            \
    } while (0)

#define DPRNN_CUDA(call)                                                              \
    do {                                                                             \
        cudaError_t e__ = (call);                                                    \
        if (e__ != cudaSuccess) {                                                    \
            ::dprnn::set_error("%s:%d: %s: %s", __FILE__, __LINE__, #call,            \
                               cudaGetErrorString(e__));                             \
            return 1;                                                                \
        }                                                                            \
    } while (0)

// 16-bit storage / tensor-core operand formats (the `h16` argument of the C ABI, DPRNN_H16_* in the header): bf16 has
// fp32's range and 8 significand bits, fp16 11 significand bits - the format of the 1e-3 tolerance mode
constexpr int H16_BF16 = 0, H16_FP16 = 1;
template <bool kF16>
__device__ __forceinline__ uint32_t pack_h16x2(float a, float b) {
    if constexpr (kF16) {
        // fp16 has 5 exponent bits: a value beyond +-65504 saturates instead of becoming inf (one F2FP either way), so an
        // outlier activation degrades one element instead of poisoning the recurrence with inf - inf
        uint32_t v;
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(v) : "f"(b), "f"(a));
        return v;
    } else {
        __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&v);
    }
}
template <bool kF16>
__device__ __forceinline__ float2 unpack_h16x2(uint32_t v) {
    if constexpr (kF16) return __half22float2(*reinterpret_cast<const __half2*>(&v));
    else return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v));
}

// x + norm(y) of one element of a half-block tail (dprnn.py:90-92,98-99).  ONE definition shared by the stand-alone norm
// kernels and the LSTM kernel that applies the pending norm while it loads its input (lstm_tc_pp.cu, kFuse), so that both
// round identically.
__device__ __forceinline__ float norm_res1(float x, float y, float mean, float rstd, float g, float be) {
    return x + ((y - mean) * rstd * g + be);
}

static inline unsigned cdiv(long a, long b) { return (unsigned)((a + b - 1) / b); }

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum of a double for blocks of up to 1024 threads. All threads get the result.
__device__ __forceinline__ double block_sum(double v, double* scratch /*[32]*/) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    double r = (lane < nw) ? scratch[lane] : 0.0;
    r = warp_sum(r);
    return r;
}

__device__ __forceinline__ float4 ld_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

// The two decoder taps of one frame for kernel 2 / stride 1 / N = 64: p_j = sum_c mask[c]*enc[c]*w[c,j], in a fixed
// summation order (shared by the uniform and the ragged decoder so that packing a batch changes no bit).
__device__ __forceinline__ void decode_taps64(const float* __restrict__ m, const float* __restrict__ e,
                                              const float* __restrict__ sw, float& p0, float& p1) {
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int c4 = 0; c4 < 16; ++c4) {
        const float4 mv = ld_stream(reinterpret_cast<const float4*>(m) + c4);
        const float4 ev = ld_stream(reinterpret_cast<const float4*>(e) + c4);
        const float z0 = mv.x * ev.x, z1 = mv.y * ev.y, z2 = mv.z * ev.z, z3 = mv.w * ev.w;
        a0 = fmaf(z0, sw[8 * c4 + 0], a0); a1 = fmaf(z0, sw[8 * c4 + 1], a1);
        a0 = fmaf(z1, sw[8 * c4 + 2], a0); a1 = fmaf(z1, sw[8 * c4 + 3], a1);
        a0 = fmaf(z2, sw[8 * c4 + 4], a0); a1 = fmaf(z2, sw[8 * c4 + 5], a1);
        a0 = fmaf(z3, sw[8 * c4 + 6], a0); a1 = fmaf(z3, sw[8 * c4 + 7], a1);
    }
    p0 = a0; p1 = a1;
}

}  // namespace dprnn
