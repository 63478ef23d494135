// Exact-fp32 LSTM recurrence (CUDA cores).  One CTA owns 32 sequences of one direction for every
// time step: h_{t-1} stays in shared memory, the cell state in registers, and W_hh^T (256 KiB in
// fp32 - larger than one SM's shared memory) is streamed from L2 through a cp.async double buffer
// in 16-row slabs that wrap around seamlessly from one step to the next.
//
// Thread map (256 threads): warp w -> sequences 4w..4w+3, lane -> hidden units 4*lane..4*lane+3 of
// all four gates, i.e. gate columns {g*128 + 4*lane + u}: a thread owns complete (i,f,g,o) quadruples
// so the cell update is thread-local and PyTorch's native [i;f;g;o] row order needs no permutation.
#include "common.cuh"
#include "../../include/dprnn_b200.h"

namespace dprnn {

constexpr int H = 128, G4 = 4 * H, SEQ_TILE = 32, KC = 16, NCHUNK = H / KC;

struct LstmParams {
    const float* gx;     // [rows, ndir*4H]  x W_ih^T + b_ih + b_hh
    const float* whhT;   // [ndir][H][4H]
    float* hout;         // [rows, ndir*H]
    long nseq; int T;
    long seq_div, seq_outer, seq_inner, step_stride;
    int ndir;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
__device__ __forceinline__ void cp_async_wait0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__global__ void __launch_bounds__(256, 2) lstm_simt_kernel(const LstmParams p) {
    extern __shared__ __align__(16) float smem[];
    float* hs = smem;                       // [H][SEQ_TILE], column group rotated by (k>>2) (see below)
    float* ws = smem + H * SEQ_TILE;        // [2][KC][G4]

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int dir = blockIdx.y;
    const long seq0 = (long)blockIdx.x * SEQ_TILE + w * 4;
    const float* whh = p.whhT + (long)dir * H * G4;
    const int ldg = p.ndir * G4, ldh = p.ndir * H;

    long base[4];
    bool ok[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const long n = seq0 + s;
        ok[s] = n < p.nseq;
        base[s] = ok[s] ? (n / p.seq_div) * p.seq_outer + (n % p.seq_div) * p.seq_inner : 0;
    }

    for (int i = tid; i < H * SEQ_TILE; i += 256) hs[i] = 0.f;
    float c[4][4];
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int u = 0; u < 4; ++u) c[s][u] = 0.f;

    auto issue_chunk = [&](int chunk, int buf) {   // 16 x 512 floats = 2048 float4, 8 per thread
        const float* src = whh + (long)chunk * KC * G4;
        float* dst = ws + buf * KC * G4;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int idx = (tid + i * 256) * 4;
            cp_async16(dst + idx, src + idx);
        }
        cp_async_commit();
    };

    issue_chunk(0, 0);
    int gchunk = 0;   // running slab counter; slab = gchunk % NCHUNK, buffer = gchunk & 1

    for (int step = 0; step < p.T; ++step) {
        const int t = dir ? p.T - 1 - step : step;
        // accumulators start from the input projection
        float acc[4][4][4];   // [seq][gate][unit]
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const float* g = p.gx + (base[s] + (long)t * p.step_stride) * ldg + dir * G4 + lane * 4;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok[s]) v = ld_stream(reinterpret_cast<const float4*>(g + q * H));
                acc[s][q][0] = v.x; acc[s][q][1] = v.y; acc[s][q][2] = v.z; acc[s][q][3] = v.w;
            }
        }
        if (step + 1 < p.T) {   // pull next step's projection rows towards L2
            const int tn = dir ? t - 1 : t + 1;
#pragma unroll
            for (int s = 0; s < 4; ++s)
                if (ok[s]) {
                    const float* g = p.gx + (base[s] + (long)tn * p.step_stride) * ldg + dir * G4 + lane * 16;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(g));
                }
        }

        for (int kc = 0; kc < NCHUNK; ++kc, ++gchunk) {
            cp_async_wait0();
            __syncthreads();   // slab visible; everyone is done with the other buffer (and, at kc==0, h is written)
            const bool more = (step + 1 < p.T) || (kc + 1 < NCHUNK);
            if (more) issue_chunk((gchunk + 1) % NCHUNK, (gchunk + 1) & 1);
            const float* wb = ws + (gchunk & 1) * KC * G4 + lane * 4;
#pragma unroll
            for (int kk = 0; kk < KC; ++kk) {
                const int k = kc * KC + kk;
                const float4 hv = *reinterpret_cast<const float4*>(hs + k * SEQ_TILE + 4 * ((w + (k >> 2)) & 7));
                const float hq[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 wv = *reinterpret_cast<const float4*>(wb + kk * G4 + q * H);
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        acc[s][q][0] = fmaf(hq[s], wv.x, acc[s][q][0]);
                        acc[s][q][1] = fmaf(hq[s], wv.y, acc[s][q][1]);
                        acc[s][q][2] = fmaf(hq[s], wv.z, acc[s][q][2]);
                        acc[s][q][3] = fmaf(hq[s], wv.w, acc[s][q][3]);
                    }
                }
            }
        }
        __syncthreads();   // all reads of h_{t-1} finished before it is overwritten

        float hn[4][4];
#pragma unroll
        for (int s = 0; s < 4; ++s) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float ig = sigmoid_acc(acc[s][0][u]);
                const float fg = sigmoid_acc(acc[s][1][u]);
                const float gg = tanhf(acc[s][2][u]);
                const float og = sigmoid_acc(acc[s][3][u]);
                c[s][u] = fg * c[s][u] + ig * gg;
                hn[s][u] = og * tanhf(c[s][u]);
            }
            if (ok[s]) {
                float* o = p.hout + (base[s] + (long)t * p.step_stride) * ldh + dir * H + lane * 4;
                *reinterpret_cast<float4*>(o) = make_float4(hn[s][0], hn[s][1], hn[s][2], hn[s][3]);
            }
        }
        // h_t into shared memory, transposed: row k = unit, 4 sequences of this warp as one float4.
        // Rotating the column group by (k>>2) = lane spreads a warp's 32 stores over all banks.
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = lane * 4 + u;
            *reinterpret_cast<float4*>(hs + k * SEQ_TILE + 4 * ((w + lane) & 7)) =
                make_float4(hn[0][u], hn[1][u], hn[2][u], hn[3][u]);
        }
        // the __syncthreads at the top of the next step's first slab orders these writes before the reads
    }
}

}  // namespace dprnn

using namespace dprnn;

extern "C" int dprnn_lstm_recurrence_f32(const float* gx, const float* whhT, float* hout, long nseq, int T,
                                         long seq_div, long seq_outer_stride, long seq_inner_stride,
                                         long step_stride, int hidden, int ndir, void* stream) {
    DPRNN_CHECK_ARG(gx && whhT && hout && nseq > 0 && T > 0 && seq_div > 0);
    DPRNN_CHECK_ARG(hidden == H && (ndir == 1 || ndir == 2));
    DPRNN_CHECK_ARG(((uintptr_t)gx | (uintptr_t)whhT | (uintptr_t)hout) % 16 == 0);
    const size_t smem = (size_t)(H * SEQ_TILE + 2 * KC * G4) * sizeof(float);
    DPRNN_CUDA(cudaFuncSetAttribute(lstm_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LstmParams p{gx, whhT, hout, nseq, T, seq_div, seq_outer_stride, seq_inner_stride, step_stride, ndir};
    dim3 grid(cdiv(nseq, SEQ_TILE), ndir);
    lstm_simt_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(p);
    DPRNN_CHECK_LAUNCH();
    return 0;
}
