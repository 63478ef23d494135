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
    float* gates_out;    // training: [rows, ndir*4H] gate ACTIVATIONS i,f,g,o of every step (for BPTT), or NULL
    float* c_out;        // training: [rows, ndir*H] cell state after every step, or NULL
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
                acc[s][0][u] = ig; acc[s][1][u] = fg; acc[s][2][u] = gg; acc[s][3][u] = og;
            }
            if (ok[s] && p.gates_out) {
                const long rowi = base[s] + (long)t * p.step_stride;
                float* go = p.gates_out + rowi * ldg + dir * G4 + lane * 4;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    *reinterpret_cast<float4*>(go + q * H) = make_float4(acc[s][q][0], acc[s][q][1], acc[s][q][2], acc[s][q][3]);
                *reinterpret_cast<float4*>(p.c_out + rowi * ldh + dir * H + lane * 4) = make_float4(c[s][0], c[s][1], c[s][2], c[s][3]);
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
    LstmParams p{gx, whhT, hout, nseq, T, seq_div, seq_outer_stride, seq_inner_stride, step_stride, ndir, nullptr, nullptr};
    dim3 grid(cdiv(nseq, SEQ_TILE), ndir);
    lstm_simt_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(p);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

// Training forward: as above, additionally storing the gate activations and the cell state of every step for BPTT.
extern "C" int dprnn_lstm_recurrence_f32_train(const float* gx, const float* whhT, float* hout, float* gates,
                                               float* cstate, long nseq, int T, long seq_div, long seq_outer_stride,
                                               long seq_inner_stride, long step_stride, int hidden, int ndir,
                                               void* stream) {
    DPRNN_CHECK_ARG(gx && whhT && hout && gates && cstate && nseq > 0 && T > 0 && seq_div > 0);
    DPRNN_CHECK_ARG(hidden == H && (ndir == 1 || ndir == 2));
    DPRNN_CHECK_ARG(((uintptr_t)gx | (uintptr_t)whhT | (uintptr_t)hout | (uintptr_t)gates | (uintptr_t)cstate) % 16 == 0);
    const size_t smem = (size_t)(H * SEQ_TILE + 2 * KC * G4) * sizeof(float);
    DPRNN_CUDA(cudaFuncSetAttribute(lstm_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LstmParams p{gx, whhT, hout, nseq, T, seq_div, seq_outer_stride, seq_inner_stride, step_stride, ndir, gates, cstate};
    dim3 grid(cdiv(nseq, SEQ_TILE), ndir);
    lstm_simt_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(p);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// BPTT: given d h_t for every step (gradient of the layer output), the saved gate activations and cell states,
// produce the gradient of the gate PRE-activations of every step.  Same tiling as the forward: one CTA owns 32
// sequences of one direction, walks the steps in reverse direction order, keeps d c and the recurrent d h in
// registers (a thread owns 4 sequences x 4 hidden units), and forms d h_{t-1} = d gates_t @ W_hh with the 512
// gate gradients of the step staged in shared memory and W_hh ([4H, H], PyTorch layout) streamed from L2.
// dx, dW_ih, dW_hh and the bias gradient follow from d gates as time-parallel contractions (backward.cu).
// ---------------------------------------------------------------------------------------------------------
namespace dprnn {

struct LstmBwdParams {
    const float* dh_out;   // [rows, ndir*H]
    const float* gates;    // [rows, ndir*4H] activations i,f,g,o
    const float* cstate;   // [rows, ndir*H]
    const float* whh;      // [ndir][4H][H]
    float* dgates;         // [rows, ndir*4H] gradient of the gate pre-activations
    long nseq; int T;
    long seq_div, seq_outer, seq_inner, step_stride;
    int ndir;
};

constexpr int GC = 16;      // gate rows of W_hh per slab

__global__ void __launch_bounds__(256, 2) lstm_bwd_kernel(const LstmBwdParams p) {
    extern __shared__ __align__(16) float smem[];
    float* dgs = smem;                      // [G4][SEQ_TILE], column group rotated like hs in the forward
    float* ws = smem + G4 * SEQ_TILE;       // [2][GC][H]

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int dir = blockIdx.y;
    const long seq0 = (long)blockIdx.x * SEQ_TILE + w * 4;
    const float* whh = p.whh + (long)dir * G4 * H;
    const int ldg = p.ndir * G4, ldh = p.ndir * H;
    long base[4];
    bool ok[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const long n = seq0 + s;
        ok[s] = n < p.nseq;
        base[s] = ok[s] ? (n / p.seq_div) * p.seq_outer + (n % p.seq_div) * p.seq_inner : 0;
    }
    float dc[4][4], dhr[4][4];
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int u = 0; u < 4; ++u) { dc[s][u] = 0.f; dhr[s][u] = 0.f; }

    auto issue_slab = [&](int slab, int buf) {     // 16 x 128 floats = 512 float4, 2 per thread
        const float* src = whh + (long)slab * GC * H;
        float* dst = ws + buf * GC * H;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int idx = (tid + i * 256) * 4;
            cp_async16(dst + idx, src + idx);
        }
        cp_async_commit();
    };
    constexpr int NSLAB = G4 / GC;
    issue_slab(0, 0);
    int gslab = 0;

    for (int step = 0; step < p.T; ++step) {
        // the forward walked t = 0..T-1 (dir 0) or T-1..0 (dir 1); BPTT walks it backwards
        const int fstep = p.T - 1 - step;                       // forward step index being differentiated
        const int t = dir ? p.T - 1 - fstep : fstep;
        const int tprev = dir ? t + 1 : t - 1;                  // time index of the forward's previous step
        __syncthreads();                                        // previous step's reads of dgs are complete
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const long rowi = base[s] + (long)t * p.step_stride;
            float4 gi = make_float4(0.f, 0.f, 0.f, 0.f), gf = gi, gg = gi, go = gi, cv = gi, cp = gi, dho = gi;
            if (ok[s]) {
                const float* g = p.gates + rowi * ldg + dir * G4 + lane * 4;
                gi = *reinterpret_cast<const float4*>(g); gf = *reinterpret_cast<const float4*>(g + H);
                gg = *reinterpret_cast<const float4*>(g + 2 * H); go = *reinterpret_cast<const float4*>(g + 3 * H);
                cv = *reinterpret_cast<const float4*>(p.cstate + rowi * ldh + dir * H + lane * 4);
                if (fstep > 0) cp = *reinterpret_cast<const float4*>(p.cstate + (base[s] + (long)tprev * p.step_stride) * ldh + dir * H + lane * 4);
                dho = *reinterpret_cast<const float4*>(p.dh_out + rowi * ldh + dir * H + lane * 4);
            }
            const float ia[4] = {gi.x, gi.y, gi.z, gi.w}, fa[4] = {gf.x, gf.y, gf.z, gf.w}, ga[4] = {gg.x, gg.y, gg.z, gg.w},
                        oa[4] = {go.x, go.y, go.z, go.w}, ca[4] = {cv.x, cv.y, cv.z, cv.w}, pa[4] = {cp.x, cp.y, cp.z, cp.w},
                        da[4] = {dho.x, dho.y, dho.z, dho.w};
            float dpi[4], dpf[4], dpg[4], dpo[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float dh = da[u] + dhr[s][u];
                const float tc = tanhf(ca[u]);
                const float dct = dc[s][u] + dh * oa[u] * (1.f - tc * tc);
                dpo[u] = dh * tc * oa[u] * (1.f - oa[u]);
                dpi[u] = dct * ga[u] * ia[u] * (1.f - ia[u]);
                dpf[u] = dct * pa[u] * fa[u] * (1.f - fa[u]);
                dpg[u] = dct * ia[u] * (1.f - ga[u] * ga[u]);
                dc[s][u] = dct * fa[u];
            }
            if (ok[s]) {
                float* o = p.dgates + rowi * ldg + dir * G4 + lane * 4;
                *reinterpret_cast<float4*>(o) = make_float4(dpi[0], dpi[1], dpi[2], dpi[3]);
                *reinterpret_cast<float4*>(o + H) = make_float4(dpf[0], dpf[1], dpf[2], dpf[3]);
                *reinterpret_cast<float4*>(o + 2 * H) = make_float4(dpg[0], dpg[1], dpg[2], dpg[3]);
                *reinterpret_cast<float4*>(o + 3 * H) = make_float4(dpo[0], dpo[1], dpo[2], dpo[3]);
            }
            // stage into shared memory as dgs[gate row][seq]; gate row g = q*H + 4*lane + u
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int k = lane * 4 + u;
                const int col = 4 * ((w + ((k >> 2) & 7)) & 7) + s;
                dgs[(0 * H + k) * SEQ_TILE + col] = dpi[u];
                dgs[(1 * H + k) * SEQ_TILE + col] = dpf[u];
                dgs[(2 * H + k) * SEQ_TILE + col] = dpg[u];
                dgs[(3 * H + k) * SEQ_TILE + col] = dpo[u];
            }
        }
        if (fstep == 0) break;                                   // no earlier step to propagate to
        float acc[4][4];
#pragma unroll
        for (int s = 0; s < 4; ++s)
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[s][u] = 0.f;
        for (int sl = 0; sl < NSLAB; ++sl, ++gslab) {
            cp_async_wait0();
            __syncthreads();                                     // slab visible; dgs written (first slab)
            issue_slab((gslab + 1) % NSLAB, (gslab + 1) & 1);
            const float* wb = ws + (gslab & 1) * GC * H + lane * 4;
#pragma unroll
            for (int gg2 = 0; gg2 < GC; ++gg2) {
                const int g = sl * GC + gg2;
                const int k = g & (H - 1);
                const float4 dv = *reinterpret_cast<const float4*>(dgs + g * SEQ_TILE + 4 * ((w + ((k >> 2) & 7)) & 7));
                const float4 wv = *reinterpret_cast<const float4*>(wb + gg2 * H);
                const float dq[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    acc[s][0] = fmaf(dq[s], wv.x, acc[s][0]);
                    acc[s][1] = fmaf(dq[s], wv.y, acc[s][1]);
                    acc[s][2] = fmaf(dq[s], wv.z, acc[s][2]);
                    acc[s][3] = fmaf(dq[s], wv.w, acc[s][3]);
                }
            }
        }
#pragma unroll
        for (int s = 0; s < 4; ++s)
#pragma unroll
            for (int u = 0; u < 4; ++u) dhr[s][u] = acc[s][u];
    }
    cp_async_wait0();
}

}  // namespace dprnn

extern "C" int dprnn_lstm_bptt_f32(const float* dh_out, const float* gates, const float* cstate, const float* whh,
                                   float* dgates, long nseq, int T, long seq_div, long seq_outer_stride,
                                   long seq_inner_stride, long step_stride, int hidden, int ndir, void* stream) {
    DPRNN_CHECK_ARG(dh_out && gates && cstate && whh && dgates && nseq > 0 && T > 0 && seq_div > 0);
    DPRNN_CHECK_ARG(hidden == H && (ndir == 1 || ndir == 2));
    DPRNN_CHECK_ARG(((uintptr_t)dh_out | (uintptr_t)gates | (uintptr_t)cstate | (uintptr_t)whh | (uintptr_t)dgates) % 16 == 0);
    const size_t smem = (size_t)(G4 * SEQ_TILE + 2 * GC * H) * sizeof(float);
    DPRNN_CUDA(cudaFuncSetAttribute(lstm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LstmBwdParams p{dh_out, gates, cstate, whh, dgates, nseq, T, seq_div, seq_outer_stride, seq_inner_stride, step_stride, ndir};
    dim3 grid(cdiv(nseq, SEQ_TILE), ndir);
    lstm_bwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(p);
    DPRNN_CHECK_LAUNCH();
    return 0;
}
