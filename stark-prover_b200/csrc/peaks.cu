// peaks.cu — measures the integer-pipe issue peak on the device the context lives on.
// MEASURED_PEAKS.json carries HBM GB/s and bf16 TFLOP/s only; SHA-256 is bound by the 32-bit integer pipes
// (SHF / LOP3 / IADD3 on the ALU pipe, IMAD on the FMA pipe), so the denominator of its roofline is
// measured here with register-only dependency chains (8 independent chains per thread).
#include "../../include/stark_b200.h"
#include "handles.hpp"

namespace starkb200 {

template <bool WITH_IMAD>
__global__ void __launch_bounds__(256) int_peak_kernel(uint32_t* sink, int iters, uint32_t seed, uint32_t mulc) {
    uint32_t x[8], y = seed ^ threadIdx.x, z = seed + blockIdx.x;
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = seed * (i + 1) + threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                uint32_t t;
                asm volatile("shf.r.wrap.b32 %0, %1, %1, 7;" : "=r"(t) : "r"(x[i]));        // SHF
                asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(t) : "r"(t), "r"(y), "r"(z));   // LOP3 (xor3)
                asm volatile("{ .reg .u32 q; add.u32 q, %1, %2; add.u32 %0, q, %3; }" : "=r"(t) : "r"(t), "r"(y), "r"(z));  // IADD3
                if (WITH_IMAD) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(t) : "r"(t), "r"(mulc), "r"(z));       // IMAD
                x[i] = t;
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) acc ^= x[i];
    if (acc == 0x12345678u) sink[0] = acc;     // practically never; keeps the chains alive
}

// Diagnostic mixes for pipe-balancing decisions (tools/exp_pipe_mix.py).  Every step issues the three ALU-pipe
// instructions above plus, per MODE: 1 = IMAD; 2 = a rotate emulated on the FMA pipe (mul.wide.u32 by 2^k, then
// lo + hi as mad.lo); 3 = mul.hi.u32 (IMAD.HI, a right shift); 4 = two IMADs.
template <int MODE>
__global__ void __launch_bounds__(256) pipe_mix_kernel(uint32_t* sink, int iters, uint32_t seed, uint32_t mulc) {
    uint32_t x[8], y = seed ^ threadIdx.x, z = seed + blockIdx.x;
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = seed * (i + 1) + threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                uint32_t t = x[i];
                if (MODE < 5) {
                    asm volatile("shf.r.wrap.b32 %0, %1, %1, 7;" : "=r"(t) : "r"(t));
                    asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(t) : "r"(t), "r"(y), "r"(z));
                    asm volatile("{ .reg .u32 q; add.u32 q, %1, %2; add.u32 %0, q, %3; }" : "=r"(t) : "r"(t), "r"(y), "r"(z));
                }
                if (MODE == 5) {                                   // FMA pipe alone: 3 IMAD
#pragma unroll
                    for (int k = 0; k < 3; k++) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(t) : "r"(t), "r"(mulc), "r"(z));
                }
                if (MODE == 6) {                                   // 3 IMAD.HI
#pragma unroll
                    for (int k = 0; k < 3; k++) asm volatile("mad.hi.u32 %0, %1, %2, %3;" : "=r"(t) : "r"(t), "r"(y), "r"(z));
                }
                if (MODE == 7) {                                   // 3 IMAD.WIDE (the high word feeds the next one)
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        uint32_t lo;
                        asm volatile("{ .reg .u64 w; mul.wide.u32 w, %2, %3; mov.b64 {%0, %1}, w; }" : "=r"(lo), "=r"(t) : "r"(t), "r"(y));
                        z ^= lo & 0u;
                    }
                }
                if (MODE == 1 || MODE == 4) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(t) : "r"(t), "r"(mulc), "r"(z));
                if (MODE == 4) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(t) : "r"(t), "r"(mulc), "r"(y));
                if (MODE == 2) {
                    uint32_t lo, hi;
                    asm volatile("{ .reg .u64 w; mul.wide.u32 w, %2, %3; mov.b64 {%0, %1}, w; }" : "=r"(lo), "=r"(hi) : "r"(t), "r"(mulc));
                    asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(t) : "r"(lo), "r"(seed | 1u), "r"(hi));
                }
                if (MODE == 3) asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(t) : "r"(t), "r"(mulc));
                x[i] = t;
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) acc ^= x[i];
    if (acc == 0x12345678u) sink[0] = acc;
}

}  // namespace starkb200

using namespace starkb200;

// steps_per_s[m], m = 0..7 (5..7 = three IMAD / IMAD.HI / IMAD.WIDE per step with no ALU-pipe work): 1e12 chain steps per second for mode m (mode 0 = the three ALU instructions alone), so
// the ALU-pipe instruction rate is 3x the figure and the cost of the extra FMA-pipe work is read off the ratio to mode 0.
extern "C" int stark_measure_pipe_mix(stark_ctx* ctx, double steps_per_s[8]) {
    if (!ctx || !steps_per_s) return ST_INVALID;
    try {
        std::lock_guard<std::recursive_mutex> lk(ctx->mu);
        STARK_CUDA(cudaSetDevice(ctx->device));
        DevBuf sink(256, ctx->stream);
        cudaEvent_t e0, e1;
        STARK_CUDA(cudaEventCreate(&e0)); STARK_CUDA(cudaEventCreate(&e1));
        const int iters = 512, blocks = ctx->sm_count * 16, threads = 256;
        const uint32_t mulc = 1u << 25;
        for (int mode = 0; mode < 8; mode++) {
            double best = 0;
            for (int rep = 0; rep < 4; rep++) {
                STARK_CUDA(cudaEventRecord(e0, ctx->stream));
                uint32_t* s = sink.as<uint32_t>();
                uint32_t seed = 0x9e3779b9u + rep;
                switch (mode) {
                    case 0: pipe_mix_kernel<0><<<blocks, threads, 0, ctx->stream>>>(s, iters, seed, mulc); break;
                    case 1: pipe_mix_kernel<1><<<blocks, threads, 0, ctx->stream>>>(s, iters, seed, mulc); break;
                    case 2: pipe_mix_kernel<2><<<blocks, threads, 0, ctx->stream>>>(s, iters, seed, mulc); break;
                    case 3: pipe_mix_kernel<3><<<blocks, threads, 0, ctx->stream>>>(s, iters, seed, mulc); break;
                    case 4: pipe_mix_kernel<4><<<blocks, threads, 0, ctx->stream>>>(s, iters, seed, mulc); break;
                    case 5: pipe_mix_kernel<5><<<blocks, threads, 0, ctx->stream>>>(s, iters, seed, mulc); break;
                    case 6: pipe_mix_kernel<6><<<blocks, threads, 0, ctx->stream>>>(s, iters, seed, mulc); break;
                    default: pipe_mix_kernel<7><<<blocks, threads, 0, ctx->stream>>>(s, iters, seed, mulc); break;
                }
                STARK_CUDA(cudaEventRecord(e1, ctx->stream));
                STARK_CUDA(cudaEventSynchronize(e1));
                float ms = 0;
                STARK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
                double t = (double)blocks * threads * iters * 4 * 8 / (ms * 1e-3) / 1e12;
                if (rep > 0 && t > best) best = t;
            }
            steps_per_s[mode] = best;
            ctx->launches += 4;
        }
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        return ST_OK;
    } catch (const std::exception&) {
        return ST_CUDA;
    }
}

// tops[0]: SHF+LOP3+IADD3 only (ALU pipe);  tops[1]: the same plus one IMAD per three ALU ops (both pipes).
// Unit: 1e12 thread-level integer instructions per second.
extern "C" int stark_measure_int_peak(stark_ctx* ctx, double tops[2]) {
    if (!ctx || !tops) return ST_INVALID;
    try {
        std::lock_guard<std::recursive_mutex> lk(ctx->mu);
        STARK_CUDA(cudaSetDevice(ctx->device));
        DevBuf sink(256, ctx->stream);
        cudaEvent_t e0, e1;
        STARK_CUDA(cudaEventCreate(&e0)); STARK_CUDA(cudaEventCreate(&e1));
        const int iters = 512, blocks = ctx->sm_count * 16, threads = 256;
        for (int variant = 0; variant < 2; variant++) {
            double best = 0;
            for (int rep = 0; rep < 4; rep++) {
                STARK_CUDA(cudaEventRecord(e0, ctx->stream));
                if (variant == 0) int_peak_kernel<false><<<blocks, threads, 0, ctx->stream>>>(sink.as<uint32_t>(), iters, 0x9e3779b9u + rep, 5u);
                else int_peak_kernel<true><<<blocks, threads, 0, ctx->stream>>>(sink.as<uint32_t>(), iters, 0x9e3779b9u + rep, 5u);
                STARK_CUDA(cudaEventRecord(e1, ctx->stream));
                STARK_CUDA(cudaEventSynchronize(e1));
                float ms = 0;
                STARK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
                double ops = (double)blocks * threads * iters * 4 * 8 * (variant == 0 ? 3 : 4);
                double t = ops / (ms * 1e-3) / 1e12;
                if (rep > 0 && t > best) best = t;
            }
            tops[variant] = best;
            ctx->launches += 4;
        }
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        return ST_OK;
    } catch (const std::exception&) {
        return ST_CUDA;
    }
}
