// fri.cu — element-wise field kernels around the NTT and Merkle kernels (HBM bound):
// u64<->u32 boundary conversion, FRI folds in evaluation and coefficient space, zero-safe batched
// inverse (Montgomery trick), point-wise products, coset domain generation.
//
// Reference semantics restated here:
//   fold          src/fri/fri_commit.rs:32-50 (even + beta*odd), evaluated on the squared half
//                 domain (:18-24, :60-63) — done in evaluation space:
//                 e'[i] = (e[i]+e[i+n/2])/2 + beta*(e[i]-e[i+n/2])/(2*D[i]),  D[i] = offset*w^i
//   degree        src/polynomial/ops.rs:19-37 (trailing zeros trimmed) drives `while poly.degree >= 1`
//                 (fri_commit.rs:89), so the exact degree of the folded polynomial is tracked
//   inverse       src/fields/element.rs:54-57: a^(p-2), hence inverse(0) == 0
#include <stdlib.h>

#include "kernels.hpp"
#include "coeff_job.cuh"

namespace starkb200 {

KernelTimer::KernelTimer(stark_ctx* c, int category, double units) : ctx(c), cat(category) {
    if (!ctx->timing) return;
    auto get = [&]() {
        cudaEvent_t e;
        if (!ctx->ev_free.empty()) { e = ctx->ev_free.back(); ctx->ev_free.pop_back(); }
        else STARK_CUDA(cudaEventCreate(&e));
        return e;
    };
    e0 = get(); e1 = get();
    ctx->algo_units[cat] += units;
    cudaEventRecord(e0, ctx->stream);
}
KernelTimer::~KernelTimer() {
    if (!e0) return;
    cudaEventRecord(e1, ctx->stream);
    ctx->ev_used[cat].emplace_back(e0, e1);
}

// ---------------- boundary conversions ----------------
__global__ void narrow_kernel(const uint64_t* in, uint32_t* out, size_t n, uint32_t p) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t v = in[i];
    out[i] = v < p ? (uint32_t)v : (uint32_t)(v % p);     // FieldElement::new: value % MODULUS (element.rs:13-17)
}
__global__ void widen_kernel(const uint32_t* in, uint64_t* out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}
void narrow_u64(stark_ctx* ctx, const uint64_t* in, uint32_t* out, size_t n) {
    KernelTimer kt(ctx, stark_ctx::CAT_OTHER, 12.0 * n);
    if (!n) return;
    narrow_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(in, out, n, ctx->fp.p);
    ctx->launches++;
    STARK_CUDA(cudaGetLastError());
}
void widen_u32(stark_ctx* ctx, const uint32_t* in, uint64_t* out, size_t n) {
    KernelTimer kt(ctx, stark_ctx::CAT_OTHER, 12.0 * n);
    if (!n) return;
    widen_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(in, out, n);
    ctx->launches++;
    STARK_CUDA(cudaGetLastError());
}
// the same conversion on another stream (FRI layers streamed to the host under the hashing, api.cu: sink_push)
__global__ void widen4_kernel(const uint4* in, ulonglong2* out, size_t n4) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const uint4 v = in[i];
    out[2 * i] = make_ulonglong2(v.x, v.y);
    out[2 * i + 1] = make_ulonglong2(v.z, v.w);
}
void widen_u32_on(stark_ctx* ctx, cudaStream_t s, const uint32_t* in, uint64_t* out, size_t n) {
    if (!n) return;
    if (n % 4 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0)
        widen4_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, s>>>(reinterpret_cast<const uint4*>(in), reinterpret_cast<ulonglong2*>(out), n / 4);
    else
        widen_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(in, out, n);
    ctx->launches++;
    STARK_CUDA(cudaGetLastError());
}
void fill_zero(stark_ctx* ctx, uint32_t* p, size_t n) {
    if (n) STARK_CUDA(cudaMemsetAsync(p, 0, n * sizeof(uint32_t), ctx->stream));
}

// ---------------- coefficient-space fold + exact degree ----------------
DegScratch* deg_scratch(stark_ctx* ctx);
// {running max, ticket}: lives in the context (zeroed once); the last block of every launch resets it

// exact degree of a coefficient vector (ops.rs:19-37: trailing zeros trimmed): block max -> atomic max -> last block publishes
__global__ void poly_degree_kernel(const uint32_t* c, size_t len, HostResult* result, DegScratch* scratch) {
    int mine = 0;
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < len; j += (size_t)gridDim.x * blockDim.x)
        if (c[j] != 0) mine = (int)(j + 1);
    __shared__ int smax;
    if (threadIdx.x == 0) smax = 0;
    __syncthreads();
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) mine = max(mine, __shfl_xor_sync(0xffffffffu, mine, d));
    if ((threadIdx.x & 31) == 0 && mine) atomicMax(&smax, mine);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (smax) atomicMax(&scratch->maxv, smax);
        __threadfence();
        unsigned t = atomicAdd(&scratch->ticket, 1u);
        if (t == gridDim.x - 1) {
            __threadfence();
            result->degree_plus1 = atomicExch(&scratch->maxv, 0);
            scratch->ticket = 0;
        }
    }
}
DegScratch* deg_scratch(stark_ctx* ctx) {
    if (!ctx->deg_scratch.p) {
        ctx->deg_scratch = DevBuf(sizeof(DegScratch), ctx->stream);
        STARK_CUDA(cudaMemsetAsync(ctx->deg_scratch.p, 0, sizeof(DegScratch), ctx->stream));
    }
    return ctx->deg_scratch.as<DegScratch>();
}
void coeff_fold_job(stark_ctx* ctx, const uint32_t* c, size_t len, uint32_t beta_m, uint32_t* out, HostResult* result, unsigned cta_threads, CoeffJob& job) {
    const size_t out_len = (len + 1) / 2;
    STARK_REQUIRE(out_len >= 1 && len <= 0x7fffffffu, "coeff_fold: empty or oversized polynomial");
    job.c = c; job.out = out; job.len = (uint32_t)len; job.out_len = (uint32_t)out_len; job.beta_m = beta_m;
    const size_t want = (out_len + cta_threads - 1) / cta_threads;
    static const unsigned cap = [] { const char* e = getenv("STARK_COEFF_JOB_CTAS"); unsigned v = e ? (unsigned)atoi(e) : 0; return v ? v : COEFF_JOB_MAX_CTAS; }();
    job.ctas = (unsigned)(want < cap ? want : cap);
    job.result = result; job.scratch = deg_scratch(ctx);
}
// the coefficient fold as a launch of its own (a fold whose tree is built elsewhere has no hashing launch to ride in)
__global__ void coeff_fold_kernel(CoeffJob job, FieldParams fp) { coeff_job_run(job, blockIdx.x, fp); }
void coeff_fold(stark_ctx* ctx, const uint32_t* c, size_t len, uint32_t beta_m, uint32_t* out, HostResult* result) {
    CoeffJob job{};
    coeff_fold_job(ctx, c, len, beta_m, out, result, 256, job);
    coeff_fold_kernel<<<job.ctas, 256, 0, ctx->stream>>>(job, ctx->fp);
    ctx->launches++;
    STARK_CUDA(cudaGetLastError());
}
void poly_degree(stark_ctx* ctx, const uint32_t* c, size_t len, HostResult* result) {
    KernelTimer kt(ctx, stark_ctx::CAT_OTHER, 8.0 * len);
    STARK_REQUIRE(len > 0, "poly_degree: empty");
    poly_degree_kernel<<<(unsigned)std::min<size_t>((len + 255) / 256, 592), 256, 0, ctx->stream>>>(c, len, result, deg_scratch(ctx));
    ctx->launches++;
    STARK_CUDA(cudaGetLastError());
}

// ---------------- zero-safe batched inverse (Montgomery trick), optional numerator ----------------
// One Fermat exponentiation (63 Montgomery products for this p) per CTA of 4096 elements instead of one per element
// (element.rs:54-57) or one per thread: a thread multiplies its INV_K elements together, the 256 thread totals are
// combined with warp-shuffle prefix and suffix product scans plus a pass over the 8 warp totals in shared memory, one
// lane inverts the CTA total, and every thread gets the inverse of its own total as prefix * suffix * inverse-of-all
// and unwinds it over its elements.  Integer-bound (a Montgomery product is ~18 issue cycles, two of its five
// instructions are IMAD.HI): 3 products per element + ~2 per element of scan and inverse, against 13 per element with
// a Fermat inverse per 8 elements (2^24 elements: 0.119 -> 0.070 ms; sweep of K / CTA size / register cap in
// profiles/r01_variants.txt).
// Raw inputs are used as the Montgomery forms of x/R, which keeps every product consistent without a to_mont pass; the
// result then carries R^2 (R with a numerator), removed by one product with `fix` folded into the CTA inverse.
// Zeros are skipped (factor 1) and map to 0, like FieldElement::inverse (element.rs:54-57: 0^(p-2) = 0).
#ifndef STARK_INV_K
#define STARK_INV_K 16
#endif
#ifndef STARK_INV_THREADS
#define STARK_INV_THREADS 256
#endif
#ifndef STARK_INV_MINB
#define STARK_INV_MINB 6          // resident CTAs per SM: the CTA-wide inversion is a ~1 us serial chain that only other CTAs hide
#endif
constexpr int INV_K = STARK_INV_K;
constexpr int INV_THREADS = STARK_INV_THREADS;
__global__ void __launch_bounds__(INV_THREADS, STARK_INV_MINB)
batch_inverse_kernel(const uint32_t* a, const uint32_t* num, uint32_t* out, size_t n, FieldParams fp, uint32_t fix) {
    __shared__ uint32_t s_tot[INV_THREADS / 32], s_pre[INV_THREADS / 32], s_suf[INV_THREADS / 32], s_inv;
    const size_t base = (size_t)blockIdx.x * (INV_THREADS * INV_K) + threadIdx.x;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t x[INV_K], pre[INV_K];
    uint32_t acc = fp.one;
#pragma unroll
    for (int j = 0; j < INV_K; j++) {
        size_t i = base + (size_t)j * INV_THREADS;
        x[j] = i < n ? a[i] : 0u;                  // 0 marks "skip": zero (or out of range) contributes a factor 1
        pre[j] = acc;
        if (x[j]) acc = mont_mul(acc, x[j], fp);
    }
    // products of the thread totals before / after this thread within the warp
    uint32_t pfx = acc, sfx = acc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t up = __shfl_up_sync(0xffffffffu, pfx, d), dn = __shfl_down_sync(0xffffffffu, sfx, d);
        if (lane >= (unsigned)d) pfx = mont_mul(pfx, up, fp);
        if (lane + d < 32) sfx = mont_mul(sfx, dn, fp);
    }
    if (lane == 31) s_tot[warp] = pfx;
    uint32_t before = __shfl_up_sync(0xffffffffu, pfx, 1), after = __shfl_down_sync(0xffffffffu, sfx, 1);
    if (lane == 0) before = fp.one;
    if (lane == 31) after = fp.one;
    __syncthreads();
    if (threadIdx.x == 0) {                         // 8 warp totals: prefixes, suffixes, and the one inversion
        uint32_t run = fp.one;
        for (int w = 0; w < INV_THREADS / 32; w++) { s_pre[w] = run; run = mont_mul(run, s_tot[w], fp); }
        s_inv = mont_mul(mont_inv(run, fp), fix, fp);
        run = fp.one;
        for (int w = INV_THREADS / 32 - 1; w >= 0; w--) { s_suf[w] = run; run = mont_mul(run, s_tot[w], fp); }
    }
    __syncthreads();
    uint32_t inv = mont_mul(mont_mul(mont_mul(before, s_pre[warp], fp), mont_mul(after, s_suf[warp], fp), fp), s_inv, fp);
#pragma unroll
    for (int j = INV_K - 1; j >= 0; j--) {
        size_t i = base + (size_t)j * INV_THREADS;
        if (i >= n) continue;
        uint32_t r = 0;
        if (x[j]) {
            r = mont_mul(inv, pre[j], fp);
            inv = mont_mul(inv, x[j], fp);
            if (num) r = mont_mul(r, num[i], fp);
        }
        out[i] = r;
    }
}
void batch_inverse(stark_ctx* ctx, const uint32_t* a, const uint32_t* num, uint32_t* out, size_t n) {
    KernelTimer kt(ctx, stark_ctx::CAT_OTHER, 16.0 * n);
    if (!n) return;
    size_t per = (size_t)INV_THREADS * INV_K;
    // results come out as x^-1 * R^2: times R^-1 / R (plain output) or times 1 / R (then num / R in the last product)
    const uint64_t p = ctx->modulus;
    uint32_t fix = num ? 1u : (uint32_t)h_inv(((uint64_t)1 << 32) % p, p);
    batch_inverse_kernel<<<(unsigned)((n + per - 1) / per), INV_THREADS, 0, ctx->stream>>>(a, num, out, n, ctx->fp, fix);
    ctx->launches++;
    STARK_CUDA(cudaGetLastError());
}

__global__ void pointwise_mul_kernel(const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n, FieldParams fp) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = mont_mul(to_mont(a[i], fp), b[i], fp);
}
void pointwise_mul(stark_ctx* ctx, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n) {
    KernelTimer kt(ctx, stark_ctx::CAT_OTHER, 24.0 * n);
    if (!n) return;
    pointwise_mul_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(a, b, out, n, ctx->fp);
    ctx->launches++;
    STARK_CUDA(cudaGetLastError());
}

// ---------------- coset domain D[i] = offset * w^i (coset_fri.rs:32-36) ----------------
__global__ void coset_domain_kernel(uint32_t offset, PowTable tw, uint32_t* out, size_t n, FieldParams fp) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = mont_mul(pow_lookup(tw, (uint32_t)i, fp), offset, fp);
}
void coset_domain(stark_ctx* ctx, uint64_t offset, unsigned log_n, uint32_t* out) {
    KernelTimer kt(ctx, stark_ctx::CAT_OTHER, 8.0 * ((size_t)1 << log_n));
    const TwiddleSet& tws = ctx->twiddles(log_n);
    size_t n = (size_t)1 << log_n;
    coset_domain_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((uint32_t)(offset % ctx->modulus), tws.fwd(), out, n, ctx->fp);
    ctx->launches++;
    STARK_CUDA(cudaGetLastError());
}

// ---------------- index-dependent powers (four-step twiddles, coset scaling of a column-distributed array) -------
__global__ void pow_mul_kernel(uint32_t* v, size_t n, size_t inner_len, size_t outer0, int product, size_t inner_stride,
                               PowTable table, FieldParams fp) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    size_t inner = i % inner_len, outer = outer0 + i / inner_len;
    size_t e = product ? inner * outer : inner * inner_stride + outer;
    v[i] = mont_mul(v[i], pow_lookup(table, (uint32_t)e, fp), fp);      // e < 2^log_table by construction
}
void pow_mul(stark_ctx* ctx, uint32_t* v, size_t n, size_t inner_len, size_t outer0, bool product, size_t inner_stride,
             const PowTable& table) {
    if (!n) return;
    KernelTimer kt(ctx, stark_ctx::CAT_OTHER, 16.0 * n);
    pow_mul_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(v, n, inner_len, outer0, product, inner_stride, table, ctx->fp);
    ctx->launches++;
    STARK_CUDA(cudaGetLastError());
}

// ---------------- plain evaluation-space fold (the fused version lives in merkle.cu) ----------------
__global__ void fri_fold_kernel(LeafSource src, FieldParams fp) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= src.half) return;
    uint32_t a = src.prev[i], b = src.prev[i + src.half];
    uint32_t s = mont_mul(pow_lookup(src.winv, (uint32_t)i, fp), src.sb_m, fp);
    src.fold_out[i] = fadd(mont_mul(fadd(a, b, fp), src.inv2_m, fp), mont_mul(fsub(a, b, fp), s, fp), fp);
}
void fri_fold(stark_ctx* ctx, const LeafSource& src) {
    KernelTimer kt(ctx, stark_ctx::CAT_OTHER, 12.0 * 2 * src.half);
    if (!src.half) return;
    fri_fold_kernel<<<(unsigned)((src.half + 255) / 256), 256, 0, ctx->stream>>>(src, ctx->fp);
    ctx->launches++;
    STARK_CUDA(cudaGetLastError());
}
void fri_fold_on(stark_ctx* ctx, cudaStream_t s, const LeafSource& src) {
    if (!src.half) return;
    fri_fold_kernel<<<(unsigned)((src.half + 255) / 256), 256, 0, s>>>(src, ctx->fp);
    ctx->launches++;
    STARK_CUDA(cudaGetLastError());
}

}  // namespace starkb200
