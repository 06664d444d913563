// pipes.cu — instruction-mix micro-benchmarks behind the transform kernels' butterfly (tools/ubench/run.sh).
// Every kernel keeps 8 or 16 independent register chains per thread; the table printed is cycles per STEP per
// warp per SM sub-partition (16 resident warps per sub-partition), so a pipe that takes 2 cycles per warp instruction
// shows up as 2.0 per instruction of a step.  Inspect the SASS (cuobjdump) next to the numbers: ptxas chooses between
// IADD3 / VIADD / IMAD.IADD for a plain add on its own.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CHAINS 8
#define REPS 8

#define KERNEL_HEAD(name)                                                                   \
    __global__ void __launch_bounds__(256) name(uint32_t* sink, int iters, uint32_t p, uint32_t seed, const uint32_t* gp) { \
        uint32_t x[CHAINS], y = seed ^ threadIdx.x, z = seed + blockIdx.x;                  \
        (void)y; (void)z; (void)gp;                                                                   \
        _Pragma("unroll") for (int i = 0; i < CHAINS; i++) x[i] = seed * (i + 1) + threadIdx.x; \
        for (int it = 0; it < iters; it++) {                                                \
            _Pragma("unroll") for (int r = 0; r < REPS; r++) {                              \
                _Pragma("unroll") for (int i = 0; i < CHAINS; i++) {                        \
                    uint32_t t = x[i];
#define KERNEL_TAIL                                                                         \
                    x[i] = t;                                                               \
                }                                                                           \
            }                                                                               \
        }                                                                                   \
        uint32_t acc = 0;                                                                   \
        _Pragma("unroll") for (int i = 0; i < CHAINS; i++) acc ^= x[i];                     \
        if (acc == 0x12345678u) sink[0] = acc;                                              \
    }

// 1 step = ISETP + predicated add with a uniform operand (ptxas: VIADD)
KERNEL_HEAD(k_setp_padd_uniform)
    asm volatile("{ .reg .pred q; setp.lt.u32 q, %0, %1; @q add.u32 %0, %0, %1; }" : "+r"(t) : "r"(p));
KERNEL_TAIL
// 1 step = ISETP + predicated IMAD (x * 1 + p with an opaque 1)
KERNEL_HEAD(k_setp_pimad)
    asm volatile("{ .reg .pred q; setp.lt.u32 q, %0, %1; @q mad.lo.u32 %0, %0, %2, %1; }" : "+r"(t) : "r"(p), "r"(z | 1u));
KERNEL_TAIL
// 1 step = ISETP alone (the setp chain feeds nothing else)
KERNEL_HEAD(k_setp_lop)
    asm volatile("{ .reg .pred q; setp.lt.u32 q, %0, %1; @q lop3.b32 %0, %0, %1, %2, 0x96; }" : "+r"(t) : "r"(p), "r"(y));
KERNEL_TAIL
// 1 step = IMAD.HI alone
KERNEL_HEAD(k_hi)
    asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(t) : "r"(y));
KERNEL_TAIL
// 1 step = IMAD.HI + IMAD (lo)
KERNEL_HEAD(k_hi_imad)
    asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(t) : "r"(y));
    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(t) : "r"(z | 1u), "r"(y));
KERNEL_TAIL
// half of the chains: IMAD.HI; other half: ISETP + predicated add with a uniform operand (VIADD) -- 1 step = one of each
KERNEL_HEAD(k_hi_and_viadd)
    if (i & 1) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(t) : "r"(y));
    else asm volatile("{ .reg .pred q; setp.lt.u32 q, %0, %1; @q add.u32 %0, %0, %1; }" : "+r"(t) : "r"(p));
KERNEL_TAIL
// half IMAD.HI, half ISETP + predicated IMAD
KERNEL_HEAD(k_hi_and_pimad)
    if (i & 1) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(t) : "r"(y));
    else asm volatile("{ .reg .pred q; setp.lt.u32 q, %0, %1; @q mad.lo.u32 %0, %0, %2, %1; }" : "+r"(t) : "r"(p), "r"(z | 1u));
KERNEL_TAIL
// half IMAD.HI, half ISETP + predicated LOP3 (ALU)
KERNEL_HEAD(k_hi_and_plop)
    if (i & 1) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(t) : "r"(y));
    else asm volatile("{ .reg .pred q; setp.lt.u32 q, %0, %1; @q lop3.b32 %0, %0, %1, %2, 0x96; }" : "+r"(t) : "r"(p), "r"(y));
KERNEL_TAIL
// half IMAD.HI, half IADD3
KERNEL_HEAD(k_hi_and_add3)
    if (i & 1) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(t) : "r"(y));
    else asm volatile("{ .reg .u32 u; add.u32 u, %0, %1; add.u32 %0, u, %2; }" : "+r"(t) : "r"(p), "r"(z));
KERNEL_TAIL
// 1 step = IMAD.WIDE (both halves used) alone
KERNEL_HEAD(k_wide)
    { uint32_t lo; asm volatile("{ .reg .u64 w; mul.wide.u32 w, %2, %3; mov.b64 {%0, %1}, w; }" : "=r"(lo), "=r"(t) : "r"(t), "r"(y)); t ^= lo; }
KERNEL_TAIL
// 1 step = IADD3 alone / IADD3 + LOP3
KERNEL_HEAD(k_add3)
    asm volatile("{ .reg .u32 u; add.u32 u, %0, %1; add.u32 %0, u, %2; }" : "+r"(t) : "r"(p), "r"(z));
KERNEL_TAIL
// 1 step = LDS.32 (conflict-free, lane-strided) + IADD3
__global__ void __launch_bounds__(256) k_lds_add3(uint32_t* sink, int iters, uint32_t p, uint32_t seed, const uint32_t* gp) {
    __shared__ uint32_t sm[256 * 9];
    for (int i = threadIdx.x; i < 256 * 9; i += 256) sm[i] = seed + i;
    __syncthreads();
    uint32_t x[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < REPS; r++) {
#pragma unroll
            for (int i = 0; i < CHAINS; i++) {
                uint32_t v = sm[(threadIdx.x + 256 * i + r) ];
                asm volatile("{ .reg .u32 u; add.u32 u, %0, %1; add.u32 %0, u, %2; }" : "+r"(x[i]) : "r"(v), "r"(p));
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) acc ^= x[i];
    if (acc == 0x12345678u) sink[0] = acc;
}

// ---- the register-resident radix-16 butterfly group of the transform kernels, 32 butterflies per step ----------------
struct FP { uint32_t p, pinv; };
static __constant__ uint32_t c_zero = 0;
static __constant__ uint32_t c_one = 1;
template <int FORM> __device__ __forceinline__ uint32_t sub_fix(uint32_t a, uint32_t b, uint32_t p) {
    uint32_t d;
    if (FORM == 1) {   // correction as a predicated IMAD
        const uint32_t one = c_one;
        asm("{ .reg .pred q; .reg .u32 c;\n\t sub.cc.u32 %0, %1, %2;\n\t addc.u32 c, 0, 0;\n\t setp.eq.u32 q, c, 0;\n\t @q mad.lo.u32 %0, %3, %4, %0;\n\t}" : "=&r"(d) : "r"(a), "r"(b), "r"(one), "r"(p));
    } else if (FORM == 2) {
        const uint32_t one = c_one, pm = p - one;
        asm("{ .reg .pred q; .reg .u32 c;\n\t sub.cc.u32 %0, %1, %2;\n\t addc.u32 c, 0, 0;\n\t setp.eq.u32 q, c, 0;\n\t @q add.u32 %0, %0, %3;\n\t @q add.u32 %0, %0, %4;\n\t}" : "=&r"(d) : "r"(a), "r"(b), "r"(pm), "r"(one));
    } else {
        asm("{ .reg .pred q; .reg .u32 c;\n\t sub.cc.u32 %0, %1, %2;\n\t addc.u32 c, 0, 0;\n\t setp.eq.u32 q, c, 0;\n\t @q add.u32 %0, %0, %3;\n\t}" : "=&r"(d) : "r"(a), "r"(b), "r"(p));
    }
    return d;
}
template <int FORM> __device__ __forceinline__ uint32_t add_fix(uint32_t a, uint32_t b, uint32_t p) {
    uint32_t s; const uint32_t np = 0u - p;
    if (FORM == 1) {
        const uint32_t one = c_one;
        asm("{ .reg .pred q; .reg .u32 c;\n\t add.cc.u32 %0, %1, %2;\n\t addc.u32 c, 0, 0;\n\t setp.ne.u32 q, c, 0;\n\t @q mad.lo.u32 %0, %3, %4, %0;\n\t}" : "=&r"(s) : "r"(a), "r"(b), "r"(one), "r"(np));
    } else if (FORM == 2) {
        const uint32_t one = c_one, npm = np - one;
        asm("{ .reg .pred q; .reg .u32 c;\n\t add.cc.u32 %0, %1, %2;\n\t addc.u32 c, 0, 0;\n\t setp.ne.u32 q, c, 0;\n\t @q add.u32 %0, %0, %3;\n\t @q add.u32 %0, %0, %4;\n\t}" : "=&r"(s) : "r"(a), "r"(b), "r"(npm), "r"(one));
    } else {
        asm("{ .reg .pred q; .reg .u32 c;\n\t add.cc.u32 %0, %1, %2;\n\t addc.u32 c, 0, 0;\n\t setp.ne.u32 q, c, 0;\n\t @q add.u32 %0, %0, %3;\n\t}" : "=&r"(s) : "r"(a), "r"(b), "r"(np));
    }
    return s;
}
// MUL: 0 = production round 2 (x*wp, mad.hi with an opaque addend, mul.hi), 1 = one mul.wide whose low half feeds q
template <int MUL, int CORR> __device__ __forceinline__ uint32_t mm(uint32_t x, uint32_t w, uint32_t wp, const FP f) {
    if (MUL == 0) {
        uint32_t q = x * wp, hi; const uint32_t zero = c_zero;
        asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(hi) : "r"(x), "r"(w), "r"(zero));
        return sub_fix<CORR>(hi, __umulhi(q, f.p), f.p);
    } else {
        const uint64_t t = (uint64_t)x * w;
        const uint32_t q = (uint32_t)t * f.pinv;
        return sub_fix<CORR>((uint32_t)(t >> 32), __umulhi(q, f.p), f.p);
    }
}
// MUL as above; CM / CA / CS: where the correction of the product / lazy add / lazy subtract goes (0 = add.u32, ptxas picks; 1 = IMAD)
template <int MUL, int CM, int CA, int CS, bool PREG>
__global__ void __launch_bounds__(256) k_bfly(uint32_t* sink, int iters, FP f, uint32_t seed, const uint32_t* gp) {
    __shared__ uint2 tws[16];
    if (threadIdx.x < 16) tws[threadIdx.x] = make_uint2(seed * (threadIdx.x + 3) % f.p, seed * (threadIdx.x + 7));
    __syncthreads();
    if (PREG) { f.p = gp[threadIdx.x & 1]; }          // p in a vector register (as when the compiler keeps FieldParams there)
    uint32_t x[16];
#pragma unroll
    for (int j = 0; j < 16; j++) x[j] = (seed * (j + 1) + threadIdx.x) % f.p;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int jj = 0; jj < 16; jj++) {
                if (jj & (1 << u)) continue;
                const uint2 w = tws[((jj & ((1 << u) - 1)) << (3 - u)) + (it & 1)];
                uint32_t a = x[jj], b = mm<MUL, CM>(x[jj + (1 << u)], w.x, w.y, f);
                x[jj] = add_fix<CA>(a, b, f.p);
                x[jj + (1 << u)] = sub_fix<CS>(a, b, f.p);
            }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) acc ^= x[j];
    if (acc == 0x12345678u) sink[0] = acc;
}

typedef void (*kern_t)(uint32_t*, int, uint32_t, uint32_t, const uint32_t*);
typedef void (*kernf_t)(uint32_t*, int, FP, uint32_t, const uint32_t*);

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = prop.multiProcessorCount;
    uint32_t *sink, *gp; cudaMalloc(&sink, 256); cudaMalloc(&gp, 256);
    const uint32_t P = 3221225473u; uint32_t hp[2] = {P, P}; cudaMemcpy(gp, hp, 8, cudaMemcpyHostToDevice);
    uint32_t pinv = 1; for (int i = 0; i < 5; i++) pinv *= 2 - P * pinv;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = sms * 8, threads = 256, iters = 256;        // 8 CTAs x 8 warps = 16 warps per sub-partition
    printf("device %s, %d SMs, nominal %d MHz\n", prop.name, sms, clk_khz / 1000);
    printf("%-34s %10s %14s\n", "kernel", "ms", "ns/step/warp-slot");
    auto report = [&](const char* name, float ms, double steps_per_thread) {
        // time for one sub-partition to issue one step of one warp: ms / (steps per thread * warps per sub-partition)
        double ns = ms * 1e6 / (steps_per_thread * 16.0);
        printf("%-34s %10.4f %14.3f\n", name, ms, ns);
    };
    struct { const char* n; kern_t k; } ks[] = {
        {"setp + @add uniform (VIADD)", k_setp_padd_uniform}, {"setp + @imad", k_setp_pimad},
        {"setp + @lop3", k_setp_lop}, {"imad.hi", k_hi}, {"imad.hi + imad", k_hi_imad},
        {"4x imad.hi | 4x (setp+@viadd) [x2]", k_hi_and_viadd}, {"4x imad.hi | 4x (setp+@imad) [x2]", k_hi_and_pimad},
        {"4x imad.hi | 4x (setp+@lop3) [x2]", k_hi_and_plop}, {"4x imad.hi | 4x add3 [x2]", k_hi_and_add3},
        {"imad.wide", k_wide}, {"add3", k_add3},
        {"lds + add3", k_lds_add3},
    };
    for (auto& e : ks) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(e0); e.k<<<blocks, threads>>>(sink, iters, P, 0x9e3779b9u + rep, gp); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
        }
        report(e.n, best, (double)iters * REPS * CHAINS);
    }
    struct { const char* n; kernf_t k; } kb[] = {
        {"bfly r2 (mad.hi opaque, wp)", k_bfly<0, 0, 0, 0, false>}, {"bfly wide", k_bfly<1, 0, 0, 0, false>},
        {"bfly wide, corr(mul) on imad", k_bfly<1, 1, 0, 0, false>}, {"bfly wide, corr(mul,add) on imad", k_bfly<1, 1, 1, 0, false>},
        {"bfly wide, all corr on imad", k_bfly<1, 1, 1, 1, false>}, {"bfly wide, p in a vector reg", k_bfly<1, 0, 0, 0, true>},
        {"bfly r2, p in a vector reg", k_bfly<0, 0, 0, 0, true>},
        {"bfly wide, all corr add3", k_bfly<1, 2, 2, 2, false>}, {"bfly wide, corr(mul) add3", k_bfly<1, 2, 0, 0, false>},
        {"bfly wide, corr(mul,sub) add3", k_bfly<1, 2, 0, 2, false>}, {"bfly wide, corr(add,sub) add3", k_bfly<1, 0, 2, 2, false>},
    };
    FP f{P, pinv};
    for (auto& e : kb) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(e0); e.k<<<blocks, threads>>>(sink, iters, f, 0x9e3779b9u + rep, gp); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
        }
        report(e.n, best, (double)iters * 32);      // a step = one butterfly
    }
    cudaError_t err = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(err));
    return err != cudaSuccess;
}
