// field.cuh — device arithmetic in F_p for an odd prime p < 2^32.
//
// Replaces FieldElement<MODULUS>'s `%`-based operators (reference src/fields/element.rs:72-136) on the
// device.  HBM holds canonical values in [0,p) as u32 (the reference's domain of validity is p < 2^32,
// element.rs:45,47 multiply in u64); twiddles and other constants are kept in Montgomery form so that
// one mont_mul(x, w_mont) yields the canonical product x*w mod p with no conversions on the data.
#pragma once
#include <stdint.h>

namespace starkb200 {

struct FieldParams {
    uint32_t p;      // modulus
    uint32_t pinv;   // p^-1 mod 2^32
    uint32_t r2;     // 2^64 mod p  (to Montgomery form: mont_mul(a, r2))
    uint32_t one;    // 2^32 mod p  (1 in Montgomery form)
};

// a*b*2^-32 mod p.  Requires a*b < p*2^32 (true when one operand is < p); result in [0,p).
// Signed form (t - q*p)/2^32 so that nothing overflows for p > 2^31.
#ifndef STARK_MONT_WIDE
#define STARK_MONT_WIDE 0
#endif
__device__ __forceinline__ uint32_t mont_mul(uint32_t a, uint32_t b, const FieldParams& f) {
#if STARK_MONT_WIDE
    uint64_t t = (uint64_t)a * b;                 // one IMAD.WIDE instead of IMAD + IMAD.HI
    uint32_t lo = (uint32_t)t, hi = (uint32_t)(t >> 32);
#else
    uint32_t lo = a * b;
    uint32_t hi = __umulhi(a, b);
#endif
    uint32_t q = lo * f.pinv;
    uint32_t h = __umulhi(q, f.p);
#if defined(STARK_NTT_BFLY) && (STARK_NTT_BFLY & 4)
    uint32_t r;      // experiment: predicated correction instead of compare + select + three-input add
    asm("{ .reg .pred q;\n\t"
        "setp.lt.u32 q, %1, %2;\n\t"
        "sub.u32 %0, %1, %2;\n\t"
        "@q add.u32 %0, %0, %3;\n\t}"
        : "=&r"(r) : "r"(hi), "r"(h), "r"(f.p));
    return r;
#else
    uint32_t r = hi - h;
    return hi < h ? r + f.p : r;
#endif
}
__device__ __forceinline__ uint32_t fadd(uint32_t a, uint32_t b, const FieldParams& f) {
    uint32_t s = a + b;
    return (s < a || s >= f.p) ? s - f.p : s;
}
__device__ __forceinline__ uint32_t fsub(uint32_t a, uint32_t b, const FieldParams& f) {
    uint32_t d = a - b;
    return a < b ? d + f.p : d;
}
__device__ __forceinline__ uint32_t to_mont(uint32_t a, const FieldParams& f) { return mont_mul(a, f.r2, f); }
__device__ __forceinline__ uint32_t from_mont(uint32_t a, const FieldParams& f) { return mont_mul(a, 1u, f); }

// base^e with base in Montgomery form; result in Montgomery form.
__device__ __forceinline__ uint32_t mont_pow(uint32_t base, uint64_t e, const FieldParams& f) {
    uint32_t r = f.one;
    while (e) {
        if (e & 1) r = mont_mul(r, base, f);
        base = mont_mul(base, base, f);
        e >>= 1;
    }
    return r;
}
// Fermat inverse in Montgomery form; inverse(0) == 0 like element.rs:54-57.
__device__ __forceinline__ uint32_t mont_inv(uint32_t a, const FieldParams& f) {
    return mont_pow(a, (uint64_t)f.p - 2, f);
}

// Two-level power table: value(e) = lo[e & mask] * hi[e >> shift], both Montgomery form.
struct PowTable {
    const uint32_t* lo;
    const uint32_t* hi;
    uint32_t shift;
    uint32_t mask;
};
__device__ __forceinline__ uint32_t pow_lookup(const PowTable& t, uint32_t e, const FieldParams& f) {
    return mont_mul(__ldg(t.lo + (e & t.mask)), __ldg(t.hi + (e >> t.shift)), f);
}

}  // namespace starkb200
