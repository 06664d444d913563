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
    static constexpr bool is_ref = false;
    uint32_t p;      // modulus
    uint32_t pinv;   // p^-1 mod 2^32
    uint32_t r2;     // 2^64 mod p  (to Montgomery form: mont_mul(a, r2))
    uint32_t one;    // 2^32 mod p  (1 in Montgomery form)
};
// The reference's own field, p = 3 * 2^30 + 1 (the crate's MODULUS, src/fields/element.rs), as compile-time constants.
// The device functions below are templates over the field description: with FieldRef the modulus, its negation and
// p^-1 = 2^30 + 1 are instruction immediates; every other modulus passes the run-time FieldParams.
struct FieldRef {
    static constexpr bool is_ref = true;
    static constexpr uint32_t p = 0xC0000001u, pinv = 0x40000001u, r2 = 0x6AAAAAADu, one = 0x3FFFFFFFu;
};
static_assert((uint32_t)(FieldRef::p * FieldRef::pinv) == 1u, "p^-1 mod 2^32");
static_assert((((uint64_t)1 << 32) % FieldRef::p) == FieldRef::one && ((uint64_t)FieldRef::one * FieldRef::one) % FieldRef::p == FieldRef::r2,
              "Montgomery constants of the reference field");

// ---- conditional corrections --------------------------------------------------------------------------------
// Every modular add / subtract / Montgomery product ends in "add or subtract p if the 32-bit result went the wrong
// way".  Written as compare + select + add that is three instructions.  The forms below take the carry-out of the
// add / subtract itself as the predicate (IADD3 Rd, P0, ...) and apply the correction as ONE predicated add: ptxas maps
// `add.cc / sub.cc ; addc c,0,0 ; setp c` onto the carry predicate directly (checked in the SASS).  Which pipe the
// predicated add lands on is ptxas' choice: IADD3 (ALU pipe) or VIADD / IMAD.IADD (FMA-heavy pipe, measured in
// tools/ubench/pipes.cu) -- it alternates to balance instruction counts and cannot be steered from PTX (a predicated
// three-input add comes back as VIADD + IADD3, a predicated add-with-carry as IADD3.X + SEL).  STARK_CORR_FMA forces
// corrections onto the FMA pipe as predicated multiply-adds (`one` is a 1 the compiler cannot see through); every such
// variant measured slower (profiles/r02_ntt.md), the multiplier pipe being the one the transforms saturate.
// STARK_FIELD_CARRY=0 restores the compare + select forms (kernel experiments, tools/variants_ntt.sh).
#ifndef STARK_FIELD_CARRY
#define STARK_FIELD_CARRY 1
#endif
// which pipe applies the predicated +-p: bit 0 = Montgomery product, bit 1 = add, bit 2 = subtract; set = FMA pipe
// (predicated IMAD), clear = ALU pipe (predicated IADD3).  Measured in profiles/r02_ntt.md.
#ifndef STARK_CORR_FMA
#define STARK_CORR_FMA 0
#endif
static __constant__ uint32_t c_field_one = 1;
static __constant__ uint32_t c_field_zero = 0;
// d = a - b, plus p when a < b.  Any u32 a, b: the result is congruent to a - b; canonical when both are.
template <bool ON_FMA = ((STARK_CORR_FMA & 4) != 0)>
__device__ __forceinline__ uint32_t sub_fix(uint32_t a, uint32_t b, uint32_t p) {
#if STARK_FIELD_CARRY
    uint32_t d;
    if (ON_FMA) {
        const uint32_t one = c_field_one;
        asm("{ .reg .pred q; .reg .u32 c;\n\t"
            "sub.cc.u32 %0, %1, %2;\n\t"
            "addc.u32 c, 0, 0;\n\t"                  // the carry of a + ~b + 1: 1 when a >= b
            "setp.eq.u32 q, c, 0;\n\t"
            "@q mad.lo.u32 %0, %3, %4, %0;\n\t}"
            : "=&r"(d) : "r"(a), "r"(b), "r"(one), "r"(p));
    } else {
        asm("{ .reg .pred q; .reg .u32 c;\n\t"
            "sub.cc.u32 %0, %1, %2;\n\t"
            "addc.u32 c, 0, 0;\n\t"
            "setp.eq.u32 q, c, 0;\n\t"
            "@q add.u32 %0, %0, %3;\n\t}"
            : "=&r"(d) : "r"(a), "r"(b), "r"(p));
    }
    return d;
#else
    uint32_t d = a - b;
    return a < b ? d + p : d;
#endif
}
// s = a + b, minus p when the sum wrapped past 2^32 (a "weak" sum: congruent, not necessarily < p).
__device__ __forceinline__ uint32_t add_wrap_fix(uint32_t a, uint32_t b, uint32_t p) {
#if STARK_FIELD_CARRY
    uint32_t s;
    const uint32_t np = 0u - p;
    if (STARK_CORR_FMA & 2) {
        const uint32_t one = c_field_one;
        asm("{ .reg .pred q; .reg .u32 c;\n\t"
            "add.cc.u32 %0, %1, %2;\n\t"
            "addc.u32 c, 0, 0;\n\t"
            "setp.ne.u32 q, c, 0;\n\t"
            "@q mad.lo.u32 %0, %3, %4, %0;\n\t}"
            : "=&r"(s) : "r"(a), "r"(b), "r"(one), "r"(np));
    } else {
        asm("{ .reg .pred q; .reg .u32 c;\n\t"
            "add.cc.u32 %0, %1, %2;\n\t"
            "addc.u32 c, 0, 0;\n\t"
            "setp.ne.u32 q, c, 0;\n\t"
            "@q add.u32 %0, %0, %3;\n\t}"
            : "=&r"(s) : "r"(a), "r"(b), "r"(np));
    }
    return s;
#else
    uint32_t s = a + b;
    return s < a ? s - p : s;
#endif
}

// a*b*2^-32 mod p.  Requires a*b < p*2^32 (true when one operand is < p); result in [0,p).
// Signed form (t - q*p)/2^32 so that nothing overflows for p > 2^31.
#ifndef STARK_MONT_WIDE
#define STARK_MONT_WIDE 0
#endif
template <class F>
__device__ __forceinline__ uint32_t mont_mul(uint32_t a, uint32_t b, const F& f) {
#if STARK_MONT_WIDE
    uint64_t t = (uint64_t)a * b;                 // one IMAD.WIDE instead of IMAD + IMAD.HI
    uint32_t lo = (uint32_t)t, hi = (uint32_t)(t >> 32);
#else
    uint32_t lo = a * b;
    uint32_t hi = __umulhi(a, b);
#endif
    uint32_t q = lo * f.pinv;
    uint32_t h = __umulhi(q, f.p);
    // compare + select here: the general product sits on long dependency chains (batched inverse, power walks), where the
    // carry form measured slower (batch_inverse 2^24: 0.060 -> 0.068 ms, profiles/r02_ntt.md); the transform butterflies
    // use mont_mul_tw below
    uint32_t r = hi - h;
    return hi < h ? r + f.p : r;
}
// The product of the transform butterflies (second factor = a twiddle in Montgomery form), with the carry-predicated
// correction.  x * w is formed as ONE 64-bit multiply whose low word feeds q: ptxas cannot fold the subtraction below into
// a multiply whose own low half it depends on (with separate mul.lo / mul.hi it folds `hi - h` into IMAD.HI with a
// negated register-pair addend -- a negate, a pair move and a zeroing per product, and a carry that is wrong for h == 0;
// round 2 first avoided that with an opaque addend and a second table word w * p^-1, which cost ~1.4 register moves per
// product on the FMA-heavy pipe, the pipe that bounds these kernels: profiles/r02_ubench_pipes.txt, r02_ntt.md).
template <class F>
__device__ __forceinline__ uint32_t mont_mul_tw(uint32_t x, uint32_t w, const F& f) {
    const uint64_t t = (uint64_t)x * w;
    uint32_t q;
    if constexpr (F::is_ref) {
        // p = 3 * 2^30 + 1: p^-1 = 2^30 + 1 mod 2^32, q = lo + (lo << 30) (ptxas picks a shift-add or a multiply by the
        // immediate as its pipe balance sees fit)
        asm("{ .reg .u32 s; shl.b32 s, %1, 30; add.u32 %0, s, %1; }" : "=r"(q) : "r"((uint32_t)t));
    } else {
        q = (uint32_t)t * f.pinv;
    }
    return sub_fix<(STARK_CORR_FMA & 1) != 0>((uint32_t)(t >> 32), __umulhi(q, f.p), f.p);
}
// canonical a, b -> canonical a + b:  a + (b - p) carries out of 32 bits exactly when a + b >= p
template <class F>
__device__ __forceinline__ uint32_t fadd(uint32_t a, uint32_t b, const F& f) {
#if STARK_FIELD_CARRY
    uint32_t s;
    const uint32_t one = c_field_one, t = b - f.p;
    asm("{ .reg .pred q; .reg .u32 c;\n\t"
        "add.cc.u32 %0, %1, %2;\n\t"
        "addc.u32 c, 0, 0;\n\t"
        "setp.eq.u32 q, c, 0;\n\t"
        "@q mad.lo.u32 %0, %3, %4, %0;\n\t}"
        : "=&r"(s) : "r"(a), "r"(t), "r"(one), "r"(f.p));
    return s;
#else
    uint32_t s = a + b;
    return (s < a || s >= f.p) ? s - f.p : s;
#endif
}
template <class F>
__device__ __forceinline__ uint32_t fsub(uint32_t a, uint32_t b, const F& f) { return sub_fix(a, b, f.p); }
template <class F>
__device__ __forceinline__ uint32_t to_mont(uint32_t a, const F& f) { return mont_mul(a, f.r2, f); }
template <class F>
__device__ __forceinline__ uint32_t from_mont(uint32_t a, const F& f) { return mont_mul(a, 1u, f); }

// base^e with base in Montgomery form; result in Montgomery form.
template <class F>
__device__ __forceinline__ uint32_t mont_pow(uint32_t base, uint64_t e, const F& f) {
    uint32_t r = f.one;
    while (e) {
        if (e & 1) r = mont_mul(r, base, f);
        base = mont_mul(base, base, f);
        e >>= 1;
    }
    return r;
}
// Fermat inverse in Montgomery form; inverse(0) == 0 like element.rs:54-57.
template <class F>
__device__ __forceinline__ uint32_t mont_inv(uint32_t a, const F& f) {
    return mont_pow(a, (uint64_t)f.p - 2, f);
}

// Two-level power table: value(e) = lo[e & mask] * hi[e >> shift], both Montgomery form.
struct PowTable {
    const uint32_t* lo;
    const uint32_t* hi;
    uint32_t shift;
    uint32_t mask;
};
template <class F>
__device__ __forceinline__ uint32_t pow_lookup(const PowTable& t, uint32_t e, const F& f) {
    return mont_mul(__ldg(t.lo + (e & t.mask)), __ldg(t.hi + (e >> t.shift)), f);
}

}  // namespace starkb200
