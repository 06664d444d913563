// sha256.cuh — SHA-256 compression on the integer pipes (SHF / LOP3 / IADD3), fully unrolled.
//
// Implements exactly the hash the reference's Merkle tree uses (src/merkle/mod.rs:13-16 via
// rs_merkle 1.4.2 `algorithms::Sha256` = sha2 0.10.8):
//   leaf   = SHA-256(value.to_be_bytes())  — one 64-byte block: W0:W1 = value, W2 = 0x80000000, W15 = 64
//   parent = SHA-256(left || right)        — block 1 = the children's 16 state words,
//                                            block 2 = constant padding (W0 = 0x80000000, W15 = 512)
// Digests live in registers / HBM as the eight big-endian state WORDS, so no byte swaps are needed
// until a digest leaves the device.
#pragma once
#include <stdint.h>

namespace starkb200 {

struct Digest { uint32_t w[8]; };

__device__ __forceinline__ uint32_t rotr32(uint32_t x, unsigned n) { return __funnelshift_r(x, x, n); }
__device__ __forceinline__ uint32_t big_s0(uint32_t x) { return rotr32(x, 2) ^ rotr32(x, 13) ^ rotr32(x, 22); }
__device__ __forceinline__ uint32_t big_s1(uint32_t x) { return rotr32(x, 6) ^ rotr32(x, 11) ^ rotr32(x, 25); }
// Variant (off): the plain right shifts of the message schedule as the high word of x * 2^(32-n) with a factor
// the compiler cannot see through, i.e. IMAD.HI on the FMA pipe instead of SHF on the ALU pipe (8 -> 6 ALU
// instructions per schedule step).  Measured neutral on B200 (2^24-leaf tree 2.629 ms vs 2.627 ms), so the
// plain shift stays.
#ifndef STARK_SHA_SHR_ON_FMA
#define STARK_SHA_SHR_ON_FMA 0
#endif
static __constant__ uint32_t c_sha_two29 = 1u << 29;
static __constant__ uint32_t c_sha_two22 = 1u << 22;
#if STARK_SHA_SHR_ON_FMA
__device__ __forceinline__ uint32_t shr3(uint32_t x) { return __umulhi(x, c_sha_two29); }
__device__ __forceinline__ uint32_t shr10(uint32_t x) { return __umulhi(x, c_sha_two22); }
#else
__device__ __forceinline__ uint32_t shr3(uint32_t x) { return x >> 3; }
__device__ __forceinline__ uint32_t shr10(uint32_t x) { return x >> 10; }
#endif
__device__ __forceinline__ uint32_t sml_s0(uint32_t x) { return rotr32(x, 7) ^ rotr32(x, 18) ^ shr3(x); }
__device__ __forceinline__ uint32_t sml_s1(uint32_t x) { return rotr32(x, 17) ^ rotr32(x, 19) ^ shr10(x); }
__device__ __forceinline__ uint32_t ch(uint32_t e, uint32_t f, uint32_t g) { return (e & f) ^ (~e & g); }
__device__ __forceinline__ uint32_t maj(uint32_t a, uint32_t b, uint32_t c) { return (a & b) ^ (a & c) ^ (b & c); }

#define STARK_SHA_K                                                                                        \
    { 0x428a2f98u,0x71374491u,0xb5c0fbcfu,0xe9b5dba5u,0x3956c25bu,0x59f111f1u,0x923f82a4u,0xab1c5ed5u,      \
      0xd807aa98u,0x12835b01u,0x243185beu,0x550c7dc3u,0x72be5d74u,0x80deb1feu,0x9bdc06a7u,0xc19bf174u,      \
      0xe49b69c1u,0xefbe4786u,0x0fc19dc6u,0x240ca1ccu,0x2de92c6fu,0x4a7484aau,0x5cb0a9dcu,0x76f988dau,      \
      0x983e5152u,0xa831c66du,0xb00327c8u,0xbf597fc7u,0xc6e00bf3u,0xd5a79147u,0x06ca6351u,0x14292967u,      \
      0x27b70a85u,0x2e1b2138u,0x4d2c6dfcu,0x53380d13u,0x650a7354u,0x766a0abbu,0x81c2c92eu,0x92722c85u,      \
      0xa2bfe8a1u,0xa81a664bu,0xc24b8b70u,0xc76c51a3u,0xd192e819u,0xd6990624u,0xf40e3585u,0x106aa070u,      \
      0x19a4c116u,0x1e376c08u,0x2748774cu,0x34b0bcb5u,0x391c0cb3u,0x4ed8aa4au,0x5b9cca4fu,0x682e6ff3u,      \
      0x748f82eeu,0x78a5636fu,0x84c87814u,0x8cc70208u,0x90befffau,0xa4506cebu,0xbef9a3f7u,0xc67178f2u }

// K[i] + W[i] of the constant second block of a 64-byte message (generated offline; checked by the
// parity tests against hashlib through every Merkle root).
#define STARK_SHA_KW_PAD64                                                                                 \
    { 0xc28a2f98u,0x71374491u,0xb5c0fbcfu,0xe9b5dba5u,0x3956c25bu,0x59f111f1u,0x923f82a4u,0xab1c5ed5u,      \
      0xd807aa98u,0x12835b01u,0x243185beu,0x550c7dc3u,0x72be5d74u,0x80deb1feu,0x9bdc06a7u,0xc19bf374u,      \
      0x649b69c1u,0xf0fe4786u,0x0fe1edc6u,0x240cf254u,0x4fe9346fu,0x6cc984beu,0x61b9411eu,0x16f988fau,      \
      0xf2c65152u,0xa88e5a6du,0xb019fc65u,0xb9d99ec7u,0x9a1231c3u,0xe70eeaa0u,0xfdb1232bu,0xc7353eb0u,      \
      0x3069bad5u,0xcb976d5fu,0x5a0f118fu,0xdc1eeefdu,0x0a35b689u,0xde0b7a04u,0x58f4ca9du,0xe15d5b16u,      \
      0x007f3e86u,0x37088980u,0xa507ea32u,0x6fab9537u,0x17406110u,0x0d8cd6f1u,0xcdaa3b6du,0xc0bbbe37u,      \
      0x83613bdau,0xdb48a363u,0x0b02e931u,0x6fd15ca7u,0x521afacau,0x31338431u,0x6ed41a95u,0x6d437890u,      \
      0xc39c91f2u,0x9eccabbdu,0xb5c9a0e6u,0x532fb63cu,0xd2c741c6u,0x07237ea3u,0xa4954b68u,0x4c191d76u }

// Pipe balancing.  Rotates and boolean functions (SHF, LOP3) can only issue on the ALU pipe, which
// retires one warp instruction every 2 cycles per SM sub-partition and is what bounds this kernel.  Integer
// adds can also run on the FMA pipe as IMAD (x * 1 + y).  `c_sha_one` is a 1 the compiler cannot see
// through (constant memory), so every add below becomes an IMAD and the ALU pipe is left with the
// 10 (round) + 8 (schedule) instructions that have no other home.
static __constant__ uint32_t c_sha_one = 1;
#ifndef STARK_SHA_ADDS_ON_FMA
#define STARK_SHA_ADDS_ON_FMA 1
#endif
#if STARK_SHA_ADDS_ON_FMA == 2
// mad.lo with a literal 1: SASS `IMAD.IADD Rd, Ra, 0x1, Rb` -- FMA pipe, two register reads
__device__ __forceinline__ uint32_t sha_add_imad(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, 1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
#define SHA_ADD(x, y) sha_add_imad((x), (y))
#elif STARK_SHA_ADDS_ON_FMA
#define SHA_ADD(x, y) ((x) * sha_one + (y))
#else
#define SHA_ADD(x, y) ((x) + (y))
#endif

// Variant (off): some schedule steps do two of their three adds as ONE three-input IADD3 on the ALU pipe instead of two
// IMADs (bit j of STARK_SHA_SCHED_IADD3 selects the form for schedule word j of each group of 16).  Motivation: a
// register-chain micro-benchmark (tools/exp_pipe_mix.py) issues 3 ALU + 1 IMAD in the 6.0 cycles the ALU pipe needs, but
// 3 ALU + 2 IMAD in 7.0 (IMAD.HI / IMAD.WIDE cost ~5 / ~3 issue cycles, which is why shifts and rotates stay on the ALU
// pipe), and a compression is 1024 ALU : ~600 IMAD.  Measured on the 2^24-leaf tree: masks 0x0000 / 0x1111 / 0x5555 /
// 0x7777 / 0xFFFF -> 2.831 / 2.868 / 2.865 / 2.858 / 2.893 ms, i.e. every instruction moved back to the ALU pipe costs
// time and the all-IMAD form stays (profiles/r01_variants.txt).
#ifndef STARK_SHA_SCHED_IADD3
#define STARK_SHA_SCHED_IADD3 0x0000u
#endif
#define STARK_SHA_SCHED(w0, s0, w9, s1, j) \
    (((STARK_SHA_SCHED_IADD3 >> (j)) & 1u) ? SHA_ADD(((w0) + (s0) + (w9)), (s1)) : SHA_ADD(SHA_ADD(SHA_ADD((w0), (s0)), (w9)), (s1)))

// Operand order keeps the per-round dependency chain short: h + kw is known a round early, Ch needs one
// LOP3 after the new e, Sigma1 two dependent ALU ops, so the new e is 4 dependent instructions after the old
// one (SHF, LOP3, add, add) and the new a likewise.  Latency-bound launches (the top of every tree) run
// one warp per scheduler, so this chain is their speed.
#define STARK_SHA_ROUND(a, b, c, d, e, f, g, h, kw)                                        \
    {                                                                                      \
        uint32_t t0 = SHA_ADD((h), (kw));                                                  \
        t0 = SHA_ADD(t0, ch(e, f, g));                                                     \
        uint32_t t1 = SHA_ADD(t0, big_s1(e));                                              \
        (d) = SHA_ADD((d), t1);                                                            \
        t1 = SHA_ADD(t1, maj(a, b, c));                                                    \
        (h) = SHA_ADD(t1, big_s0(a));                                                      \
    }

// Round constants live in the constant bank: with a compile-time index they are instruction operands, with
// the loop index of the rolled rounds they are uniform-datapath loads (the index is warp-uniform), so they
// cost no ALU-pipe slot either way.
static __constant__ uint32_t c_sha_k[64] = STARK_SHA_K;
static __constant__ uint32_t c_sha_kw_pad64[64] = STARK_SHA_KW_PAD64;

#define STARK_SHA_8ROUNDS(a, b, c, d, e, f, g, h, KW)                 \
    STARK_SHA_ROUND(a, b, c, d, e, f, g, h, KW(0));                   \
    STARK_SHA_ROUND(h, a, b, c, d, e, f, g, KW(1));                   \
    STARK_SHA_ROUND(g, h, a, b, c, d, e, f, KW(2));                   \
    STARK_SHA_ROUND(f, g, h, a, b, c, d, e, KW(3));                   \
    STARK_SHA_ROUND(e, f, g, h, a, b, c, d, KW(4));                   \
    STARK_SHA_ROUND(d, e, f, g, h, a, b, c, KW(5));                   \
    STARK_SHA_ROUND(c, d, e, f, g, h, a, b, KW(6));                   \
    STARK_SHA_ROUND(b, c, d, e, f, g, h, a, KW(7));

// One compression of state `st` with the 16 message words in `w` (clobbered).  The 64 rounds are rolled
// into 4 iterations of 16 (the a..h rotation and the w[] window both close after 16 rounds), so one
// compression is ~600 instructions of code instead of ~1400: a thread that reduces a whole subtree
// (22 compressions) otherwise streams half a megabyte of straight-line code through the instruction cache.
__device__ __forceinline__ void sha256_compress(uint32_t st[8], uint32_t w[16]) {
    const uint32_t sha_one = c_sha_one; (void)sha_one;
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#define KW_A(j) SHA_ADD(w[j], c_sha_k[j])
#define KW_B(j) SHA_ADD(w[8 + j], c_sha_k[8 + j])
    STARK_SHA_8ROUNDS(a, b, c, d, e, f, g, h, KW_A)
    STARK_SHA_8ROUNDS(a, b, c, d, e, f, g, h, KW_B)
#undef KW_A
#undef KW_B
#pragma unroll 1
    for (int i = 16; i < 64; i += 16) {
#pragma unroll
        for (int j = 0; j < 16; j++)
            w[j] = STARK_SHA_SCHED(w[j], sml_s0(w[(j + 1) & 15]), w[(j + 9) & 15], sml_s1(w[(j + 14) & 15]), j);
#define KW_A(j) SHA_ADD(w[j], c_sha_k[i + j])
#define KW_B(j) SHA_ADD(w[8 + j], c_sha_k[i + 8 + j])
        STARK_SHA_8ROUNDS(a, b, c, d, e, f, g, h, KW_A)
        STARK_SHA_8ROUNDS(a, b, c, d, e, f, g, h, KW_B)
#undef KW_A
#undef KW_B
    }
    st[0] = SHA_ADD(st[0], a); st[1] = SHA_ADD(st[1], b); st[2] = SHA_ADD(st[2], c); st[3] = SHA_ADD(st[3], d);
    st[4] = SHA_ADD(st[4], e); st[5] = SHA_ADD(st[5], f); st[6] = SHA_ADD(st[6], g); st[7] = SHA_ADD(st[7], h);
}

// Compression with the constant padding block of a 64-byte message: no schedule, K+W precomputed.
__device__ __forceinline__ void sha256_compress_pad64(uint32_t st[8]) {
    const uint32_t sha_one = c_sha_one; (void)sha_one;
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll 1
    for (int i = 0; i < 64; i += 16) {
#define KW_A(j) (c_sha_kw_pad64[i + j])
#define KW_B(j) (c_sha_kw_pad64[i + 8 + j])
        STARK_SHA_8ROUNDS(a, b, c, d, e, f, g, h, KW_A)
        STARK_SHA_8ROUNDS(a, b, c, d, e, f, g, h, KW_B)
#undef KW_A
#undef KW_B
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}

__device__ __forceinline__ void sha256_init(uint32_t st[8]) {
    st[0] = 0x6a09e667u; st[1] = 0xbb67ae85u; st[2] = 0x3c6ef372u; st[3] = 0xa54ff53au;
    st[4] = 0x510e527fu; st[5] = 0x9b05688cu; st[6] = 0x1f83d9abu; st[7] = 0x5be0cd19u;
}

// leaf = SHA-256(BE8(value)) for any u64 (the reference's rule); the kernels use sha256_leaf32 below.
__device__ __forceinline__ void sha256_leaf(uint32_t hi, uint32_t lo, Digest& out) {
    uint32_t w[16];
    w[0] = hi; w[1] = lo; w[2] = 0x80000000u;
#pragma unroll
    for (int i = 3; i < 15; i++) w[i] = 0;
    w[15] = 64;
    sha256_init(out.w);
    sha256_compress(out.w, w);
}

// The same hash for value < 2^32 (every canonical element of a field with p < 2^32), specialised: the block is
// W0 = 0, W1 = value, W2 = 0x80000000, W3..W14 = 0, W15 = 64 and the state starts from the IV, so
//   * round 0 is a constant and rounds 1-3 fold to a few adds (written with plain adds so the compiler folds them),
//   * rounds 2..15 take K+W as literals,
//   * schedule words 16..31 lose every term that is sigma of a constant or a zero word (15 sigma0's, 1 sigma1),
// about 75 fewer ALU-pipe instructions than the generic compression (~7 % of a leaf hash).  Rounds 32..63 are the
// generic rolled code.
__device__ __forceinline__ void sha256_leaf32(uint32_t v, Digest& out) {
    const uint32_t sha_one = c_sha_one; (void)sha_one;
    constexpr uint32_t K[64] = STARK_SHA_K;
    uint32_t a = 0x6a09e667u, b = 0xbb67ae85u, c = 0x3c6ef372u, d = 0xa54ff53au;
    uint32_t e = 0x510e527fu, f = 0x9b05688cu, g = 0x1f83d9abu, h = 0x5be0cd19u;
#define STARK_SHA_ROUND_PLAIN(a, b, c, d, e, f, g, h, kw)                                  \
    {                                                                                      \
        uint32_t t1 = (h) + (kw) + ch(e, f, g) + big_s1(e);                                \
        (d) = (d) + t1;                                                                    \
        (h) = t1 + maj(a, b, c) + big_s0(a);                                               \
    }
    STARK_SHA_ROUND_PLAIN(a, b, c, d, e, f, g, h, K[0]);                 // W0 = 0: a compile-time constant round
    STARK_SHA_ROUND_PLAIN(h, a, b, c, d, e, f, g, K[1] + v);
    STARK_SHA_ROUND_PLAIN(g, h, a, b, c, d, e, f, K[2] + 0x80000000u);
    STARK_SHA_ROUND_PLAIN(f, g, h, a, b, c, d, e, K[3]);
#undef STARK_SHA_ROUND_PLAIN
    STARK_SHA_ROUND(e, f, g, h, a, b, c, d, K[4]);
    STARK_SHA_ROUND(d, e, f, g, h, a, b, c, K[5]);
    STARK_SHA_ROUND(c, d, e, f, g, h, a, b, K[6]);
    STARK_SHA_ROUND(b, c, d, e, f, g, h, a, K[7]);
#define KW_C(j) (K[8 + j] + ((j) == 7 ? 64u : 0u))
    STARK_SHA_8ROUNDS(a, b, c, d, e, f, g, h, KW_C)
#undef KW_C
    // schedule words 16..31 (W[t] = s1(W[t-2]) + W[t-7] + s0(W[t-15]) + W[t-16]) with the constant terms folded
    uint32_t w[16];
    constexpr uint32_t S1_64 = (64u >> 17 | 64u << 15) ^ (64u >> 19 | 64u << 13) ^ (64u >> 10);
    constexpr uint32_t S0_TOP = (0x80000000u >> 7) ^ (0x80000000u >> 18) ^ (0x80000000u >> 3);
    constexpr uint32_t S0_64 = (64u >> 7 | 64u << 25) ^ (64u >> 18 | 64u << 14) ^ (64u >> 3);
    w[0] = sml_s0(v);                                                    // W16
    w[1] = SHA_ADD(v, S1_64 + S0_TOP);                                   // W17
    w[2] = SHA_ADD(sml_s1(w[0]), 0x80000000u);                           // W18
    w[3] = sml_s1(w[1]);                                                 // W19
    w[4] = sml_s1(w[2]);                                                 // W20
    w[5] = sml_s1(w[3]);                                                 // W21
    w[6] = SHA_ADD(sml_s1(w[4]), 64u);                                   // W22
    w[7] = SHA_ADD(sml_s1(w[5]), w[0]);                                  // W23
    w[8] = SHA_ADD(sml_s1(w[6]), w[1]);                                  // W24
    w[9] = SHA_ADD(sml_s1(w[7]), w[2]);                                  // W25
    w[10] = SHA_ADD(sml_s1(w[8]), w[3]);                                 // W26
    w[11] = SHA_ADD(sml_s1(w[9]), w[4]);                                 // W27
    w[12] = SHA_ADD(sml_s1(w[10]), w[5]);                                // W28
    w[13] = SHA_ADD(sml_s1(w[11]), w[6]);                                // W29
    w[14] = SHA_ADD(SHA_ADD(sml_s1(w[12]), w[7]), S0_64);                // W30
    w[15] = SHA_ADD(SHA_ADD(SHA_ADD(sml_s1(w[13]), w[8]), sml_s0(w[0])), 64u);   // W31
#define KW_A(j) SHA_ADD(w[j], K[16 + j])
#define KW_B(j) SHA_ADD(w[8 + j], K[24 + j])
    STARK_SHA_8ROUNDS(a, b, c, d, e, f, g, h, KW_A)
    STARK_SHA_8ROUNDS(a, b, c, d, e, f, g, h, KW_B)
#undef KW_A
#undef KW_B
#pragma unroll 1
    for (int i = 32; i < 64; i += 16) {
#pragma unroll
        for (int j = 0; j < 16; j++)
            w[j] = STARK_SHA_SCHED(w[j], sml_s0(w[(j + 1) & 15]), w[(j + 9) & 15], sml_s1(w[(j + 14) & 15]), j);
#define KW_A(j) SHA_ADD(w[j], c_sha_k[i + j])
#define KW_B(j) SHA_ADD(w[8 + j], c_sha_k[i + 8 + j])
        STARK_SHA_8ROUNDS(a, b, c, d, e, f, g, h, KW_A)
        STARK_SHA_8ROUNDS(a, b, c, d, e, f, g, h, KW_B)
#undef KW_A
#undef KW_B
    }
    out.w[0] = SHA_ADD(a, 0x6a09e667u); out.w[1] = SHA_ADD(b, 0xbb67ae85u); out.w[2] = SHA_ADD(c, 0x3c6ef372u);
    out.w[3] = SHA_ADD(d, 0xa54ff53au); out.w[4] = SHA_ADD(e, 0x510e527fu); out.w[5] = SHA_ADD(f, 0x9b05688cu);
    out.w[6] = SHA_ADD(g, 0x1f83d9abu); out.w[7] = SHA_ADD(h, 0x5be0cd19u);
}

// parent = SHA-256(left || right)
__device__ __forceinline__ void sha256_node(const Digest& l, const Digest& r, Digest& out) {
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 8; i++) { w[i] = l.w[i]; w[8 + i] = r.w[i]; }
    uint32_t st[8];
    sha256_init(st);
    sha256_compress(st, w);
    sha256_compress_pad64(st);
#pragma unroll
    for (int i = 0; i < 8; i++) out.w[i] = st[i];
}

__device__ __forceinline__ Digest load_digest(const uint32_t* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    Digest d;
    d.w[0] = a.x; d.w[1] = a.y; d.w[2] = a.z; d.w[3] = a.w;
    d.w[4] = b.x; d.w[5] = b.y; d.w[6] = b.z; d.w[7] = b.w;
    return d;
}
__device__ __forceinline__ void store_digest(uint32_t* p, const Digest& d) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(d.w[0], d.w[1], d.w[2], d.w[3]);
    q[1] = make_uint4(d.w[4], d.w[5], d.w[6], d.w[7]);
}

}  // namespace starkb200
