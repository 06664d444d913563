/*
 * stark_oracle.c — CPU oracle (TEST INFRASTRUCTURE; see stark_oracle.h for the rules and pinning).
 * Every function cites the reference file:line it restates (paths relative to /root/reference).
 */
#define _GNU_SOURCE
#include "stark_oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#if defined(__x86_64__)
#include <cpuid.h>
#include <immintrin.h>
#endif

typedef unsigned __int128 u128;

/* ======================================================================================
 * field — src/fields/element.rs
 * ====================================================================================== */
uint64_t or_fe_new(uint64_t v, uint64_t M) { return v % M; }                       /* :13-17 */
uint64_t or_fe_add(uint64_t a, uint64_t b, uint64_t M) { return (a + b) % M; }     /* :72-78 (u64 add, as written) */
uint64_t or_fe_sub(uint64_t a, uint64_t b, uint64_t M) { return ((M + a - b) % M) % M; } /* :86-92 */
uint64_t or_fe_mul(uint64_t a, uint64_t b, uint64_t M) { return (uint64_t)((u128)a * b % M); } /* :102-108 */
uint64_t or_fe_neg(uint64_t a, uint64_t M) { return (M - a) % M; }                 /* :130-136 */
uint64_t or_fe_pow(uint64_t a, uint64_t e, uint64_t M) {                            /* :38-51 */
    uint64_t result = 1, base = a;
    while (e > 0) {
        if (e & 1) result = (result * base) % M;   /* u64 product: valid for M < 2^32 only, like the reference */
        base = (base * base) % M;
        e >>= 1;
    }
    return result;
}
uint64_t or_fe_inverse(uint64_t a, uint64_t M) { return or_fe_pow(a, M - 2, M); }  /* :54-57 */
uint64_t or_fe_div(uint64_t a, uint64_t b, uint64_t M) { return or_fe_mul(a, or_fe_inverse(b, M), M); } /* :116-122 */
void or_fe_to_bytes(uint64_t a, uint8_t out[8]) {                                   /* :59-61 */
    for (int i = 0; i < 8; i++) out[i] = (uint8_t)(a >> (56 - 8 * i));
}
uint64_t or_fe_from_i128(int64_t v, uint64_t M) {                                   /* :138-147 */
    __int128 m = (__int128)M, x = (__int128)v % m;
    if (x < 0) x += m;
    return (uint64_t)x % M;
}

/* ======================================================================================
 * polynomial — src/polynomial/ops.rs (literal tier)
 * ====================================================================================== */
size_t or_poly_trim(const uint64_t* c, size_t len) {                                /* :19-37, :47-60 */
    while (len > 0 && c[len - 1] == 0) len--;
    return len;
}
uint64_t or_poly_evaluate(const uint64_t* c, size_t len, uint64_t x, uint64_t M) { /* :76-83 */
    uint64_t r = 0;
    for (size_t i = len; i-- > 0;) r = or_fe_add(or_fe_mul(r, x, M), c[i], M);
    return r;
}
size_t or_poly_add(const uint64_t* a, size_t al, const uint64_t* b, size_t bl, uint64_t* out, uint64_t M) { /* :87-99 */
    size_t n = al > bl ? al : bl;
    for (size_t i = 0; i < n; i++) out[i] = or_fe_add(i < al ? a[i] : 0, i < bl ? b[i] : 0, M);
    return or_poly_trim(out, n);
}
size_t or_poly_sub(const uint64_t* a, size_t al, const uint64_t* b, size_t bl, uint64_t* out, uint64_t M) { /* :102-112 */
    size_t n = al > bl ? al : bl;
    for (size_t i = 0; i < n; i++) out[i] = or_fe_sub(i < al ? a[i] : 0, i < bl ? b[i] : 0, M);
    return or_poly_trim(out, n);
}
size_t or_poly_mul(const uint64_t* a, size_t al, const uint64_t* b, size_t bl, uint64_t* out, uint64_t M) { /* :114-138 */
    al = or_poly_trim(a, al); bl = or_poly_trim(b, bl);
    if (al == 0 || bl == 0) return 0;
    size_t n = al + bl - 1;
    uint64_t* prod = (uint64_t*)calloc(n, sizeof(uint64_t));
    for (size_t i = 0; i < al; i++) {
        if (a[i] == 0) continue;
        for (size_t j = 0; j < bl; j++) prod[i + j] = or_fe_add(prod[i + j], or_fe_mul(a[i], b[j], M), M);
    }
    memcpy(out, prod, n * sizeof(uint64_t));
    free(prod);
    return or_poly_trim(out, n);
}
int or_poly_div_rem(const uint64_t* a, size_t al, const uint64_t* b, size_t bl,
                    uint64_t* q, size_t* ql, uint64_t* r, size_t* rl, uint64_t M) {  /* :141-191 */
    al = or_poly_trim(a, al); bl = or_poly_trim(b, bl);
    if (bl == 0) return -1;                                   /* panic!("Division by zero polynomial") :142-144 */
    if (al == 0 || al < bl) {                                 /* :145-147 */
        *ql = 0; memcpy(r, a, al * sizeof(uint64_t)); *rl = al; return 0;
    }
    uint64_t* rem = (uint64_t*)malloc(al * sizeof(uint64_t));
    memcpy(rem, a, al * sizeof(uint64_t));
    long rem_deg = (long)al - 1, den_deg = (long)bl - 1;
    size_t q_len = al - bl + 1;
    memset(q, 0, q_len * sizeof(uint64_t));
    uint64_t den_lead = b[den_deg];
    size_t rlen = al;
    while (rem_deg >= den_deg && rem_deg != -1) {             /* :157 */
        uint64_t ratio = or_fe_mul(rem[rem_deg], or_fe_inverse(den_lead, M), M);  /* inverse inside the loop, :161 */
        size_t shift = (size_t)(rem_deg - den_deg);
        q[shift] = or_fe_add(q[shift], ratio, M);
        for (long i = 0; i <= den_deg; i++)
            rem[i + shift] = or_fe_sub(rem[i + shift], or_fe_mul(ratio, b[i], M), M);
        rlen = or_poly_trim(rem, rlen);                       /* :173-185 */
        rem_deg = (long)rlen - 1;
    }
    *ql = or_poly_trim(q, q_len);
    memcpy(r, rem, rlen * sizeof(uint64_t)); *rl = rlen;
    free(rem);
    return 0;
}

/* ---- src/polynomial/interpolation.rs ---- */
size_t or_poly_from_roots(const uint64_t* roots, size_t n, uint64_t* out, uint64_t M) {  /* :9-23 */
    if (n == 0) return 0;
    uint64_t* p = (uint64_t*)calloc(n + 1, sizeof(uint64_t));
    uint64_t* t = (uint64_t*)calloc(n + 1, sizeof(uint64_t));
    size_t pl = 1; p[0] = or_fe_new(1, M);
    pl = or_poly_trim(p, pl);
    for (size_t k = 0; k < n; k++) {
        uint64_t lin[2] = { or_fe_neg(roots[k], M), or_fe_new(1, M) };
        pl = or_poly_mul(p, pl, lin, 2, t, M);
        memcpy(p, t, pl * sizeof(uint64_t));
    }
    memcpy(out, p, pl * sizeof(uint64_t));
    free(p); free(t);
    return pl;
}
/* one basis polynomial, interpolation.rs:57-75; returns trimmed len or (size_t)-1 */
static size_t lagrange_one(const uint64_t* xs, size_t n, size_t i, const uint64_t* Z, size_t zl,
                           uint64_t* li, uint64_t M) {
    uint64_t denom = or_fe_new(1, M);
    for (size_t j = 0; j < n; j++) {
        if (i == j) continue;
        denom = or_fe_mul(denom, or_fe_sub(xs[i], xs[j], M), M);
    }
    uint64_t denom_inv = or_fe_inverse(denom, M);
    uint64_t div[2]; size_t dl = or_poly_from_roots(&xs[i], 1, div, M);
    uint64_t* q = (uint64_t*)calloc(zl + 1, sizeof(uint64_t));
    uint64_t* r = (uint64_t*)calloc(zl + 1, sizeof(uint64_t));
    size_t ql = 0, rl = 0;
    int rc = or_poly_div_rem(Z, zl, div, dl, q, &ql, r, &rl, M);
    if (rc != 0 || rl != 0) { free(q); free(r); return (size_t)-1; }  /* panic "Z(x) should be divisible" :69-71 */
    for (size_t k = 0; k < ql; k++) li[k] = or_fe_mul(q[k], denom_inv, M);   /* scalar_mul: no re-trim, ops.rs:194-198 */
    free(q); free(r);
    return ql;
}
int or_lagrange_basis(const uint64_t* xs, size_t n, uint64_t* out, uint64_t M) {      /* :46-78 */
    if (n == 0) return 0;
    uint64_t* Z = (uint64_t*)calloc(n + 1, sizeof(uint64_t));
    size_t zl = or_poly_from_roots(xs, n, Z, M);
    memset(out, 0, n * n * sizeof(uint64_t));
    int bad = 0;
    #pragma omp parallel for schedule(dynamic)
    for (long i = 0; i < (long)n; i++) {
        uint64_t* li = (uint64_t*)calloc(n + 1, sizeof(uint64_t));
        size_t l = lagrange_one(xs, n, (size_t)i, Z, zl, li, M);
        if (l == (size_t)-1) bad = 1; else memcpy(out + (size_t)i * n, li, (l < n ? l : n) * sizeof(uint64_t));
        free(li);
    }
    free(Z);
    return bad ? -1 : 0;
}
size_t or_poly_interpolate(const uint64_t* xs, const uint64_t* ys, size_t n, uint64_t* out, uint64_t M) { /* :121-152 */
    if (n == 0) return 0;
    uint64_t* L = (uint64_t*)calloc(n * n, sizeof(uint64_t));
    if (or_lagrange_basis(xs, n, L, M) != 0) { free(L); return (size_t)-1; }
    uint64_t* acc = (uint64_t*)calloc(n, sizeof(uint64_t));
    uint64_t* term = (uint64_t*)calloc(n, sizeof(uint64_t));
    size_t al = 0;
    for (size_t i = 0; i < n; i++) {
        /* term = L_i * ys[i] (scalar_mul keeps length; add_assign skips a term whose degree is -1, :144-150) */
        size_t ll = or_poly_trim(L + i * n, n);
        for (size_t k = 0; k < ll; k++) term[k] = or_fe_mul(L[i * n + k], ys[i], M);
        if (ll == 0) continue;
        al = or_poly_add(acc, al, term, ll, acc, M);
    }
    memcpy(out, acc, al * sizeof(uint64_t));
    free(L); free(acc); free(term);
    return al;
}

/* ======================================================================================
 * SHA-256 — FIPS 180-4 (what sha2 0.10.8 and sha256 1.5.0 compute)
 * ====================================================================================== */
static const uint32_t K256[64] = {
    0x428a2f98,0x71374491,0xb5c0fbcf,0xe9b5dba5,0x3956c25b,0x59f111f1,0x923f82a4,0xab1c5ed5,
    0xd807aa98,0x12835b01,0x243185be,0x550c7dc3,0x72be5d74,0x80deb1fe,0x9bdc06a7,0xc19bf174,
    0xe49b69c1,0xefbe4786,0x0fc19dc6,0x240ca1cc,0x2de92c6f,0x4a7484aa,0x5cb0a9dc,0x76f988da,
    0x983e5152,0xa831c66d,0xb00327c8,0xbf597fc7,0xc6e00bf3,0xd5a79147,0x06ca6351,0x14292967,
    0x27b70a85,0x2e1b2138,0x4d2c6dfc,0x53380d13,0x650a7354,0x766a0abb,0x81c2c92e,0x92722c85,
    0xa2bfe8a1,0xa81a664b,0xc24b8b70,0xc76c51a3,0xd192e819,0xd6990624,0xf40e3585,0x106aa070,
    0x19a4c116,0x1e376c08,0x2748774c,0x34b0bcb5,0x391c0cb3,0x4ed8aa4a,0x5b9cca4f,0x682e6ff3,
    0x748f82ee,0x78a5636f,0x84c87814,0x8cc70208,0x90befffa,0xa4506ceb,0xbef9a3f7,0xc67178f2 };
static const uint32_t H256[8] = { 0x6a09e667,0xbb67ae85,0x3c6ef372,0xa54ff53a,0x510e527f,0x9b05688c,0x1f83d9ab,0x5be0cd19 };

#define ROR(x, n) (((x) >> (n)) | ((x) << (32 - (n))))
static void sha256_block_c(uint32_t st[8], const uint8_t* p) {
    uint32_t w[64];
    for (int i = 0; i < 16; i++)
        w[i] = ((uint32_t)p[4*i] << 24) | ((uint32_t)p[4*i+1] << 16) | ((uint32_t)p[4*i+2] << 8) | p[4*i+3];
    for (int i = 16; i < 64; i++) {
        uint32_t s0 = ROR(w[i-15], 7) ^ ROR(w[i-15], 18) ^ (w[i-15] >> 3);
        uint32_t s1 = ROR(w[i-2], 17) ^ ROR(w[i-2], 19) ^ (w[i-2] >> 10);
        w[i] = w[i-16] + s0 + w[i-7] + s1;
    }
    uint32_t a=st[0],b=st[1],c=st[2],d=st[3],e=st[4],f=st[5],g=st[6],h=st[7];
    for (int i = 0; i < 64; i++) {
        uint32_t S1 = ROR(e, 6) ^ ROR(e, 11) ^ ROR(e, 25);
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t t1 = h + S1 + ch + K256[i] + w[i];
        uint32_t S0 = ROR(a, 2) ^ ROR(a, 13) ^ ROR(a, 22);
        uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = S0 + mj;
        h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    st[0]+=a; st[1]+=b; st[2]+=c; st[3]+=d; st[4]+=e; st[5]+=f; st[6]+=g; st[7]+=h;
}

#if defined(__x86_64__)
/* x86 SHA extensions: the code path sha2 0.10.8 selects at run time when cpuid reports `sha`. */
__attribute__((target("sha,sse4.1,ssse3")))
static void sha256_block_ni(uint32_t st[8], const uint8_t* p) {
    const __m128i MASK = _mm_set_epi64x(0x0c0d0e0f08090a0bULL, 0x0405060700010203ULL);
    __m128i TMP = _mm_loadu_si128((const __m128i*)&st[0]);
    __m128i STATE1 = _mm_loadu_si128((const __m128i*)&st[4]);
    TMP = _mm_shuffle_epi32(TMP, 0xB1);            /* CDAB */
    STATE1 = _mm_shuffle_epi32(STATE1, 0x1B);      /* EFGH */
    __m128i STATE0 = _mm_alignr_epi8(TMP, STATE1, 8);   /* ABEF */
    STATE1 = _mm_blend_epi16(STATE1, TMP, 0xF0);        /* CDGH */
    __m128i ABEF_SAVE = STATE0, CDGH_SAVE = STATE1;
    __m128i M[4];
    for (int i = 0; i < 4; i++) M[i] = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i*)(p + 16 * i)), MASK);
    for (int r = 0; r < 16; r++) {
        __m128i MSG = _mm_add_epi32(M[r & 3], _mm_loadu_si128((const __m128i*)&K256[4 * r]));
        STATE1 = _mm_sha256rnds2_epu32(STATE1, STATE0, MSG);
        if (r >= 3 && r < 15) {                     /* finish w[4(r+1)..] into M[(r+1)&3] */
            __m128i T = _mm_alignr_epi8(M[r & 3], M[(r + 3) & 3], 4);
            M[(r + 1) & 3] = _mm_sha256msg2_epu32(_mm_add_epi32(M[(r + 1) & 3], T), M[r & 3]);
        }
        MSG = _mm_shuffle_epi32(MSG, 0x0E);
        STATE0 = _mm_sha256rnds2_epu32(STATE0, STATE1, MSG);
        if (r >= 1 && r < 13) M[(r + 3) & 3] = _mm_sha256msg1_epu32(M[(r + 3) & 3], M[r & 3]);
    }
    STATE0 = _mm_add_epi32(STATE0, ABEF_SAVE);
    STATE1 = _mm_add_epi32(STATE1, CDGH_SAVE);
    TMP = _mm_shuffle_epi32(STATE0, 0x1B);         /* FEBA */
    STATE1 = _mm_shuffle_epi32(STATE1, 0xB1);      /* DCHG */
    STATE0 = _mm_blend_epi16(TMP, STATE1, 0xF0);   /* DCBA */
    STATE1 = _mm_alignr_epi8(STATE1, TMP, 8);      /* ABEF -> HGFE */
    _mm_storeu_si128((__m128i*)&st[0], STATE0);
    _mm_storeu_si128((__m128i*)&st[4], STATE1);
}
static int cpu_has_sha(void) {
    unsigned a, b, c, d;
    if (!__get_cpuid_count(7, 0, &a, &b, &c, &d)) return 0;
    return (b >> 29) & 1;
}
#endif

static int g_accel_want = 1, g_accel = -1;
static void accel_init(void) {
#if defined(__x86_64__)
    g_accel = g_accel_want && cpu_has_sha();
    if (g_accel) {   /* self-check the intrinsic path against the portable rounds before trusting it */
        uint8_t blk[64]; for (int i = 0; i < 64; i++) blk[i] = (uint8_t)(i * 37 + 11);
        uint32_t a[8], b[8]; memcpy(a, H256, 32); memcpy(b, H256, 32);
        sha256_block_c(a, blk); sha256_block_ni(b, blk);
        sha256_block_c(a, blk); sha256_block_ni(b, blk);
        if (memcmp(a, b, 32) != 0) g_accel = 0;
    }
#else
    g_accel = 0;
#endif
}
void or_sha256_set_accel(int on) { g_accel_want = on; accel_init(); }
int  or_sha256_accel_active(void) { if (g_accel < 0) accel_init(); return g_accel; }
static inline void sha256_block(uint32_t st[8], const uint8_t* p) {
#if defined(__x86_64__)
    if (g_accel > 0) { sha256_block_ni(st, p); return; }
#endif
    sha256_block_c(st, p);
}
void or_sha256(const uint8_t* msg, size_t len, uint8_t out[32]) {
    if (g_accel < 0) accel_init();
    uint32_t st[8]; memcpy(st, H256, sizeof st);
    size_t off = 0;
    for (; off + 64 <= len; off += 64) sha256_block(st, msg + off);
    uint8_t tail[128]; size_t r = len - off;
    memset(tail, 0, sizeof tail);
    memcpy(tail, msg + off, r);
    tail[r] = 0x80;
    size_t tl = (r + 9 <= 64) ? 64 : 128;
    uint64_t bits = (uint64_t)len * 8;
    for (int i = 0; i < 8; i++) tail[tl - 1 - i] = (uint8_t)(bits >> (8 * i));
    sha256_block(st, tail);
    if (tl == 128) sha256_block(st, tail + 64);
    for (int i = 0; i < 8; i++) { out[4*i] = st[i] >> 24; out[4*i+1] = st[i] >> 16; out[4*i+2] = st[i] >> 8; out[4*i+3] = st[i]; }
}
static void hex_lower(const uint8_t* b, size_t n, char* out) {   /* const-hex 1.14 / rs_merkle root_hex: lowercase */
    static const char* d = "0123456789abcdef";
    for (size_t i = 0; i < n; i++) { out[2*i] = d[b[i] >> 4]; out[2*i+1] = d[b[i] & 15]; }
    out[2*n] = 0;
}

/* ======================================================================================
 * merkle — src/merkle/mod.rs:10-26 over rs_merkle 1.4.2
 *   leaf      = SHA-256(value.to_be_bytes())                    (mod.rs:13-16; from_leaves does not re-hash)
 *   parent    = SHA-256(left || right); a lone node is promoted (rs_merkle Hasher::concat_and_hash default)
 * ====================================================================================== */
struct or_tree {
    size_t n, depth;
    size_t* len;         /* len[l] = nodes at level l, l = 0..depth */
    uint8_t** lvl;       /* lvl[l] = len[l]*32 bytes */
};
static void leaf_digest(uint64_t v, uint8_t out[32]) { uint8_t b[8]; or_fe_to_bytes(v, b); or_sha256(b, 8, out); }
static void node_digest(const uint8_t* l, const uint8_t* r, uint8_t out[32]) {
    uint8_t cat[64]; memcpy(cat, l, 32); memcpy(cat + 32, r, 32); or_sha256(cat, 64, out);
}
static void level_up(const uint8_t* in, size_t m, uint8_t* out) {
    size_t pairs = m / 2;
    #pragma omp parallel for schedule(static) if (pairs > 4096)
    for (long j = 0; j < (long)pairs; j++) node_digest(in + 64 * j, in + 64 * j + 32, out + 32 * j);
    if (m & 1) memcpy(out + 32 * pairs, in + 32 * (m - 1), 32);
}
or_tree* or_merkle_new(const uint64_t* leaves, size_t n) {
    if (n == 0) return NULL;
    if (g_accel < 0) accel_init();
    or_tree* t = (or_tree*)calloc(1, sizeof *t);
    t->n = n;
    size_t d = 0; for (size_t m = n; m > 1; m = (m + 1) / 2) d++;
    t->depth = d;
    t->len = (size_t*)calloc(d + 1, sizeof(size_t));
    t->lvl = (uint8_t**)calloc(d + 1, sizeof(uint8_t*));
    t->len[0] = n; t->lvl[0] = (uint8_t*)malloc(n * 32);
    #pragma omp parallel for schedule(static) if (n > 4096)
    for (long i = 0; i < (long)n; i++) leaf_digest(leaves[i], t->lvl[0] + 32 * i);
    for (size_t l = 1; l <= d; l++) {
        t->len[l] = (t->len[l-1] + 1) / 2;
        t->lvl[l] = (uint8_t*)malloc(t->len[l] * 32);
        level_up(t->lvl[l-1], t->len[l-1], t->lvl[l]);
    }
    return t;
}
void or_merkle_free(or_tree* t) {
    if (!t) return;
    for (size_t l = 0; l <= t->depth; l++) free(t->lvl[l]);
    free(t->lvl); free(t->len); free(t);
}
size_t or_merkle_num_leaves(const or_tree* t) { return t->n; }
size_t or_merkle_depth(const or_tree* t) { return t->depth; }
void or_merkle_root(const or_tree* t, uint8_t out[32]) { memcpy(out, t->lvl[t->depth], 32); }
void or_merkle_root_hex(const or_tree* t, char out[65]) { hex_lower(t->lvl[t->depth], 32, out); }  /* mod.rs:24-26 */
int or_merkle_node(const or_tree* t, size_t level, size_t j, uint8_t out[32]) {
    if (level > t->depth || j >= t->len[level]) return -1;
    memcpy(out, t->lvl[level] + 32 * j, 32); return 0;
}
size_t or_merkle_path(const or_tree* t, size_t idx, uint8_t* out) {
    size_t w = 0, j = idx;
    for (size_t l = 0; l < t->depth; l++, j >>= 1) {
        size_t sib = j ^ 1;
        if (sib < t->len[l]) { memcpy(out + w, t->lvl[l] + 32 * sib, 32); w += 32; }
    }
    return w;
}
int or_merkle_verify(const uint8_t root[32], size_t n_leaves, size_t idx, uint64_t value,
                     const uint8_t* path, size_t path_len) {
    uint8_t cur[32], nxt[32]; leaf_digest(value, cur);
    size_t j = idx, m = n_leaves, used = 0;
    while (m > 1) {
        size_t sib = j ^ 1;
        if (sib < m) {
            if (used + 32 > path_len) return 0;
            if (j & 1) node_digest(path + used, cur, nxt); else node_digest(cur, path + used, nxt);
            memcpy(cur, nxt, 32); used += 32;
        }
        j >>= 1; m = (m + 1) / 2;
    }
    return used == path_len && memcmp(cur, root, 32) == 0;
}
void or_merkle_root_from_digests(const uint8_t* digests, size_t n, uint8_t out[32]) {
    /* rs_merkle MerkleTree::from_leaves(&[[u8;32]]) + root(): the tree rule alone, for published rs_merkle vectors */
    if (n == 0) { memset(out, 0, 32); return; }
    uint8_t* a = (uint8_t*)malloc(n * 32); memcpy(a, digests, n * 32);
    uint8_t* b = (uint8_t*)malloc(((n + 1) / 2) * 32 + 32);
    size_t m = n;
    while (m > 1) { level_up(a, m, b); m = (m + 1) / 2; uint8_t* t = a; a = b; b = t; }
    memcpy(out, a, 32);
    free(a); free(b);
}
void or_merkle_root_only(const uint64_t* leaves, size_t n, uint8_t out[32]) {
    if (n == 0) { memset(out, 0, 32); return; }
    if (g_accel < 0) accel_init();
    uint8_t* a = (uint8_t*)malloc(n * 32);
    #pragma omp parallel for schedule(static) if (n > 4096)
    for (long i = 0; i < (long)n; i++) leaf_digest(leaves[i], a + 32 * i);
    size_t m = n;
    uint8_t* b = (uint8_t*)malloc(((n + 1) / 2) * 32 + 32);
    while (m > 1) { level_up(a, m, b); m = (m + 1) / 2; uint8_t* t = a; a = b; b = t; }
    memcpy(out, a, 32);
    free(a); free(b);
}

/* ======================================================================================
 * channel — src/channel/channel.rs
 * ====================================================================================== */
typedef struct { uint8_t* d; size_t n; } bytes_t;
typedef struct { bytes_t* v; size_t n, cap; } bvec_t;
struct or_channel { bvec_t proof, cproof; char state[65]; uint64_t M; };
static void bvec_push(bvec_t* b, const uint8_t* d, size_t n) {
    if (b->n == b->cap) { b->cap = b->cap ? 2 * b->cap : 16; b->v = (bytes_t*)realloc(b->v, b->cap * sizeof(bytes_t)); }
    b->v[b->n].d = (uint8_t*)malloc(n ? n : 1); memcpy(b->v[b->n].d, d, n); b->v[b->n].n = n; b->n++;
}
static void bvec_free(bvec_t* b) { for (size_t i = 0; i < b->n; i++) free(b->v[i].d); free(b->v); }
or_channel* or_channel_new(uint64_t M) {                                            /* :24-30: state = "" */
    or_channel* c = (or_channel*)calloc(1, sizeof *c); c->M = M; c->state[0] = 0; return c;
}
void or_channel_free(or_channel* c) { if (!c) return; bvec_free(&c->proof); bvec_free(&c->cproof); free(c); }
void or_channel_send(or_channel* c, const uint8_t* msg, size_t len) {               /* :35-44 */
    size_t sl = strlen(c->state);
    char* cat = (char*)malloc(sl + 2 * len + 1);
    memcpy(cat, c->state, sl);
    hex_lower(msg, len, cat + sl);                       /* old_state + hex::encode(message) */
    uint8_t dg[32]; or_sha256((const uint8_t*)cat, sl + 2 * len, dg);
    hex_lower(dg, 32, c->state);                         /* sha256::digest -> lowercase hex String */
    free(cat);
    bvec_push(&c->proof, msg, len);
    bvec_push(&c->cproof, msg, len);
}
uint64_t or_channel_receive_random_int(or_channel* c, uint64_t min, uint64_t max, int show) {  /* :58-84 */
    size_t sl = strlen(c->state);
    if (sl == 0) { fprintf(stderr, "oracle: receive_random_int before any send (U256::from_str_radix(\"\") is undefined)\n"); abort(); }
    uint64_t range = (max - min) + 1;                    /* :68 */
    /* num = (U256(state,16) + min) % range   (:72) — computed digit-wise; the 2^256 wrap of `+` is unreachable */
    u128 acc = 0;
    for (size_t i = 0; i < sl; i++) {
        char ch = c->state[i];
        unsigned d = (ch >= '0' && ch <= '9') ? (unsigned)(ch - '0') : (unsigned)(ch - 'a' + 10);
        acc = (acc * 16 + d) % range;
    }
    uint64_t num = (uint64_t)((acc + (u128)(min % range)) % range);
    uint8_t dg[32]; or_sha256((const uint8_t*)c->state, sl, dg);   /* :75-76 */
    hex_lower(dg, 32, c->state);
    if (show) { uint8_t b[8]; or_fe_to_bytes(num, b); bvec_push(&c->proof, b, 8); }   /* :78-80 proof only */
    return num;                                                                       /* :83 */
}
uint64_t or_channel_receive_random_field_element(or_channel* c) {                   /* :47-55 */
    uint64_t num = or_channel_receive_random_int(c, 0, c->M - 1, 0);
    uint8_t b[8]; or_fe_to_bytes(num, b); bvec_push(&c->proof, b, 8);
    return or_fe_new(num, c->M);
}
size_t or_channel_proof_size(const or_channel* c) { size_t s = 0; for (size_t i = 0; i < c->proof.n; i++) s += c->proof.v[i].n; return s; }
size_t or_channel_compressed_proof_size(const or_channel* c) { size_t s = 0; for (size_t i = 0; i < c->cproof.n; i++) s += c->cproof.v[i].n; return s; }
const char* or_channel_state(const or_channel* c) { return c->state; }
size_t or_channel_proof_len(const or_channel* c) { return c->proof.n; }
size_t or_channel_proof_msg(const or_channel* c, size_t i, const uint8_t** data) { *data = c->proof.v[i].d; return c->proof.v[i].n; }
size_t or_channel_compressed_len(const or_channel* c) { return c->cproof.n; }
size_t or_channel_compressed_msg(const or_channel* c, size_t i, const uint8_t** data) { *data = c->cproof.v[i].d; return c->cproof.v[i].n; }
size_t or_channel_proof_flat(const or_channel* c, uint8_t* out) {
    size_t w = 0;
    for (size_t i = 0; i < c->proof.n; i++) {
        uint32_t n = (uint32_t)c->proof.v[i].n;
        if (out) { out[w] = n & 255; out[w+1] = (n >> 8) & 255; out[w+2] = (n >> 16) & 255; out[w+3] = n >> 24; memcpy(out + w + 4, c->proof.v[i].d, n); }
        w += 4 + n;
    }
    return w;
}

/* ======================================================================================
 * domains — src/fri/coset_fri.rs:32-36, src/fri/fri_commit.rs:18-24, :32-50
 * ====================================================================================== */
void or_coset_domain(uint64_t offset, uint64_t omega, size_t n, uint64_t* out, uint64_t M) {
    /* D[i] = offset * omega.pow(i); the running product gives the same canonical values */
    uint64_t w = or_fe_new(1, M);
    for (size_t i = 0; i < n; i++) { out[i] = or_fe_mul(offset, w, M); w = or_fe_mul(w, omega, M); }
}
void or_next_fri_domain(const uint64_t* d, size_t n, uint64_t* out, uint64_t M) {
    for (size_t i = 0; i < n / 2; i++) out[i] = or_fe_pow(d[i], 2, M);
}
size_t or_next_fri_polynomial(const uint64_t* c, size_t len, uint64_t beta, uint64_t* out, uint64_t M) {
    /* odd = Polynomial::new(a1,a3,..) * beta ; even = Polynomial::new(a0,a2,..) ; odd + even  (:32-50) */
    size_t ne = (len + 1) / 2, no = len / 2;
    for (size_t j = 0; j < ne; j++) {
        uint64_t e = c[2 * j], o = (j < no) ? or_fe_mul(c[2 * j + 1], beta, M) : 0;
        out[j] = or_fe_add(o, e, M);
    }
    return or_poly_trim(out, ne);
}

/* ======================================================================================
 * fast tier: 32-bit Montgomery NTT (M must be an odd prime < 2^32)
 * ====================================================================================== */
typedef struct { uint32_t p, pinv, r2, one; } mont_t;
static mont_t mont_init(uint64_t M) {
    mont_t m; m.p = (uint32_t)M;
    uint32_t inv = m.p;                               /* Newton: inv = p^-1 mod 2^32 */
    for (int i = 0; i < 5; i++) inv *= 2u - m.p * inv;
    m.pinv = inv;
    m.one = (uint32_t)(((uint64_t)1 << 32) % M);
    m.r2 = (uint32_t)((u128)m.one * m.one % M);
    return m;
}
static inline uint32_t mmul(uint32_t a, uint32_t b, const mont_t* m) {   /* a*b*2^-32 mod p, inputs/outputs in [0,p) */
    uint64_t t = (uint64_t)a * b;
    uint32_t q = (uint32_t)t * m->pinv;
    uint32_t h = (uint32_t)(((uint64_t)q * m->p) >> 32);
    uint32_t hi = (uint32_t)(t >> 32);
    uint32_t r = hi - h;
    return hi < h ? r + m->p : r;
}
static inline uint32_t madd(uint32_t a, uint32_t b, const mont_t* m) { uint64_t s = (uint64_t)a + b; return (uint32_t)(s >= m->p ? s - m->p : s); }
static inline uint32_t msub(uint32_t a, uint32_t b, const mont_t* m) { return a >= b ? a - b : a + (m->p - b); }
static inline uint32_t to_mont(uint64_t a, const mont_t* m) { return mmul((uint32_t)(a % m->p), m->r2, m); }

uint64_t or_root_of_unity(uint64_t generator, unsigned log_n, uint64_t M) {
    return or_fe_pow(generator, (M - 1) >> log_n, M);
}
static int fast_ok(uint64_t M, unsigned log_n) {
    if (M >= ((uint64_t)1 << 32) || (M & 1) == 0 || M < 3) return 0;
    if (log_n > 40 || ((M - 1) & ((((uint64_t)1) << log_n) - 1)) != 0) return 0;
    return 1;
}
/* in-place radix-2 DIT over Montgomery-free canonical values; twiddles kept in Montgomery form so that
 * mmul(x, tw) is the canonical product.  a: canonical u32 values. */
static void ntt_core(uint32_t* a, unsigned log_n, uint64_t omega, const mont_t* m) {
    size_t n = (size_t)1 << log_n;
    /* bit reversal */
    for (size_t i = 0, j = 0; i < n; i++) {
        if (i < j) { uint32_t t = a[i]; a[i] = a[j]; a[j] = t; }
        size_t bit = n >> 1;
        for (; bit && (j & bit); bit >>= 1) j ^= bit;
        j ^= bit;
    }
    if (n < 2) return;
    uint32_t* tw = (uint32_t*)malloc((n / 2) * sizeof(uint32_t));   /* tw[k] = omega^k, Montgomery form */
    uint32_t w = to_mont(omega, m);
    tw[0] = m->one;
    for (size_t k = 1; k < n / 2; k++) tw[k] = mmul(tw[k - 1], w, m);
    for (unsigned s = 1; s <= log_n; s++) {
        size_t half = (size_t)1 << (s - 1), step = n >> s;
        #pragma omp parallel for schedule(static) if (n >= 65536)
        for (long b = 0; b < (long)(n / 2); b++) {
            size_t grp = (size_t)b / half, k = (size_t)b % half;
            size_t i = grp * 2 * half + k, j = i + half;
            uint32_t u = a[i], v = mmul(a[j], tw[k * step], m);
            a[i] = madd(u, v, m); a[j] = msub(u, v, m);
        }
    }
    free(tw);
}
int or_ntt(uint64_t* a, unsigned log_n, uint64_t omega, uint64_t M) {
    if (!fast_ok(M, log_n)) return -1;
    mont_t m = mont_init(M); size_t n = (size_t)1 << log_n;
    uint32_t* t = (uint32_t*)malloc(n * sizeof(uint32_t));
    for (size_t i = 0; i < n; i++) t[i] = (uint32_t)(a[i] % M);
    ntt_core(t, log_n, omega, &m);
    for (size_t i = 0; i < n; i++) a[i] = t[i];
    free(t); return 0;
}
int or_intt(uint64_t* a, unsigned log_n, uint64_t omega, uint64_t M) {
    if (!fast_ok(M, log_n)) return -1;
    mont_t m = mont_init(M); size_t n = (size_t)1 << log_n;
    uint32_t* t = (uint32_t*)malloc(n * sizeof(uint32_t));
    for (size_t i = 0; i < n; i++) t[i] = (uint32_t)(a[i] % M);
    ntt_core(t, log_n, or_fe_inverse(omega, M), &m);
    uint32_t ninv = to_mont(or_fe_inverse(n % M, M), &m);
    for (size_t i = 0; i < n; i++) a[i] = mmul(t[i], ninv, &m);
    free(t); return 0;
}
int or_coset_evaluate(const uint64_t* c, size_t len, unsigned log_n, uint64_t offset, uint64_t omega,
                      uint64_t* out, uint64_t M) {
    size_t n = (size_t)1 << log_n;
    if (!fast_ok(M, log_n) || len > n) return -1;
    mont_t m = mont_init(M);
    uint32_t* t = (uint32_t*)calloc(n, sizeof(uint32_t));
    uint32_t off = to_mont(offset, &m), pw = m.one;            /* c_j * offset^j */
    for (size_t j = 0; j < len; j++) { t[j] = mmul((uint32_t)(c[j] % M), pw, &m); pw = mmul(pw, off, &m); }
    ntt_core(t, log_n, omega, &m);
    for (size_t i = 0; i < n; i++) out[i] = t[i];
    free(t); return 0;
}
int or_coset_interpolate(const uint64_t* evals, unsigned log_n, uint64_t offset, uint64_t omega,
                         uint64_t* out, uint64_t M) {
    size_t n = (size_t)1 << log_n;
    if (!fast_ok(M, log_n)) return -1;
    mont_t m = mont_init(M);
    uint32_t* t = (uint32_t*)malloc(n * sizeof(uint32_t));
    for (size_t i = 0; i < n; i++) t[i] = (uint32_t)(evals[i] % M);
    ntt_core(t, log_n, or_fe_inverse(omega, M), &m);
    uint32_t oinv = to_mont(or_fe_inverse(offset, M), &m);
    uint32_t pw = to_mont(or_fe_inverse(n % M, M), &m);        /* n^-1 * offset^-j */
    for (size_t j = 0; j < n; j++) { out[j] = mmul(t[j], pw, &m); pw = mmul(pw, oinv, &m); }
    free(t); return 0;
}
void or_batch_inverse(uint64_t* a, size_t n, uint64_t M) {
    /* Montgomery trick; a zero stays zero because element.rs:54-57 gives 0.pow(M-2) == 0 */
    if (n == 0) return;
    uint64_t* pre = (uint64_t*)malloc(n * sizeof(uint64_t));
    uint64_t acc = or_fe_new(1, M);
    for (size_t i = 0; i < n; i++) { pre[i] = acc; if (a[i] % M != 0) acc = or_fe_mul(acc, a[i] % M, M); }
    uint64_t inv = or_fe_inverse(acc, M);
    for (size_t i = n; i-- > 0;) {
        uint64_t v = a[i] % M;
        if (v == 0) { a[i] = 0; continue; }
        a[i] = or_fe_mul(inv, pre[i], M);
        inv = or_fe_mul(inv, v, M);
    }
    free(pre);
}
void or_fri_fold_evals(const uint64_t* e, size_t n, uint64_t beta, uint64_t offset, uint64_t omega,
                       uint64_t* out, uint64_t M) {
    /* e'[i] = (e[i]+e[i+n/2])/2 + beta*(e[i]-e[i+n/2])/(2*D[i]),  D[i] = offset*omega^i   */
    size_t h = n / 2;
    uint64_t inv2 = or_fe_inverse(2 % M, M);
    uint64_t oinv = or_fe_inverse(offset, M), winv = or_fe_inverse(omega, M);
    uint64_t s0 = or_fe_mul(or_fe_mul(beta % M, inv2, M), oinv, M);
    const size_t CH = 4096;
    #pragma omp parallel for schedule(static) if (h >= 65536)
    for (long c0 = 0; c0 < (long)h; c0 += CH) {
        uint64_t s = or_fe_mul(s0, or_fe_pow(winv, (uint64_t)c0, M), M);
        size_t end = (size_t)c0 + CH < h ? (size_t)c0 + CH : h;
        for (size_t i = (size_t)c0; i < end; i++) {
            uint64_t a = e[i], b = e[i + h];
            uint64_t sum = or_fe_mul(or_fe_add(a, b, M), inv2, M);
            uint64_t dif = or_fe_mul(or_fe_sub(a, b, M), s, M);
            out[i] = or_fe_add(sum, dif, M);
            s = or_fe_mul(s, winv, M);
        }
    }
}

/* ======================================================================================
 * FRI — src/fri/fri_commit.rs (with the SURVEY 2.3 repairs: const-generic modulus, root sent as the
 * ASCII bytes of the hex string (fri_verify.rs:24-25), auth path = or_merkle_path)
 * ====================================================================================== */
struct or_fri_proof {
    size_t n_layers, cap;
    uint64_t** layers; size_t* lens; or_tree** trees;    /* fri_layers, fri_merkles  (:9-13) */
    uint64_t final_c; size_t final_len;                  /* final_poly */
};
static void fri_push(or_fri_proof* p, uint64_t* evals, size_t n, or_tree* t) {
    if (p->n_layers == p->cap) {
        p->cap = p->cap ? 2 * p->cap : 32;
        p->layers = (uint64_t**)realloc(p->layers, p->cap * sizeof(void*));
        p->lens = (size_t*)realloc(p->lens, p->cap * sizeof(size_t));
        p->trees = (or_tree**)realloc(p->trees, p->cap * sizeof(void*));
    }
    p->layers[p->n_layers] = evals; p->lens[p->n_layers] = n; p->trees[p->n_layers] = t; p->n_layers++;
}
static void send_root(or_channel* ch, const or_tree* t) {
    char hex[65]; or_merkle_root_hex(t, hex);
    or_channel_send(ch, (const uint8_t*)hex, 64);        /* channel.send(root().as_bytes()) :86,:100 */
}
or_fri_proof* or_fri_commit_literal(const uint64_t* coeffs, size_t len, const uint64_t* domain, size_t n,
                                    or_channel* ch, uint64_t M) {                    /* :72-122 */
    or_fri_proof* p = (or_fri_proof*)calloc(1, sizeof *p);
    size_t pl = or_poly_trim(coeffs, len);
    uint64_t* poly = (uint64_t*)malloc((pl + 1) * sizeof(uint64_t)); memcpy(poly, coeffs, pl * sizeof(uint64_t));
    uint64_t* dom = (uint64_t*)malloc((n + 1) * sizeof(uint64_t)); memcpy(dom, domain, n * sizeof(uint64_t));
    uint64_t* ev = (uint64_t*)malloc((n + 1) * sizeof(uint64_t));
    #pragma omp parallel for schedule(static) if (n * (pl + 1) > 100000)
    for (long i = 0; i < (long)n; i++) ev[i] = or_poly_evaluate(poly, pl, dom[i], M);   /* :78 */
    or_tree* t = or_merkle_new(ev, n);                                                   /* :79 */
    if (!t) { free(poly); free(dom); free(ev); free(p); return NULL; }
    fri_push(p, ev, n, t);
    send_root(ch, t);                                                                    /* :86 */
    while ((long)pl - 1 >= 1) {                                                          /* while poly.degree >= 1 :89 */
        uint64_t beta = or_channel_receive_random_field_element(ch);                     /* :91 */
        uint64_t* np = (uint64_t*)malloc(((pl + 1) / 2 + 1) * sizeof(uint64_t));
        size_t npl = or_next_fri_polynomial(poly, pl, beta, np, M);                      /* :94, :32-50 */
        size_t nn = n / 2;
        uint64_t* nd = (uint64_t*)malloc((nn + 1) * sizeof(uint64_t));
        or_next_fri_domain(dom, n, nd, M);                                               /* :18-24 */
        uint64_t* ne = (uint64_t*)malloc((nn + 1) * sizeof(uint64_t));
        #pragma omp parallel for schedule(static) if (nn * (npl + 1) > 100000)
        for (long i = 0; i < (long)nn; i++) ne[i] = or_poly_evaluate(np, npl, nd[i], M);  /* :60-63 */
        or_tree* nt = or_merkle_new(ne, nn);                                             /* :97 */
        if (!nt) { free(np); free(nd); free(ne); free(poly); free(dom); or_fri_free(p); return NULL; } /* root() would panic */
        send_root(ch, nt);                                                               /* :100 */
        fri_push(p, ne, nn, nt);
        free(poly); free(dom); poly = np; pl = npl; dom = nd; n = nn;
    }
    uint64_t fv = pl == 0 ? 0 : poly[0];                                                 /* :109-113 */
    uint8_t b[8]; or_fe_to_bytes(fv, b); or_channel_send(ch, b, 8);                      /* :114 */
    p->final_c = fv; p->final_len = pl;
    free(poly); free(dom);
    return p;
}
static or_fri_proof* fri_commit_fast_impl(const uint64_t* coeffs, size_t len, unsigned log_n, uint64_t offset,
                                 uint64_t omega, or_channel* ch, uint64_t M, int keep) {
    size_t n = (size_t)1 << log_n;
    size_t pl = or_poly_trim(coeffs, len);
    if (!fast_ok(M, log_n) || pl > n) return NULL;
    or_fri_proof* p = (or_fri_proof*)calloc(1, sizeof *p);
    uint64_t* poly = (uint64_t*)malloc((pl + 1) * sizeof(uint64_t)); memcpy(poly, coeffs, pl * sizeof(uint64_t));
    uint64_t* ev = (uint64_t*)malloc(n * sizeof(uint64_t));
    or_coset_evaluate(poly, pl, log_n, offset, omega, ev, M);
    char hex[65]; uint8_t root[32];
    if (keep) { or_tree* t = or_merkle_new(ev, n); fri_push(p, ev, n, t); send_root(ch, t); }
    else { or_merkle_root_only(ev, n, root); hex_lower(root, 32, hex); or_channel_send(ch, (const uint8_t*)hex, 64); }
    while ((long)pl - 1 >= 1) {
        if (n < 2) { if (!keep) free(ev); free(poly); or_fri_free(p); return NULL; }
        uint64_t beta = or_channel_receive_random_field_element(ch);
        uint64_t* np = (uint64_t*)malloc(((pl + 1) / 2 + 1) * sizeof(uint64_t));
        size_t npl = or_next_fri_polynomial(poly, pl, beta, np, M);
        size_t nn = n / 2;
        uint64_t* ne = (uint64_t*)malloc((nn ? nn : 1) * sizeof(uint64_t));
        or_fri_fold_evals(ev, n, beta, offset, omega, ne, M);
        if (nn == 0) { free(np); free(ne); if (!keep) free(ev); free(poly); or_fri_free(p); return NULL; }
        if (keep) { or_tree* nt = or_merkle_new(ne, nn); fri_push(p, ne, nn, nt); send_root(ch, nt); }
        else { or_merkle_root_only(ne, nn, root); hex_lower(root, 32, hex); or_channel_send(ch, (const uint8_t*)hex, 64); free(ev); }
        free(poly); poly = np; pl = npl; ev = ne; n = nn;
        offset = or_fe_mul(offset, offset, M); omega = or_fe_mul(omega, omega, M);
    }
    uint64_t fv = pl == 0 ? 0 : poly[0];
    uint8_t b[8]; or_fe_to_bytes(fv, b); or_channel_send(ch, b, 8);
    p->final_c = fv; p->final_len = pl;
    free(poly); if (!keep) free(ev);
    return p;
}
or_fri_proof* or_fri_commit_fast(const uint64_t* coeffs, size_t len, unsigned log_n, uint64_t offset,
                                 uint64_t omega, or_channel* ch, uint64_t M) {
    return fri_commit_fast_impl(coeffs, len, log_n, offset, omega, ch, M, 1);
}
int or_fri_commit_fast_rootonly(const uint64_t* coeffs, size_t len, unsigned log_n, uint64_t offset,
                                 uint64_t omega, or_channel* ch, uint64_t M) {
    or_fri_proof* p = fri_commit_fast_impl(coeffs, len, log_n, offset, omega, ch, M, 0);
    if (!p) return -1;
    or_fri_free(p); return 0;
}
void or_fri_free(or_fri_proof* p) {
    if (!p) return;
    for (size_t k = 0; k < p->n_layers; k++) { free(p->layers[k]); or_merkle_free(p->trees[k]); }
    free(p->layers); free(p->lens); free(p->trees); free(p);
}
size_t or_fri_num_layers(const or_fri_proof* p) { return p->n_layers; }
size_t or_fri_layer_len(const or_fri_proof* p, size_t k) { return p->lens[k]; }
const uint64_t* or_fri_layer(const or_fri_proof* p, size_t k) { return p->layers[k]; }
const or_tree* or_fri_tree(const or_fri_proof* p, size_t k) { return p->trees[k]; }
size_t or_fri_final_poly(const or_fri_proof* p, uint64_t* out) { out[0] = p->final_c; return p->final_len ? 1 : 0; }

static void send_elem_and_path(or_channel* ch, const uint64_t* evals, const or_tree* t, size_t idx) {
    uint8_t b[8]; or_fe_to_bytes(evals[idx], b); or_channel_send(ch, b, 8);
    uint8_t* path = (uint8_t*)malloc(32 * (or_merkle_depth(t) + 1));
    size_t pl = or_merkle_path(t, idx, path);
    or_channel_send(ch, path, pl);
    free(path);
}
void or_decommit_fri_layers(size_t index, const or_fri_proof* p, or_channel* ch) {      /* :137-165 */
    for (size_t k = 0; k < p->n_layers; k++) {
        size_t length = p->lens[k];
        if (length == 1) {                                   /* :147-149 — and, as written, falls through */
            uint8_t b[8]; or_fe_to_bytes(p->layers[k][0], b); or_channel_send(ch, b, 8);
        }
        size_t idx = index % length;                         /* :152 */
        size_t sib = (idx + length / 2) % length;            /* :153 */
        send_elem_and_path(ch, p->layers[k], p->trees[k], idx);   /* :156-158 */
        send_elem_and_path(ch, p->layers[k], p->trees[k], sib);   /* :161-163 */
    }
}
void or_decommit_fri(size_t num_queries, size_t max_index, const or_fri_proof* p, or_channel* ch) { /* :168-179 */
    for (size_t q = 0; q < num_queries; q++) {
        size_t idx = (size_t)or_channel_receive_random_int(ch, 0, max_index, 1);
        or_decommit_fri_layers(idx, p, ch);
    }
}

/* ======================================================================================
 * STARK-101 FibonacciSq prover — build-defined (src/trace, src/composition, src/prover are empty
 * files in the reference).  Protocol written down in DESIGN.md "cfg1"; constants from SURVEY 8c.
 * ====================================================================================== */
void or_fibsq_trace(uint64_t a1, size_t rows, uint64_t* out, uint64_t M) {
    if (rows > 0) out[0] = or_fe_new(1, M);
    if (rows > 1) out[1] = or_fe_new(a1, M);
    for (size_t i = 2; i < rows; i++)
        out[i] = or_fe_add(or_fe_mul(out[i-1], out[i-1], M), or_fe_mul(out[i-2], out[i-2], M), M);
}
/* compose p(s*X): coefficient j scaled by s^j */
static void poly_scale_arg(const uint64_t* c, size_t len, uint64_t s, uint64_t* out, uint64_t M) {
    uint64_t pw = or_fe_new(1, M);
    for (size_t j = 0; j < len; j++) { out[j] = or_fe_mul(c[j], pw, M); pw = or_fe_mul(pw, s, M); }
}
int or_stark101_prove(uint64_t a1, unsigned log_trace, unsigned log_blowup, uint64_t generator,
                      size_t num_queries, int literal, or_channel* ch, uint64_t M) {
    size_t T = (size_t)1 << log_trace, rows = T - 1, N = T << log_blowup, blow = (size_t)1 << log_blowup;
    unsigned log_N = log_trace + log_blowup;
    uint64_t g = or_root_of_unity(generator, log_trace, M);      /* trace-domain generator, order T */
    uint64_t h = or_root_of_unity(generator, log_N, M);          /* LDE-domain generator, order N */
    uint64_t w = or_fe_new(generator, M);                        /* coset offset */
    uint64_t* a = (uint64_t*)malloc(T * sizeof(uint64_t));
    or_fibsq_trace(a1, rows, a, M);
    uint64_t* G = (uint64_t*)malloc(T * sizeof(uint64_t));
    or_coset_domain(or_fe_new(1, M), g, T, G, M);
    /* ---- trace polynomial f: degree <= rows-1 through (G[i], a[i]), i < rows ---- */
    uint64_t* f = (uint64_t*)calloc(T, sizeof(uint64_t)); size_t fl;
    if (literal) {
        fl = or_poly_interpolate(G, a, rows, f, M);
        if (fl == (size_t)-1) return -1;
    } else {
        /* pick the T-th value so that the x^(T-1) coefficient vanishes: sum_i y_i g^i = 0 */
        uint64_t s = 0;
        for (size_t i = 0; i < rows; i++) s = or_fe_add(s, or_fe_mul(a[i], G[i], M), M);
        a[rows] = or_fe_mul(or_fe_neg(s, M), or_fe_inverse(G[rows], M), M);
        or_coset_interpolate(a, log_trace, or_fe_new(1, M), g, f, M);
        fl = or_poly_trim(f, T);
    }
    /* ---- commit f on the coset ---- */
    uint64_t* D = (uint64_t*)malloc(N * sizeof(uint64_t));
    or_coset_domain(w, h, N, D, M);
    uint64_t* fe = (uint64_t*)malloc(N * sizeof(uint64_t));
    if (literal) { for (size_t i = 0; i < N; i++) fe[i] = or_poly_evaluate(f, fl, D[i], M); }
    else or_coset_evaluate(f, fl, log_N, w, h, fe, M);
    or_tree* ft = or_merkle_new(fe, N);
    {   /* the statement opens the transcript: modulus, generator, sizes, query count, claimed a_{T-2} (8 BE bytes each) */
        uint8_t stmt[48];
        or_fe_to_bytes(M, stmt); or_fe_to_bytes(or_fe_new(generator, M), stmt + 8); or_fe_to_bytes(log_trace, stmt + 16);
        or_fe_to_bytes(log_blowup, stmt + 24); or_fe_to_bytes((uint64_t)num_queries, stmt + 32); or_fe_to_bytes(a[rows - 1], stmt + 40);
        or_channel_send(ch, stmt, 48);
    }
    send_root(ch, ft);
    uint64_t al[3];
    for (int k = 0; k < 3; k++) al[k] = or_channel_receive_random_field_element(ch);
    /* ---- composition polynomial ---- */
    uint64_t last = a[rows - 1];
    uint64_t x_last = G[rows - 1];                                       /* g^(T-2) */
    uint64_t ex[3] = { G[T - 3], G[T - 2], G[T - 1] };
    uint64_t* cp = (uint64_t*)calloc(N, sizeof(uint64_t)); size_t cpl;
    if (literal) {
        size_t big = 2 * T + 8;
        uint64_t *num = (uint64_t*)calloc(big, 8), *q = (uint64_t*)calloc(big, 8), *r = (uint64_t*)calloc(big, 8),
                 *t1 = (uint64_t*)calloc(big, 8), *t2 = (uint64_t*)calloc(big, 8), *acc = (uint64_t*)calloc(big, 8);
        size_t nl, ql, rl, accl = 0;
        uint64_t one = or_fe_new(1, M);
        /* p0 = (f - 1)/(X - 1) */
        nl = or_poly_sub(f, fl, &one, 1, num, M);
        uint64_t d0[2] = { or_fe_neg(one, M), one };
        or_poly_div_rem(num, nl, d0, 2, q, &ql, r, &rl, M); if (rl) return -2;
        for (size_t k = 0; k < ql; k++) t1[k] = or_fe_mul(q[k], al[0], M);
        accl = or_poly_add(acc, accl, t1, or_poly_trim(t1, ql), acc, M);
        /* p1 = (f - last)/(X - g^(T-2)) */
        nl = or_poly_sub(f, fl, &last, 1, num, M);
        uint64_t d1[2] = { or_fe_neg(x_last, M), one };
        or_poly_div_rem(num, nl, d1, 2, q, &ql, r, &rl, M); if (rl) return -3;
        for (size_t k = 0; k < ql; k++) t1[k] = or_fe_mul(q[k], al[1], M);
        accl = or_poly_add(acc, accl, t1, or_poly_trim(t1, ql), acc, M);
        /* p2 = (f(g^2 X) - f(gX)^2 - f(X)^2) * (X-g^(T-3))(X-g^(T-2))(X-g^(T-1)) / (X^T - 1) */
        uint64_t* fg = (uint64_t*)calloc(T, 8); uint64_t* fg2 = (uint64_t*)calloc(T, 8);
        poly_scale_arg(f, fl, g, fg, M); poly_scale_arg(f, fl, or_fe_mul(g, g, M), fg2, M);
        size_t l1 = or_poly_mul(fg, or_poly_trim(fg, fl), fg, or_poly_trim(fg, fl), t1, M);
        size_t l2 = or_poly_mul(f, fl, f, fl, t2, M);
        nl = or_poly_sub(fg2, or_poly_trim(fg2, fl), t1, l1, num, M);
        nl = or_poly_sub(num, nl, t2, l2, num, M);
        uint64_t exl[4]; size_t el = or_poly_from_roots(ex, 3, exl, M);
        uint64_t* num2 = (uint64_t*)calloc(big + 4, 8);
        size_t n2l = or_poly_mul(num, nl, exl, el, num2, M);
        uint64_t* zt = (uint64_t*)calloc(T + 1, 8); zt[0] = or_fe_neg(one, M); zt[T] = one;
        uint64_t* q2 = (uint64_t*)calloc(big + 4, 8); uint64_t* r2 = (uint64_t*)calloc(big + 4, 8);
        or_poly_div_rem(num2, n2l, zt, T + 1, q2, &ql, r2, &rl, M); if (rl) return -4;
        for (size_t k = 0; k < ql; k++) t1[k] = or_fe_mul(q2[k], al[2], M);
        accl = or_poly_add(acc, accl, t1, or_poly_trim(t1, ql), acc, M);
        memcpy(cp, acc, accl * 8); cpl = accl;
        free(num); free(q); free(r); free(t1); free(t2); free(acc); free(fg); free(fg2); free(num2); free(zt); free(q2); free(r2);
    } else {
        /* pointwise on the coset: f(g x_i) = fe[i+blow], f(g^2 x_i) = fe[i+2*blow] */
        uint64_t *d0 = (uint64_t*)malloc(N * 8), *d1 = (uint64_t*)malloc(N * 8), *d2 = (uint64_t*)malloc(N * 8);
        for (size_t i = 0; i < N; i++) {
            d0[i] = or_fe_sub(D[i], or_fe_new(1, M), M);
            d1[i] = or_fe_sub(D[i], x_last, M);
            d2[i] = or_fe_sub(or_fe_pow(D[i], T, M), or_fe_new(1, M), M);
        }
        or_batch_inverse(d0, N, M); or_batch_inverse(d1, N, M); or_batch_inverse(d2, N, M);
        uint64_t* ce = (uint64_t*)malloc(N * 8);
        for (size_t i = 0; i < N; i++) {
            uint64_t x = D[i], fx = fe[i], fgx = fe[(i + blow) % N], fg2x = fe[(i + 2 * blow) % N];
            uint64_t p0 = or_fe_mul(or_fe_sub(fx, or_fe_new(1, M), M), d0[i], M);
            uint64_t p1 = or_fe_mul(or_fe_sub(fx, last, M), d1[i], M);
            uint64_t n2 = or_fe_sub(or_fe_sub(fg2x, or_fe_mul(fgx, fgx, M), M), or_fe_mul(fx, fx, M), M);
            uint64_t e3 = or_fe_mul(or_fe_mul(or_fe_sub(x, ex[0], M), or_fe_sub(x, ex[1], M), M), or_fe_sub(x, ex[2], M), M);
            uint64_t p2 = or_fe_mul(or_fe_mul(n2, e3, M), d2[i], M);
            ce[i] = or_fe_add(or_fe_add(or_fe_mul(al[0], p0, M), or_fe_mul(al[1], p1, M), M), or_fe_mul(al[2], p2, M), M);
        }
        or_coset_interpolate(ce, log_N, w, h, cp, M);
        cpl = or_poly_trim(cp, N);
        free(d0); free(d1); free(d2); free(ce);
    }
    /* ---- FRI over the composition polynomial ---- */
    or_fri_proof* fp = literal ? or_fri_commit_literal(cp, cpl, D, N, ch, M)
                               : or_fri_commit_fast(cp, cpl, log_N, w, h, ch, M);
    if (!fp) return -5;
    /* ---- queries: f(x), f(gx), f(g^2 x) with paths, then the FRI layers ---- */
    for (size_t qn = 0; qn < num_queries; qn++) {
        size_t idx = (size_t)or_channel_receive_random_int(ch, 0, N - 1 - 2 * blow, 1);
        send_elem_and_path(ch, fe, ft, idx);
        send_elem_and_path(ch, fe, ft, idx + blow);
        send_elem_and_path(ch, fe, ft, idx + 2 * blow);
        or_decommit_fri_layers(idx, fp, ch);
    }
    or_fri_free(fp); or_merkle_free(ft);
    free(a); free(G); free(f); free(D); free(fe); free(cp);
    return 0;
}

int or_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void or_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
