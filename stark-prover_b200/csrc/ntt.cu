// ntt.cu — Cooley-Tukey NTT / iNTT passes in F_p (p < 2^32, Montgomery multiplication on the integer
// pipes; no tensor cores — a butterfly network is not a dense contraction).
//
// Replaces the reference's per-point Horner evaluation over a domain (src/polynomial/ops.rs:76-83 mapped
// at src/fri/fri_commit.rs:78) and its Lagrange interpolation (src/polynomial/interpolation.rs:121-152):
// on a power-of-two coset both are an NTT, and canonical results are unique, hence bit-identical.
//
// Structure.  A size-2^log_n transform is split into passes of <= 9 bits.  Each pass stages a tile of
// 2^r points x 32 columns in shared memory (33-word rows: conflict-free for both the coalesced global
// side and the butterfly side), runs the r stages as register-resident radix-8/16 groups, applies one
// inter-pass twiddle per element (two-level table, 2 loads + 1 multiply) and writes the tile back to the
// addresses it came from.  Decimation-in-time consumes bit-reversed input and produces natural order;
// decimation-in-frequency is its mirror.  The LDE chain  iNTT(DIF) -> coset scale -> NTT(DIT)  therefore
// needs no permutation at all.  Natural->natural entry points of size >= 2^10 use the digit-split transform further
// down (ntt_natural: strided in-place passes + one transposing pass, no permutation sweep); smaller and batched ones
// add one bit-reversal gather to the passes above.
#include <stdlib.h>

#include <string>

#include "kernels.hpp"

namespace starkb200 {

constexpr int NTT_C = 32;
constexpr int NTT_TS = NTT_C + 1;
// Row stride of the tiles whose load / store phases move whole 16-byte groups of a row (blow-up-by-8 passes, strided
// natural-order passes): 36 words keeps every row 16-byte aligned, so those phases are STS.128 / LDS.128 (4x fewer
// shared-memory instructions; ncu's source view had 57 % of all warp samples in the load phase of a pass, a third of them
// on the MIO queue and issue slots of its 32 scalar stores per thread) and stays conflict-free: a quarter warp's eight
// 16-byte accesses fall on rows t, t+1 (banks 4t + 8g and 4t + 4 + 8g) or on one row (banks 4 rho + 4 g4); the butterfly
// rounds read one row per warp instruction whatever the stride.
constexpr int NTT_TS16 = NTT_C + 4;

// The hot transform kernels are instantiated once more for the reference's own field (FieldRef, field.cuh), where the
// modulus is an immediate and q = lo * p^-1 is a shift-add; every other modulus runs the same kernels on FieldParams.
template <bool PREF> struct FieldOf { using type = FieldParams; __device__ static FieldParams make(const FieldParams& fp) { return fp; } };
template <> struct FieldOf<true> { using type = FieldRef; __device__ static FieldRef make(const FieldParams&) { return FieldRef{}; } };
static bool is_ref_field(const stark_ctx* ctx) { return ctx->fp.p == FieldRef::p; }

struct NttPass {
    const uint32_t* src;
    uint32_t* dst;
    unsigned log_n;
    unsigned lo;          // lowest ADDRESS bit of this pass (0: contiguous tile)
    unsigned col_bits;    // lowest address bits that index independent transforms (column-batched); 0 = one transform
    unsigned ncols;       // valid columns per tile
    unsigned log_pad;     // DIT/contiguous only: input is zero-padded by 2^log_pad
    unsigned log_m;       // log_n - log_pad
    int has_scale;
    PowTable scale;
    PowTable tw;
    const uint32_t* small;
    unsigned small_log;
};

// Lazy (DIT only, and only for p > 2^31 -- the launcher checks): values are "weak" -- any u32 congruent to the
// element (2^32 < 2p, so canonical or canonical + p).  A Montgomery product accepts a weak operand and returns a canonical one, so in
// a + w*b / a - w*b only `a` is weak; the sum wraps past 2^32 at most once and the difference goes negative
// at most once: one carry-predicated correction each (field.cuh: add_wrap_fix / sub_fix).  The last pass canonicalises
// on store.  A butterfly is 9 instructions: the product (IMAD.WIDE, IMAD, IMAD.HI, IADD3 with carry-out, predicated add)
// and two IADD3-with-carry + predicated-add pairs.  The multiplies can only run on the FMA-heavy pipe (2 / 4 / 4 cycles per
// warp instruction) and ptxas places about half of the predicated adds there too (as VIADD / IMAD.IADD): that pipe, not
// the ALU pipe, bounds these kernels (profiles/r02_ntt.md, tools/ubench/pipes.cu).
template <class F>
__device__ __forceinline__ uint32_t ladd(uint32_t a_weak, uint32_t b, const F& f) { return add_wrap_fix(a_weak, b, f.p); }
template <class F>
__device__ __forceinline__ uint32_t lsub(uint32_t a_weak, uint32_t b, const F& f) { return sub_fix(a_weak, b, f.p); }
template <class F>
__device__ __forceinline__ uint32_t canonical(uint32_t x, const F& f) { return x >= f.p ? x - f.p : x; }

// In-tile twiddles w_R^k, k < R/2 (Montgomery form), in shared memory: a warp's lanes are columns of one butterfly, so
// every twiddle load of a register group is a broadcast.
__device__ __forceinline__ void fill_tile_twiddles(uint32_t* tws, const uint32_t* small, unsigned small_log, int r_log) {
    for (int k = threadIdx.x; k < (1 << r_log) >> 1; k += blockDim.x) tws[k] = small[(size_t)k << (small_log - r_log)];
}

template <int G, bool DIF, bool LAZY, int TS, class F>
__device__ __forceinline__ void butterfly_group(uint32_t* col, const int s, const int base, const uint32_t* tws,
                                                const int r_log, const F& fp) {
    uint32_t x[1 << G];
#pragma unroll
    for (int j = 0; j < (1 << G); j++) x[j] = col[(base + (j << s)) * TS];
    const int bl = base & ((1 << s) - 1);
#pragma unroll
    for (int uu = 0; uu < G; uu++) {
        const int u = DIF ? (G - 1 - uu) : uu;
        const int level = s + u + 1;                 // butterflies of span 2^(s+u) use w_{2^level}
#pragma unroll
        for (int jj = 0; jj < (1 << G); jj++) {
            if (jj & (1 << u)) continue;
            const int k = bl + ((jj & ((1 << u) - 1)) << s);
            const bool unit = (s == 0 && u == 0);    // w_2^0 = 1: no multiplication in the span-1 stage
            const uint32_t w = unit ? 0u : tws[k << (r_log - level)];
            if (DIF) {
                uint32_t a = x[jj], b = x[jj + (1 << u)];
                x[jj] = fadd(a, b, fp);
                uint32_t d = fsub(a, b, fp);
                x[jj + (1 << u)] = unit ? d : mont_mul_tw(d, w, fp);
            } else if (LAZY) {
                // (the span-1 stage is the first of a pass: its inputs come straight from a product, so `b` is canonical)
                uint32_t a = x[jj], b = unit ? x[jj + (1 << u)] : mont_mul_tw(x[jj + (1 << u)], w, fp);
                x[jj] = ladd(a, b, fp);
                x[jj + (1 << u)] = lsub(a, b, fp);
            } else {
                uint32_t a = x[jj], b = unit ? x[jj + (1 << u)] : mont_mul_tw(x[jj + (1 << u)], w, fp);
                x[jj] = fadd(a, b, fp);
                x[jj + (1 << u)] = fsub(a, b, fp);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < (1 << G); j++) col[(base + (j << s)) * TS] = x[j];
}

template <int R_LOG, int G, bool DIF, bool LAZY, int TS, class F>
__device__ __forceinline__ void run_round(uint32_t* tile, const int s, const uint32_t* tws, const F& fp,
                                          const unsigned ncols) {
    constexpr int items = ((1 << R_LOG) >> G) * NTT_C;
    for (int w = threadIdx.x; w < items; w += blockDim.x) {
        const int c = w % NTT_C, bi = w / NTT_C;
        if ((unsigned)c >= ncols) continue;
        const int base = ((bi >> s) << (s + G)) | (bi & ((1 << s) - 1));
        butterfly_group<G, DIF, LAZY, TS>(tile + c, s, base, tws, R_LOG, fp);
    }
    __syncthreads();
}

template <int R_LOG, bool DIF, bool LAZY = false, int TS = NTT_TS, class F = FieldParams>
__device__ __forceinline__ void run_rounds(uint32_t* tile, const uint32_t* tws, const F& fp, unsigned ncols) {
    // stage groups (sum = R_LOG); DIT walks spans upward, DIF downward
    if constexpr (R_LOG <= 4) {
        run_round<R_LOG, R_LOG, DIF, LAZY, TS>(tile, 0, tws, fp, ncols);
    } else if constexpr (R_LOG == 5) {
        if (!DIF) { run_round<5, 3, DIF, LAZY, TS>(tile, 0, tws, fp, ncols); run_round<5, 2, DIF, LAZY, TS>(tile, 3, tws, fp, ncols); }
        else      { run_round<5, 2, DIF, LAZY, TS>(tile, 3, tws, fp, ncols); run_round<5, 3, DIF, LAZY, TS>(tile, 0, tws, fp, ncols); }
    } else if constexpr (R_LOG == 6) {
        if (!DIF) { run_round<6, 3, DIF, LAZY, TS>(tile, 0, tws, fp, ncols); run_round<6, 3, DIF, LAZY, TS>(tile, 3, tws, fp, ncols); }
        else      { run_round<6, 3, DIF, LAZY, TS>(tile, 3, tws, fp, ncols); run_round<6, 3, DIF, LAZY, TS>(tile, 0, tws, fp, ncols); }
    } else if constexpr (R_LOG == 7) {
        if (!DIF) { run_round<7, 4, DIF, LAZY, TS>(tile, 0, tws, fp, ncols); run_round<7, 3, DIF, LAZY, TS>(tile, 4, tws, fp, ncols); }
        else      { run_round<7, 3, DIF, LAZY, TS>(tile, 4, tws, fp, ncols); run_round<7, 4, DIF, LAZY, TS>(tile, 0, tws, fp, ncols); }
    } else if constexpr (R_LOG == 8) {
        if (!DIF) { run_round<8, 4, DIF, LAZY, TS>(tile, 0, tws, fp, ncols); run_round<8, 4, DIF, LAZY, TS>(tile, 4, tws, fp, ncols); }
        else      { run_round<8, 4, DIF, LAZY, TS>(tile, 4, tws, fp, ncols); run_round<8, 4, DIF, LAZY, TS>(tile, 0, tws, fp, ncols); }
    } else {
        static_assert(R_LOG == 9, "pass width");
        if (!DIF) { run_round<9, 3, DIF, LAZY, TS>(tile, 0, tws, fp, ncols); run_round<9, 3, DIF, LAZY, TS>(tile, 3, tws, fp, ncols); run_round<9, 3, DIF, LAZY, TS>(tile, 6, tws, fp, ncols); }
        else      { run_round<9, 3, DIF, LAZY, TS>(tile, 6, tws, fp, ncols); run_round<9, 3, DIF, LAZY, TS>(tile, 3, tws, fp, ncols); run_round<9, 3, DIF, LAZY, TS>(tile, 0, tws, fp, ncols); }
    }
}

__device__ __forceinline__ uint32_t bitrev_bits(uint32_t x, unsigned bits) { return bits ? (__brev(x) >> (32 - bits)) : 0u; }

template <int R_LOG, bool DIF, bool STRIDED>
__global__ void ntt_pass_kernel(NttPass ps, FieldParams fp) {
    extern __shared__ uint32_t smem[];
    constexpr int R = 1 << R_LOG;
    uint32_t* tile = smem;                    // [R][33]
    uint32_t* tws = smem + R * NTT_TS;        // [R/2] twiddles of w_R
    fill_tile_twiddles(tws, ps.small, ps.small_log, R_LOG);

    const size_t tile_id = blockIdx.x;
    size_t gbase;
    uint32_t low0 = 0;
    if (STRIDED) {
        const size_t tiles_per_high = ((size_t)1 << ps.lo) / NTT_C;
        const size_t high = tile_id / tiles_per_high;
        low0 = (uint32_t)(tile_id % tiles_per_high) * NTT_C;
        gbase = (high << (ps.lo + R_LOG)) | low0;
    } else {
        gbase = tile_id * (size_t)R * ps.ncols;
    }
    const unsigned tw_shift = ps.log_n - (ps.lo - ps.col_bits) - R_LOG;   // exponent scale into the size-2^log_n tables
    const int total = R * (int)ps.ncols;

    // ---- load (coalesced) ----
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        int t, c;
        size_t g;
        if (STRIDED) { t = i / NTT_C; c = i % NTT_C; g = gbase + ((size_t)t << ps.lo) + c; }
        else         { c = i >> R_LOG; t = i & (R - 1); g = gbase + i; }
        uint32_t x;
        if (!STRIDED && !DIF && ps.log_pad) {
            if (g & (((size_t)1 << ps.log_pad) - 1)) x = 0;
            else {
                uint32_t q = (uint32_t)(g >> ps.log_pad);
                x = ps.src[q];
                if (ps.has_scale) x = mont_mul(x, pow_lookup(ps.scale, bitrev_bits(q, ps.log_m), fp), fp);
            }
        } else {
            x = ps.src[g];
            if (!STRIDED && !DIF && ps.has_scale) x = mont_mul(x, pow_lookup(ps.scale, bitrev_bits((uint32_t)g, ps.log_n), fp), fp);
        }
        if (STRIDED && !DIF) {   // DIT inter-pass twiddle on the way in: w_{2^(lo+r)}^(bitrev(t) * low)
            uint32_t e = bitrev_bits((uint32_t)t, R_LOG) * ((low0 + (uint32_t)c) >> ps.col_bits);
            x = mont_mul(x, pow_lookup(ps.tw, e << tw_shift, fp), fp);
        }
        tile[t * NTT_TS + c] = x;
    }
    __syncthreads();

    run_rounds<R_LOG, DIF>(tile, tws, fp, ps.ncols);

    // ---- store (same addresses) ----
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        int t, c;
        size_t g;
        if (STRIDED) { t = i / NTT_C; c = i % NTT_C; g = gbase + ((size_t)t << ps.lo) + c; }
        else         { c = i >> R_LOG; t = i & (R - 1); g = gbase + i; }
        uint32_t x = tile[t * NTT_TS + c];
        if (STRIDED && DIF) {    // DIF inter-pass twiddle on the way out
            uint32_t e = bitrev_bits((uint32_t)t, R_LOG) * ((low0 + (uint32_t)c) >> ps.col_bits);
            x = mont_mul(x, pow_lookup(ps.tw, e << tw_shift, fp), fp);
        }
        ps.dst[g] = x;
    }
}

template <int R_LOG, bool DIF, bool STRIDED>
static void launch_pass(stark_ctx* ctx, const NttPass& ps, size_t tiles) {
    constexpr int R = 1 << R_LOG;
    int threads = (R * NTT_C) >> 4;
    if (threads < 32) threads = 32;
    if (threads > 1024) threads = 1024;
    size_t smem = (size_t)(R * NTT_TS + (R >> 1) + 2) * sizeof(uint32_t);
    auto kern = ntt_pass_kernel<R_LOG, DIF, STRIDED>;
    if (smem > 48 * 1024)   // opt in per launch: the attribute is per device, contexts may live on several
        STARK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)tiles, threads, smem, ctx->stream>>>(ps, ctx->fp);
    ctx->launches++;
}

template <bool DIF, bool STRIDED>
static void dispatch_pass(stark_ctx* ctx, unsigned r, const NttPass& ps, size_t tiles) {
    switch (r) {
        case 1: if constexpr (!STRIDED) launch_pass<1, DIF, false>(ctx, ps, tiles); break;
        case 2: if constexpr (!STRIDED) launch_pass<2, DIF, false>(ctx, ps, tiles); break;
        case 3: if constexpr (!STRIDED) launch_pass<3, DIF, false>(ctx, ps, tiles); break;
        case 4: if constexpr (!STRIDED) launch_pass<4, DIF, false>(ctx, ps, tiles); break;
        case 5: launch_pass<5, DIF, STRIDED>(ctx, ps, tiles); break;
        case 6: launch_pass<6, DIF, STRIDED>(ctx, ps, tiles); break;
        case 7: launch_pass<7, DIF, STRIDED>(ctx, ps, tiles); break;
        case 8: launch_pass<8, DIF, STRIDED>(ctx, ps, tiles); break;
        case 9: launch_pass<9, DIF, STRIDED>(ctx, ps, tiles); break;
        default: throw StarkError(ST_INTERNAL, "ntt: bad pass width");
    }
}

__global__ void point_scale_kernel(const uint32_t* src, uint32_t* dst, int has_scale, PowTable scale, FieldParams fp) {
    uint32_t x = src[0];
    if (has_scale) x = mont_mul(x, pow_lookup(scale, 0, fp), fp);
    dst[0] = x;
}

// experiment knob (tools/variants_plan.sh): STARK_NTT_PLAN="21:8,7,6;24:8,8,8" overrides the pass widths for those sizes
static bool plan_override(unsigned log_n, std::vector<unsigned>& out) {
    const char* e = getenv("STARK_NTT_PLAN");
    if (!e) return false;
    std::string s(e);
    size_t pos = 0;
    while (pos < s.size()) {
        size_t end = s.find(';', pos);
        if (end == std::string::npos) end = s.size();
        std::string item = s.substr(pos, end - pos);
        size_t colon = item.find(':');
        if (colon != std::string::npos && (unsigned)atoi(item.substr(0, colon).c_str()) == log_n) {
            std::vector<unsigned> v;
            unsigned sum = 0;
            size_t q = colon + 1;
            while (q < item.size()) {
                size_t comma = item.find(',', q);
                if (comma == std::string::npos) comma = item.size();
                unsigned b = (unsigned)atoi(item.substr(q, comma - q).c_str());
                if (b < 5 || b > 9) return false;
                v.push_back(b); sum += b;
                q = comma + 1;
            }
            if (sum == log_n && v.size() >= 2 && v.size() <= 4) { out = v; return true; }
            return false;
        }
        pos = end + 1;
    }
    return false;
}
static std::vector<unsigned> plan_bits(unsigned log_n) {
    if (log_n <= 9) return {log_n};
    { std::vector<unsigned> o; if (plan_override(log_n, o)) return o; }
    unsigned k = (log_n + 8) / 9, base = log_n / k, rem = log_n % k;
    std::vector<unsigned> v;
    for (unsigned i = 0; i < k; i++) v.push_back(base + (i < rem ? 1u : 0u));
    return v;   // every entry is in [5, 9]
}

static void check_size(stark_ctx* ctx, unsigned log_n) {
    STARK_REQUIRE(log_n <= ctx->two_adicity, "ntt: 2^log_n does not divide p-1 (no root of unity of that order)");
    STARK_REQUIRE(log_n <= 30, "ntt: log_n > 30 not supported");
}

// The contiguous pass over `cols` groups of 2^r points: tiles of 32 groups, and one narrower tile for what is left
// (a batch whose group count is not a multiple of 32; a single transform has a power-of-two count).
template <bool DIF>
static void launch_contiguous(stark_ctx* ctx, unsigned r, NttPass ps, size_t cols) {
    const size_t full = cols / NTT_C, rem = cols % NTT_C;
    if (full) {
        ps.ncols = NTT_C;
        dispatch_pass<DIF, false>(ctx, r, ps, full);
    }
    if (rem) {
        STARK_REQUIRE(full == 0 || (ps.log_pad == 0 && !ps.has_scale), "ntt: ragged batches take no padding/scale");
        const size_t done = (full * NTT_C) << r;
        ps.src += done; ps.dst += done;
        ps.ncols = (unsigned)rem;
        dispatch_pass<DIF, false>(ctx, r, ps, 1);
    }
}

void ntt_dit(stark_ctx* ctx, const uint32_t* src, uint32_t* data, unsigned log_n, unsigned log_pad,
             const PowTable* scale, bool inverse_root, size_t batch) {
    check_size(ctx, log_n);
    STARK_REQUIRE(log_pad <= log_n, "ntt: padding larger than the transform");
    STARK_REQUIRE(batch >= 1 && (batch == 1 || (log_pad == 0 && scale == nullptr && log_n >= 1)), "ntt: batched transforms take no padding/scale");
    if (log_n == 0) {   // single point: X[0] = x[0] * scale(0)
        point_scale_kernel<<<1, 1, 0, ctx->stream>>>(src, data, scale != nullptr, scale ? *scale : PowTable{}, ctx->fp);
        ctx->launches++;
        STARK_CUDA(cudaGetLastError());
        return;
    }
    const TwiddleSet& tws = ctx->twiddles(log_n);
    const size_t n = (size_t)1 << log_n;
    std::vector<unsigned> bits = plan_bits(log_n);
    // algorithmic bytes of the whole transform at 8 B per element: read the input once, write the output once
    KernelTimer kt(ctx, stark_ctx::CAT_NTT, (8.0 * (double)(n >> log_pad) + 8.0 * (double)n) * (double)batch);
    unsigned lo = 0;
    for (size_t i = 0; i < bits.size(); i++) {
        unsigned r = bits[i];
        NttPass ps{};
        ps.src = (i == 0) ? src : data;
        ps.dst = data;
        ps.log_n = log_n; ps.lo = lo;
        ps.tw = inverse_root ? tws.inv() : tws.fwd();
        ps.small = inverse_root ? ctx->small_inv.as<uint32_t>() : ctx->small_fwd.as<uint32_t>();
        ps.small_log = ctx->small_log;
        if (i == 0) {
            ps.log_pad = log_pad; ps.log_m = log_n - log_pad;
            ps.has_scale = scale != nullptr;
            if (scale) ps.scale = *scale;
            size_t cols = (n * batch) >> r;           // groups of 2^r contiguous points, across the whole batch
            launch_contiguous<false>(ctx, r, ps, cols);
        } else {
            ps.ncols = NTT_C;
            dispatch_pass<false, true>(ctx, r, ps, batch * n / ((size_t)NTT_C << r));
        }
        lo += r;
    }
    STARK_CUDA(cudaGetLastError());
}

void ntt_dif(stark_ctx* ctx, uint32_t* data, unsigned log_n, bool inverse_root, size_t batch) {
    check_size(ctx, log_n);
    if (log_n == 0) return;
    STARK_REQUIRE(batch >= 1, "ntt: empty batch");
    const TwiddleSet& tws = ctx->twiddles(log_n);
    const size_t n = (size_t)1 << log_n;
    std::vector<unsigned> bits = plan_bits(log_n);
    KernelTimer kt(ctx, stark_ctx::CAT_NTT, 16.0 * (double)n * (double)batch);
    unsigned hi = log_n;
    for (size_t ii = bits.size(); ii-- > 0;) {
        unsigned r = bits[ii];
        unsigned lo = hi - r;
        NttPass ps{};
        ps.src = data; ps.dst = data;
        ps.log_n = log_n; ps.lo = lo;
        ps.tw = inverse_root ? tws.inv() : tws.fwd();
        ps.small = inverse_root ? ctx->small_inv.as<uint32_t>() : ctx->small_fwd.as<uint32_t>();
        ps.small_log = ctx->small_log;
        if (lo == 0) {
            size_t cols = (n * batch) >> r;
            launch_contiguous<true>(ctx, r, ps, cols);
        } else {
            ps.ncols = NTT_C;
            dispatch_pass<true, true>(ctx, r, ps, batch * n / ((size_t)NTT_C << r));
        }
        hi = lo;
    }
    STARK_CUDA(cudaGetLastError());
}

// ---- column-batched DIF: data[2^log_n rows][2^col_bits columns], every column an independent transform -----
// (phase A of the multi-GPU four-step NTT: the columns are a rank's slice of n2, contiguous in memory, so the rows
// it produces can be shipped to their owners as contiguous chunks).  Natural row order in, bit-reversed out.
void ntt_dif_columns(stark_ctx* ctx, uint32_t* data, unsigned log_n, unsigned col_bits, bool inverse_root) {
    check_size(ctx, log_n);
    STARK_REQUIRE(col_bits >= 5 && log_n + col_bits <= 31, "ntt_dif_columns: needs >= 32 columns and < 2^31 elements");
    if (log_n == 0) return;
    const TwiddleSet& tws = ctx->twiddles(log_n);
    const size_t total = (size_t)1 << (log_n + col_bits);
    std::vector<unsigned> bits = plan_bits(log_n);
    KernelTimer kt(ctx, stark_ctx::CAT_NTT, 16.0 * (double)total);
    unsigned hi = log_n;
    for (size_t ii = bits.size(); ii-- > 0;) {
        unsigned r = bits[ii], tlo = hi - r;
        if (r < 5) {   // tiny transforms: widen the pass so the strided kernel exists for it
            throw StarkError(ST_UNSUPPORTED, "ntt_dif_columns: transform smaller than 2^5 rows");
        }
        NttPass ps{};
        ps.src = data; ps.dst = data;
        ps.log_n = log_n; ps.lo = tlo + col_bits; ps.col_bits = col_bits; ps.ncols = NTT_C;
        ps.tw = inverse_root ? tws.inv() : tws.fwd();
        ps.small = inverse_root ? ctx->small_inv.as<uint32_t>() : ctx->small_fwd.as<uint32_t>();
        ps.small_log = ctx->small_log;
        dispatch_pass<true, true>(ctx, r, ps, total / ((size_t)NTT_C << r));
        hi = tlo;
    }
    STARK_CUDA(cudaGetLastError());
}

// ---- natural -> natural transform without a permutation sweep ------------------------------------------------
// (Polynomial::evaluate over a whole domain, ops.rs:76-83 @ fri_commit.rs:78, and Polynomial::interpolate,
// ops.rs:239-241: both take and return natural order.)  The index is split into digits n = [n1][n2]..[nm], n1 the
// most significant; pass i transforms digit i.  Passes 1..m-1 are strided and in place: the element at
// [k1..k(i-1)][n_i][low] is first multiplied by W^(K*n_i), W = w_{2^(log_n-lo_i)}, K = k1 + R1 k2 + .. the output
// index accumulated so far -- a per-ROW scalar of the tile (K is fixed by the tile's high address bits), kept in
// shared memory, so the inter-pass twiddle costs one product per element and no table look-up.  Rows go into the
// tile in bit-reversed order so that the lazy decimation-in-time rounds leave digit k_i in natural row order.
// The last pass transforms the contiguous digit and writes X[K + 2^(log_n-r_m) k_m]: a tile holds 32 consecutive
// k1 (= 32 consecutive K) for every n_m, so both its reads (runs of 2^r_m words) and its writes (128-byte lines)
// are coalesced and the digit reversal costs nothing.  16-byte global accesses throughout.
struct NatPass {
    const uint32_t* src;
    uint32_t* dst;
    unsigned log_n;
    unsigned lo;             // lowest address bit of this pass's digit
    unsigned src_len;        // FIRST: valid entries of src (zero above)
    int has_scale;           // FIRST: x_j *= scale(j);  LAST: X_k *= scale(k)
    PowTable scale;
    PowTable tw;             // w_{2^log_n}^e
    const uint32_t* small;
    unsigned small_log;
    unsigned col_bits;       // column-batched data: the lowest address bits index independent transforms (0 otherwise)
    // FIRST pass of a column-batched transform reading a slice of a larger row-major array (phase A of the four-step
    // NTT reads its columns of the coefficient matrix directly: no staging sweep): element (row, col) of the working array
    // comes from src[(row << src_row_log) + src_col_off + col]; src_len and the input scale are indexed the same way
    int src_map;
    unsigned src_row_log, src_col_off;
    unsigned nprev;          // digits already transformed (address bits above this digit), most significant first
    unsigned prev_bits[4];   // their widths
    unsigned prev_off[4];    // their positions in the output index: off[d] = r_1 + .. + r_(d-1)
};

__device__ __forceinline__ size_t nat_src_index(const NatPass& ps, size_t g) {
    if (!ps.src_map) return g;
    return ((g >> ps.col_bits) << ps.src_row_log) + ps.src_col_off + (g & (((size_t)1 << ps.col_bits) - 1));
}
// output index accumulated by the earlier passes from the address bits above this pass's digit
__device__ __forceinline__ uint32_t nat_kacc(const NatPass& ps, uint32_t high) {
    uint32_t k = 0;
#pragma unroll
    for (int d = 3; d >= 0; d--) {           // static indices: the parameter struct stays in constant memory
        if (d < (int)ps.nprev) {
            k |= (high & ((1u << ps.prev_bits[d]) - 1u)) << ps.prev_off[d];
            high >>= ps.prev_bits[d];
        }
    }
    return k;
}

constexpr int nat_threads(int r_log) { return (2 << r_log) < 64 ? 64 : ((2 << r_log) > 512 ? 512 : (2 << r_log)); }
// resident CTAs per SM the register allocation aims for: STARK_NTT_REGCAP = registers per thread the transforms may use
#ifndef STARK_NTT_REGCAP
#define STARK_NTT_REGCAP 42
#endif
constexpr int ntt_min_blocks(int threads, int cap) {
    int by_regs = 65536 / (STARK_NTT_REGCAP * threads), by_threads = 1536 / threads;
    int b = by_regs < by_threads ? by_regs : by_threads;
    return b > cap ? cap : (b < 1 ? 1 : b);
}
constexpr int nat_min_blocks(int r_log) { return ntt_min_blocks(nat_threads(r_log), 16); }

template <int R_LOG, bool FIRST, bool LAZY, bool PREF>
__global__ void __launch_bounds__(nat_threads(R_LOG), nat_min_blocks(R_LOG)) nat_strided_kernel(NatPass ps, FieldParams fp_arg) {
    const typename FieldOf<PREF>::type fp = FieldOf<PREF>::make(fp_arg);
    extern __shared__ uint32_t smem[];
    constexpr int R = 1 << R_LOG;
    constexpr int TS = NTT_TS16;
    uint32_t* tile = smem;                    // [R][TS], row rho holds digit value bitrev(rho) until the rounds have run
    uint32_t* tws = smem + R * TS;        // [R/2] twiddles of w_R
    uint32_t* rowtw = tws + (R >> 1);         // [R] W^(K t), t = bitrev(row)
    fill_tile_twiddles(tws, ps.small, ps.small_log, R_LOG);

    const size_t tiles_per_high = ((size_t)1 << ps.lo) / NTT_C;
    const size_t high = blockIdx.x / tiles_per_high;
    const uint32_t low0 = (uint32_t)(blockIdx.x % tiles_per_high) * NTT_C;
    const size_t gbase = (high << (ps.lo + R_LOG)) | low0;
    // ---- load: a warp covers 4 tile rows x 128 bytes.  A thread's global loads are all issued before the row
    // twiddles are looked up and before the first barrier, so that table and data latencies overlap. ----
    constexpr int T = nat_threads(R_LOG), ITER = (R * 8) / T;     // 4 (8 for 2^9 rows)
    static_assert(ITER % 4 == 0, "tile rows x 8 vectors must be a multiple of 4 per thread");
#pragma unroll 1
    for (int b = 0; b < ITER; b += 4) {
        uint4 a[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = threadIdx.x + (b + u) * T;
            const int g4 = i & 7, rho = i >> 3;
            const size_t g = gbase + ((size_t)bitrev_bits((uint32_t)rho, R_LOG) << ps.lo) + 4 * g4;
            if (FIRST) {
                const size_t gs = nat_src_index(ps, g);
                if (gs + 3 < ps.src_len) a[u] = *reinterpret_cast<const uint4*>(ps.src + gs);
                else {
                    a[u].x = (gs < ps.src_len) ? ps.src[gs] : 0u;
                    a[u].y = (gs + 1 < ps.src_len) ? ps.src[gs + 1] : 0u;
                    a[u].z = (gs + 2 < ps.src_len) ? ps.src[gs + 2] : 0u;
                    a[u].w = 0u;
                }
            } else {
                a[u] = *reinterpret_cast<const uint4*>(ps.dst + g);
            }
        }
        if (!FIRST && b == 0) {
            const uint32_t K = nat_kacc(ps, (uint32_t)high);
            for (int rho = threadIdx.x; rho < R; rho += T) {         // indexed by tile row: consecutive words for a warp's 4 rows
                rowtw[rho] = pow_lookup(ps.tw, (K * bitrev_bits((uint32_t)rho, R_LOG)) << (ps.lo - ps.col_bits), fp);
            }
            __syncthreads();
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = threadIdx.x + (b + u) * T;
            const int g4 = i & 7, rho = i >> 3;
            uint32_t v[4] = {a[u].x, a[u].y, a[u].z, a[u].w};
            if (FIRST) {
                const size_t g = nat_src_index(ps, gbase + ((size_t)bitrev_bits((uint32_t)rho, R_LOG) << ps.lo) + 4 * g4);
                if (ps.has_scale && g < ps.src_len) {      // one look-up, then walk base^(g+k) (the table ends at src_len)
                    uint32_t sc = pow_lookup(ps.scale, (uint32_t)g, fp);
                    const uint32_t step = __ldg(ps.scale.lo + 1);
#pragma unroll
                    for (int k = 0; k < 4; k++) { v[k] = mont_mul(v[k], sc, fp); if (k < 3) sc = mont_mul(sc, step, fp); }
                }
            } else {
                const uint32_t w = rowtw[rho];
#pragma unroll
                for (int k = 0; k < 4; k++) v[k] = mont_mul_tw(v[k], w, fp);
            }
            *reinterpret_cast<uint4*>(tile + rho * TS + 4 * g4) = make_uint4(v[0], v[1], v[2], v[3]);
        }
    }
    __syncthreads();

    run_rounds<R_LOG, false, LAZY, TS>(tile, tws, fp, NTT_C);

    // ---- store: digit k_i in natural row order, same addresses (weak values are fine: the next pass multiplies) ----
    for (int i = threadIdx.x; i < R * 8; i += blockDim.x) {
        const int g4 = i & 7, t = i >> 3;
        *reinterpret_cast<uint4*>(ps.dst + gbase + ((size_t)t << ps.lo) + 4 * g4) = *reinterpret_cast<const uint4*>(tile + t * TS + 4 * g4);
    }
}

template <int R_LOG, bool PREF>
__global__ void __launch_bounds__(nat_threads(R_LOG), nat_min_blocks(R_LOG)) nat_last_kernel(NatPass ps, FieldParams fp_arg) {
    const typename FieldOf<PREF>::type fp = FieldOf<PREF>::make(fp_arg);
    extern __shared__ uint32_t smem[];
    constexpr int R = 1 << R_LOG;
    uint32_t* tile = smem;                    // [R rows = n_m][33: column j = k1 - k1_0]
    uint32_t* tws = smem + R * NTT_TS;
    uint32_t* coltw = tws + (R >> 1);         // [32] W^K(j), W = w_{2^log_n}
    fill_tile_twiddles(tws, ps.small, ps.small_log, R_LOG);

    // address bits above the digit: [k1][mid]; the tile takes 32 consecutive k1 at one mid
    // (a batch of transforms: the tiles of one transform are consecutive, src / dst move on by 2^log_n per transform)
    const unsigned high_bits = ps.log_n - R_LOG, mid_bits = high_bits - ps.prev_bits[0];
    const uint32_t tile_in = blockIdx.x & ((1u << (high_bits - 5)) - 1u);
    {
        const size_t shift = (size_t)(blockIdx.x >> (high_bits - 5)) << ps.log_n;
        ps.src += shift; ps.dst += shift;
    }
    const uint32_t mid = tile_in & ((1u << mid_bits) - 1u);
    const uint32_t k1_0 = (tile_in >> mid_bits) * NTT_C;
    const uint32_t K0 = nat_kacc(ps, (k1_0 << mid_bits) | mid);     // K(j) = K0 + j  (k1 is the lowest digit of K)
    // ---- load: 8 lanes read 128 contiguous bytes of one row; a warp covers 4 values of j.  Data and table loads of
    // four vectors are in flight together, ahead of the column-twiddle look-up and the first barrier. ----
    constexpr int T = nat_threads(R_LOG), ITER = (R * 8) / T;
    static_assert(ITER % 4 == 0, "tile rows x 8 vectors must be a multiple of 4 per thread");
#pragma unroll 1
    for (int b = 0; b < ITER; b += 4) {
        uint4 a[4];
        uint32_t wl[4], wh[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = threadIdx.x + (b + u) * T;
            const int q8 = i & 7, jj = (i >> 3) & 3, rest = i >> 5;
            const int seg = rest & ((R >> 5) - 1), j = (rest >> (R_LOG - 5)) * 4 + jj;
            const uint32_t n3 = (uint32_t)seg * 32 + 4 * q8;
            const size_t row = ((size_t)(k1_0 + j) << mid_bits) | mid;
            a[u] = *reinterpret_cast<const uint4*>(ps.src + (row << R_LOG) + n3);
            const uint32_t e = n3 * (K0 + (uint32_t)j);                        // W^(n_m K)
            wl[u] = __ldg(ps.tw.lo + (e & ps.tw.mask));
            wh[u] = __ldg(ps.tw.hi + (e >> ps.tw.shift));
        }
        if (b == 0) {
            if (threadIdx.x < NTT_C) {
                coltw[threadIdx.x] = pow_lookup(ps.tw, K0 + threadIdx.x, fp);
            }
            __syncthreads();
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = threadIdx.x + (b + u) * T;
            const int q8 = i & 7, jj = (i >> 3) & 3, rest = i >> 5;
            const int seg = rest & ((R >> 5) - 1), j = (rest >> (R_LOG - 5)) * 4 + jj;
            const uint32_t n3 = (uint32_t)seg * 32 + 4 * q8;
            const uint32_t step = coltw[j];
            uint32_t w = mont_mul(wl[u], wh[u], fp);
            uint32_t* o = tile + n3 * NTT_TS + j;
            o[0] = mont_mul(a[u].x, w, fp); w = mont_mul_tw(w, step, fp);
            o[NTT_TS] = mont_mul(a[u].y, w, fp); w = mont_mul_tw(w, step, fp);
            o[2 * NTT_TS] = mont_mul(a[u].z, w, fp); w = mont_mul_tw(w, step, fp);
            o[3 * NTT_TS] = mont_mul(a[u].w, w, fp);
        }
    }
    __syncthreads();

    run_rounds<R_LOG, true>(tile, tws, fp, NTT_C);          // natural rows in, k_m = bitrev(row) out, canonical

    // ---- store: X[K0 + j + 2^(log_n - r) k_m], four j per thread ----
    for (int i = threadIdx.x; i < R * 8; i += blockDim.x) {
        const int g4 = i & 7, rho = i >> 3;
        const uint32_t km = bitrev_bits((uint32_t)rho, R_LOG);
        const uint32_t* o = tile + rho * NTT_TS + 4 * g4;
        const size_t k = ((size_t)km << high_bits) + K0 + 4 * g4;
        uint32_t v[4] = {o[0], o[1], o[2], o[3]};
        if (ps.has_scale) {
            uint32_t sc = pow_lookup(ps.scale, (uint32_t)k, fp);
            const uint32_t step = __ldg(ps.scale.lo + 1);
#pragma unroll
            for (int c = 0; c < 4; c++) { v[c] = mont_mul(v[c], sc, fp); if (c < 3) sc = mont_mul(sc, step, fp); }
        }
        *reinterpret_cast<uint4*>(ps.dst + k) = make_uint4(v[0], v[1], v[2], v[3]);
    }
}

template <int R_LOG, bool FIRST>
static void launch_nat_strided(stark_ctx* ctx, const NatPass& ps, size_t tiles) {
    constexpr int R = 1 << R_LOG;
    const int threads = nat_threads(R_LOG);
    size_t smem = (size_t)(R * NTT_TS16 + (R >> 1) + R) * sizeof(uint32_t);
    auto launch = [&](auto kern) {
        if (smem > 48 * 1024) STARK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)tiles, threads, smem, ctx->stream>>>(ps, ctx->fp);
    };
    if (is_ref_field(ctx)) launch(nat_strided_kernel<R_LOG, FIRST, true, true>);
    else if (ctx->fp.p >> 31) launch(nat_strided_kernel<R_LOG, FIRST, true, false>);     // weak butterfly values need 2^32 < 2p
    else launch(nat_strided_kernel<R_LOG, FIRST, false, false>);
    ctx->launches++;
}
template <int R_LOG>
static void launch_nat_last(stark_ctx* ctx, const NatPass& ps, size_t tiles) {
    constexpr int R = 1 << R_LOG;
    const int threads = nat_threads(R_LOG);
    size_t smem = (size_t)(R * NTT_TS + (R >> 1) + NTT_C) * sizeof(uint32_t);
    auto kern = is_ref_field(ctx) ? nat_last_kernel<R_LOG, true> : nat_last_kernel<R_LOG, false>;
    if (smem > 48 * 1024) STARK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)tiles, threads, smem, ctx->stream>>>(ps, ctx->fp);
    ctx->launches++;
}

bool ntt_natural_supported(unsigned log_n, const void* src, const void* work, const void* dst) {
    auto aligned = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    return log_n >= 10 && log_n <= 30 && aligned(src) && aligned(work) && aligned(dst);
}

void ntt_natural(stark_ctx* ctx, const uint32_t* src, size_t src_len, uint32_t* work, uint32_t* dst, unsigned log_n,
                 bool inverse_root, const PowTable* in_scale, const PowTable* out_scale, size_t batch) {
    check_size(ctx, log_n);
    STARK_REQUIRE(ntt_natural_supported(log_n, src, work, dst), "ntt_natural: size or alignment not supported");
    const TwiddleSet& tws = ctx->twiddles(log_n);
    const size_t n = (size_t)1 << log_n;
    STARK_REQUIRE(batch >= 1 && (batch == 1 || (src_len == n * batch && in_scale == nullptr)) && n * batch <= ((size_t)1 << 31),
                  "ntt_natural: a batch takes full-length inputs, no input scale and < 2^31 elements");
    if (batch == 1 && src_len > n) src_len = n;
    std::vector<unsigned> bits = plan_bits(log_n);         // >= 2 digits of 5..9 bits, most significant first
    KernelTimer kt(ctx, stark_ctx::CAT_NTT, 8.0 * (double)src_len + 8.0 * (double)n * (double)batch);
    NatPass ps{};
    ps.log_n = log_n;
    ps.tw = inverse_root ? tws.inv() : tws.fwd();
    ps.small = inverse_root ? ctx->small_inv.as<uint32_t>() : ctx->small_fwd.as<uint32_t>();
    ps.small_log = ctx->small_log;
    unsigned hi = log_n, off = 0;
    for (size_t i = 0; i < bits.size(); i++) {
        const unsigned r = bits[i];
        const bool first = (i == 0), last = (i + 1 == bits.size());
        ps.lo = hi - r;
        ps.nprev = (unsigned)i;
        ps.src = first ? src : work;
        ps.dst = last ? dst : work;
        ps.src_len = (unsigned)src_len;
        ps.has_scale = first ? (in_scale != nullptr) : (last ? (out_scale != nullptr) : 0);
        if (first && in_scale) ps.scale = *in_scale;
        if (last && out_scale) ps.scale = *out_scale;
        const size_t tiles = batch * (n / ((size_t)NTT_C << r));      // strided passes: the batch index is just more high address bits
        if (!last) {
            switch (r) {
                case 5: first ? launch_nat_strided<5, true>(ctx, ps, tiles) : launch_nat_strided<5, false>(ctx, ps, tiles); break;
                case 6: first ? launch_nat_strided<6, true>(ctx, ps, tiles) : launch_nat_strided<6, false>(ctx, ps, tiles); break;
                case 7: first ? launch_nat_strided<7, true>(ctx, ps, tiles) : launch_nat_strided<7, false>(ctx, ps, tiles); break;
                case 8: first ? launch_nat_strided<8, true>(ctx, ps, tiles) : launch_nat_strided<8, false>(ctx, ps, tiles); break;
                case 9: first ? launch_nat_strided<9, true>(ctx, ps, tiles) : launch_nat_strided<9, false>(ctx, ps, tiles); break;
                default: throw StarkError(ST_INTERNAL, "ntt_natural: bad pass width");
            }
        } else {
            switch (r) {
                case 5: launch_nat_last<5>(ctx, ps, tiles); break;
                case 6: launch_nat_last<6>(ctx, ps, tiles); break;
                case 7: launch_nat_last<7>(ctx, ps, tiles); break;
                case 8: launch_nat_last<8>(ctx, ps, tiles); break;
                case 9: launch_nat_last<9>(ctx, ps, tiles); break;
                default: throw StarkError(ST_INTERNAL, "ntt_natural: bad pass width");
            }
        }
        ps.prev_bits[i] = r; ps.prev_off[i] = off;
        off += r; hi -= r;
    }
    STARK_CUDA(cudaGetLastError());
}

// ---- column-batched transform with the same strided passes: data[2^log_n rows][2^col_bits columns], every column an
// independent transform (phase A of the multi-GPU four-step NTT).  Every digit is strided here, so all passes are
// nat_strided passes, in place; natural rows in, and slot [k_1][k_2]..[k_m] (digit k_1 in the most significant address
// bits) holds X[k_1 + R_1 k_2 + R_1 R_2 k_3 + ..]: digit-reversed, which the caller undoes for free when it scatters the
// rows to their owners.  Values are weak after the last pass (the consumer multiplies).  Returns the digit widths.
std::vector<unsigned> ntt_columns_digitrev(stark_ctx* ctx, uint32_t* data, unsigned log_n, unsigned col_bits, bool inverse_root,
                                           const ColumnSource* from) {
    check_size(ctx, log_n);
    STARK_REQUIRE(log_n >= 10 && col_bits >= 5 && log_n + col_bits <= 31, "ntt_columns_digitrev: needs >= 2^10 rows, >= 32 columns, < 2^31 elements");
    STARK_REQUIRE((reinterpret_cast<uintptr_t>(data) & 15) == 0, "ntt_columns_digitrev: unaligned buffer");
    const TwiddleSet& tws = ctx->twiddles(log_n);
    const size_t total = (size_t)1 << (log_n + col_bits);
    std::vector<unsigned> bits = plan_bits(log_n);
    KernelTimer kt(ctx, stark_ctx::CAT_NTT, 16.0 * (double)total);
    NatPass ps{};
    ps.log_n = log_n; ps.col_bits = col_bits;
    ps.tw = inverse_root ? tws.inv() : tws.fwd();
    ps.small = inverse_root ? ctx->small_inv.as<uint32_t>() : ctx->small_fwd.as<uint32_t>();
    ps.small_log = ctx->small_log;
    ps.src = data; ps.dst = data; ps.src_len = (unsigned)total; ps.has_scale = 0;
    if (from) {           // the first pass gathers the working array from a slice of `from->src` and scales it on the way in
        STARK_REQUIRE((reinterpret_cast<uintptr_t>(from->src) & 15) == 0 && from->col_off % 4 == 0 && from->src_len <= 0xffffffffull,
                      "ntt_columns_digitrev: unaligned source slice");
        ps.src = from->src; ps.src_len = (unsigned)from->src_len; ps.src_map = 1;
        ps.src_row_log = from->row_log; ps.src_col_off = from->col_off;
        ps.has_scale = from->scale != nullptr;
        if (from->scale) ps.scale = *from->scale;
    }
    unsigned hi = log_n, off = 0;
    for (size_t i = 0; i < bits.size(); i++) {
        const unsigned r = bits[i];
        ps.lo = hi - r + col_bits;
        ps.nprev = (unsigned)i;
        const size_t tiles = total / ((size_t)NTT_C << r);
        const bool first = (i == 0);
        switch (r) {
            case 5: first ? launch_nat_strided<5, true>(ctx, ps, tiles) : launch_nat_strided<5, false>(ctx, ps, tiles); break;
            case 6: first ? launch_nat_strided<6, true>(ctx, ps, tiles) : launch_nat_strided<6, false>(ctx, ps, tiles); break;
            case 7: first ? launch_nat_strided<7, true>(ctx, ps, tiles) : launch_nat_strided<7, false>(ctx, ps, tiles); break;
            case 8: first ? launch_nat_strided<8, true>(ctx, ps, tiles) : launch_nat_strided<8, false>(ctx, ps, tiles); break;
            case 9: first ? launch_nat_strided<9, true>(ctx, ps, tiles) : launch_nat_strided<9, false>(ctx, ps, tiles); break;
            default: throw StarkError(ST_INTERNAL, "ntt_columns_digitrev: bad pass width");
        }
        ps.prev_bits[i] = r; ps.prev_off[i] = off;
        off += r; hi -= r;
    }
    STARK_CUDA(cudaGetLastError());
    return bits;
}

// ---- blow-up-by-8 forward transform: one size-n NTT over 8 interleaved coset columns ------------------------
// X[8k'+s] = sum_j c_j (g w_N^s)^j w_n^(j k'): the evaluation on the size-N = 8n coset is eight size-n
// transforms (one per shift s) that share every butterfly twiddle and every inter-pass twiddle, and whose
// outputs interleave into natural order.  Stored as D[n rows][8 columns] the eight columns of a row are 32
// contiguous bytes: one twiddle is formed per ROW (not per element), global accesses are two 16-byte vectors
// per row, and four adjacent rows make the 128-byte line.  Tile = 2^r rows x (4 row-groups x 8 columns).
struct Lde8Pass {
    const uint32_t* src;     // FIRST pass: coefficient source; later passes: unused (in place on dst)
    uint32_t* dst;           // [n][8]
    unsigned log_rows;       // n = 2^log_rows
    unsigned lo;             // lowest row-index bit of this pass
    unsigned src_len;        // FIRST: valid entries of src
    int src_bitrev;          // FIRST: 1 = src is in bit-reversed row order (output of the DIF inverse), 0 = natural coefficients
    int last;                // canonicalise on store
    PowTable scale;          // FIRST: c0 * g^j
    PowTable shift;          // FIRST: w_N^j  (N = 8n)
    PowTable tw;             // later passes: w_n^e
    const uint32_t* small;
    unsigned small_log;
};

// threads per tile of 2^r rows x 4 row-groups: 2R >> STARK_LDE8_TSHIFT (experiment knob: smaller CTAs = more
// independent load / butterfly / store phases in flight per SM and a shorter last wave)
#ifndef STARK_LDE8_TSHIFT
#define STARK_LDE8_TSHIFT 1
#endif
constexpr int lde8_threads(int r_log) {
    int t = (2 << r_log) >> STARK_LDE8_TSHIFT;
    return t < 64 ? 64 : (t > 1024 ? 1024 : t);
}

template <int R_LOG, bool FIRST, bool LAZY, bool PREF>
__global__ void __launch_bounds__(lde8_threads(R_LOG), ntt_min_blocks(lde8_threads(R_LOG), 24)) lde8_pass_kernel(Lde8Pass ps, FieldParams fp_arg) {
    const typename FieldOf<PREF>::type fp = FieldOf<PREF>::make(fp_arg);
    extern __shared__ uint32_t smem[];
    constexpr int R = 1 << R_LOG;
    constexpr int TS = NTT_TS16;
    uint32_t* tile = smem;                    // [R][TS]: word g*8+s of row t
    uint32_t* tws = smem + R * TS;
    fill_tile_twiddles(tws, ps.small, ps.small_log, R_LOG);

    const size_t tile_id = blockIdx.x;
    size_t row_base;          // FIRST: first row of the tile; else row of (t = 0, g = 0)
    uint32_t low0 = 0;
    if (FIRST) {
        row_base = tile_id * (size_t)R;          // rows row_base + t + (g << (log_rows - 2))
    } else {
        const size_t tiles_per_high = ((size_t)1 << ps.lo) / 4;
        const size_t high = tile_id / tiles_per_high;
        low0 = (uint32_t)(tile_id % tiles_per_high) * 4;
        row_base = (high << (ps.lo + R_LOG)) | low0;
    }
    const unsigned tw_shift = ps.log_rows - ps.lo - R_LOG;

    // ---- load: one work item = one row (8 columns) ----
    if (!FIRST) {
        // two rows of a thread at a time: data and twiddle-table loads in flight together
        constexpr int T = lde8_threads(R_LOG);
        static_assert((4 * R) % (2 * T) == 0, "rows per thread must be even");
#pragma unroll 1
        for (int b = 0; b < 4 * R; b += 2 * T) {
        uint4 a[2], bq[2];
        uint32_t wl[2], wh[2];
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int i = b + threadIdx.x + u * T;
            const int g = i & 3, t = i >> 2;
            const uint4* p = reinterpret_cast<const uint4*>(ps.dst + (row_base + ((size_t)t << ps.lo) + g) * 8);
            a[u] = p[0]; bq[u] = p[1];
            const uint32_t e = (bitrev_bits((uint32_t)t, R_LOG) * (low0 + (uint32_t)g)) << tw_shift;
            wl[u] = __ldg(ps.tw.lo + (e & ps.tw.mask));
            wh[u] = __ldg(ps.tw.hi + (e >> ps.tw.shift));
        }
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int i = b + threadIdx.x + u * T;
            const int g = i & 3, t = i >> 2;
            const uint32_t tw = mont_mul(wl[u], wh[u], fp);
            uint4* o = reinterpret_cast<uint4*>(tile + t * TS + g * 8);
            o[0] = make_uint4(mont_mul_tw(a[u].x, tw, fp), mont_mul_tw(a[u].y, tw, fp), mont_mul_tw(a[u].z, tw, fp), mont_mul_tw(a[u].w, tw, fp));
            o[1] = make_uint4(mont_mul_tw(bq[u].x, tw, fp), mont_mul_tw(bq[u].y, tw, fp), mont_mul_tw(bq[u].z, tw, fp), mont_mul_tw(bq[u].w, tw, fp));
        }
        }
    }
    // FIRST: the tile holds all 2^r values of the lowest row digit t for four values g of the TOP two row bits, i.e. rows
    // q = t | tile << r | g << (log_rows - 2).  Row q holds coefficient bitrev(q), so the four rows of one t are four
    // ADJACENT coefficients (one aligned 16-byte load of natural-order input; the 32-byte sectors of the gather are used
    // in full by two neighbouring tiles instead of one word in eight), and their scale / shift factors are one table
    // look-up each plus a walk.  The rows of one g are 2^r consecutive rows of the output: stores stay contiguous.
    if (FIRST) {
        const unsigned top = ps.log_rows - 2, mid_bits = top - R_LOG;
        const uint32_t jmid = bitrev_bits((uint32_t)tile_id, mid_bits) << 2;
        const bool vec = !ps.src_bitrev && (reinterpret_cast<uintptr_t>(ps.src) & 15) == 0;
        for (int t = threadIdx.x; t < R; t += blockDim.x) {
            const uint32_t j0 = (bitrev_bits((uint32_t)t, R_LOG) << (ps.log_rows - R_LOG)) | jmid;     // coefficients j0 .. j0 + 3
            const uint32_t q0 = (uint32_t)t | ((uint32_t)tile_id << R_LOG);
            uint32_t c[4];
            if (vec && j0 + 3 < ps.src_len) {
                const uint4 x = *reinterpret_cast<const uint4*>(ps.src + j0);
                c[0] = x.x; c[1] = x.y; c[2] = x.z; c[3] = x.w;
            } else {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t si = ps.src_bitrev ? (q0 | (bitrev_bits((uint32_t)k, 2) << top)) : j0 + k;
                    c[k] = si < ps.src_len ? ps.src[si] : 0u;
                }
            }
            uint32_t sc = pow_lookup(ps.scale, j0, fp), b = pow_lookup(ps.shift, j0, fp);      // c0 g^j0, w_N^j0
            const uint32_t sc_step = __ldg(ps.scale.lo + 1), b_step = __ldg(ps.shift.lo + 1);  // g, w_N
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int g = (int)bitrev_bits((uint32_t)k, 2);
                uint32_t v[8];
                v[0] = mont_mul(c[k], sc, fp);                                                 // c_j g^j
#pragma unroll
                for (int e = 1; e < 8; e++) v[e] = mont_mul_tw(v[e - 1], b, fp);               // ... * w_N^(j s)
                uint4* o = reinterpret_cast<uint4*>(tile + t * TS + g * 8);
                o[0] = make_uint4(v[0], v[1], v[2], v[3]);
                o[1] = make_uint4(v[4], v[5], v[6], v[7]);
                if (k < 3) { sc = mont_mul(sc, sc_step, fp); b = mont_mul(b, b_step, fp); }
            }
        }
    }
    __syncthreads();

    run_rounds<R_LOG, false, LAZY, TS>(tile, tws, fp, NTT_C);

    // ---- store ----
    for (int i = threadIdx.x; i < 4 * R; i += blockDim.x) {
        int t, g;
        size_t row;
        if (FIRST) { g = i >> R_LOG; t = i & (R - 1); row = row_base + t + ((size_t)g << (ps.log_rows - 2)); }
        else { g = i & 3; t = i >> 2; row = row_base + ((size_t)t << ps.lo) + g; }
        const uint4* o4 = reinterpret_cast<const uint4*>(tile + t * TS + g * 8);
        const uint4 lo4 = o4[0], hi4 = o4[1];
        const uint32_t o[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
        uint32_t v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = (LAZY && ps.last) ? canonical(o[k], fp) : o[k];
        uint4* p = reinterpret_cast<uint4*>(ps.dst + row * 8);
        p[0] = make_uint4(v[0], v[1], v[2], v[3]);
        p[1] = make_uint4(v[4], v[5], v[6], v[7]);
    }
}

template <int R_LOG, bool FIRST, bool LAZY, bool PREF>
static void launch_lde8_impl(stark_ctx* ctx, const Lde8Pass& ps, size_t tiles) {
    constexpr int R = 1 << R_LOG;
    const int threads = lde8_threads(R_LOG);
    size_t smem = (size_t)(R * NTT_TS16 + (R >> 1) + 2) * sizeof(uint32_t);
    auto kern = lde8_pass_kernel<R_LOG, FIRST, LAZY, PREF>;
    if (smem > 48 * 1024) STARK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)tiles, threads, smem, ctx->stream>>>(ps, ctx->fp);
    ctx->launches++;
}
// weak (lazily reduced) butterfly values need 2^32 < 2p; smaller primes (BabyBear, 998244353, ...) take the strict form
template <int R_LOG, bool FIRST>
static void launch_lde8(stark_ctx* ctx, const Lde8Pass& ps, size_t tiles) {
    if (is_ref_field(ctx)) launch_lde8_impl<R_LOG, FIRST, true, true>(ctx, ps, tiles);
    else if (ctx->fp.p >> 31) launch_lde8_impl<R_LOG, FIRST, true, false>(ctx, ps, tiles);
    else launch_lde8_impl<R_LOG, FIRST, false, false>(ctx, ps, tiles);
}
template <bool FIRST>
static void dispatch_lde8(stark_ctx* ctx, unsigned r, const Lde8Pass& ps, size_t tiles) {
    switch (r) {
        case 5: launch_lde8<5, FIRST>(ctx, ps, tiles); break;
        case 6: launch_lde8<6, FIRST>(ctx, ps, tiles); break;
        case 7: launch_lde8<7, FIRST>(ctx, ps, tiles); break;
        case 8: launch_lde8<8, FIRST>(ctx, ps, tiles); break;
        case 9: launch_lde8<9, FIRST>(ctx, ps, tiles); break;
        default: throw StarkError(ST_INTERNAL, "lde8: bad pass width");
    }
}

bool lde8_supported(unsigned log_rows) { return log_rows >= 10; }

// dst[8k'+s] = sum_j c_j (base w_N^s)^j w_n^(j k'), c_j = c0 * src[...]:
//   src_bitrev = false: src holds natural-order coefficients (src_len of them, zero above)
//   src_bitrev = true : src holds n values in bit-reversed coefficient order (the DIF inverse's output)
void lde8_forward(stark_ctx* ctx, const uint32_t* src, size_t src_len, bool src_bitrev, uint32_t* dst, unsigned log_rows,
                  uint64_t base, uint64_t c0) {
    STARK_REQUIRE(lde8_supported(log_rows), "lde8: too small");
    const unsigned log_N = log_rows + 3;
    check_size(ctx, log_N);
    const size_t n = (size_t)1 << log_rows;
    const TwiddleSet& tw_rows = ctx->twiddles(log_rows);
    const TwiddleSet& tw_N = ctx->twiddles(log_N);
    ScaleTable st;
    build_scale_table(ctx, base % ctx->modulus, c0 % ctx->modulus, log_rows, st);
    std::vector<unsigned> bits = plan_bits(log_rows);       // every entry in [5, 9] for log_rows >= 10
    KernelTimer kt(ctx, stark_ctx::CAT_NTT, 8.0 * (double)n + 8.0 * 8.0 * (double)n);
    unsigned lo = 0;
    for (size_t i = 0; i < bits.size(); i++) {
        unsigned r = bits[i];
        Lde8Pass ps{};
        ps.src = src; ps.dst = dst; ps.log_rows = log_rows; ps.lo = lo;
        ps.src_len = (unsigned)(src_len < n ? src_len : n); ps.src_bitrev = src_bitrev;
        ps.last = (i + 1 == bits.size());
        ps.scale = st.view; ps.shift = tw_N.fwd(); ps.tw = tw_rows.fwd();
        ps.small = ctx->small_fwd.as<uint32_t>(); ps.small_log = ctx->small_log;
        size_t tiles = n / ((size_t)4 << r);
        if (i == 0) dispatch_lde8<true>(ctx, r, ps, tiles);
        else dispatch_lde8<false>(ctx, r, ps, tiles);
        lo += r;
    }
    STARK_CUDA(cudaGetLastError());
}

// ---- bit-reversal gather (natural <-> bit-reversed), optional scale ------------------------------
__global__ void bitrev_kernel(const uint32_t* in, uint32_t* out, unsigned log_n, size_t total, int has_scale, PowTable scale,
                              int by_input, FieldParams fp) {
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    uint32_t i = (uint32_t)(g & (((size_t)1 << log_n) - 1));     // index inside its transform
    size_t base = g - i;
    uint32_t j = bitrev_bits(i, log_n);
    uint32_t x = in[base + j];
    if (has_scale) x = mont_mul(x, pow_lookup(scale, by_input ? j : i, fp), fp);
    out[g] = x;
}
void bitrev_permute(stark_ctx* ctx, const uint32_t* in, uint32_t* out, unsigned log_n, const PowTable* scale,
                    bool scale_by_input_index, size_t batch) {
    size_t n = ((size_t)1 << log_n) * batch;
    PowTable sc{};
    if (scale) sc = *scale;
    KernelTimer kt(ctx, stark_ctx::CAT_NTT, 16.0 * (double)n);
    bitrev_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(in, out, log_n, n, scale != nullptr, sc,
                                                                         scale_by_input_index, ctx->fp);
    ctx->launches++;
    STARK_CUDA(cudaGetLastError());
}

// ---- per-call power tables (coset offsets, 1/n) --------------------------------------------------
__global__ void scale_table_kernel(uint32_t base_m, uint32_t c0_m, unsigned shift, unsigned n_lo, unsigned n_hi,
                                   uint32_t* lo, uint32_t* hi, FieldParams fp) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_lo) lo[i] = mont_pow(base_m, i, fp);
    else if (i < n_lo + n_hi) {
        unsigned j = i - n_lo;
        hi[j] = mont_mul(c0_m, mont_pow(base_m, (uint64_t)j << shift, fp), fp);
    }
}
void build_scale_table(stark_ctx* ctx, uint64_t base, uint64_t c0, unsigned log_n, ScaleTable& out) {
    unsigned shift = (log_n + 1) / 2;
    unsigned n_lo = 1u << shift, n_hi = 1u << (log_n - shift);
    out.lo = DevBuf(n_lo * sizeof(uint32_t), ctx->stream);
    out.hi = DevBuf(n_hi * sizeof(uint32_t), ctx->stream);
    out.view = PowTable{out.lo.as<uint32_t>(), out.hi.as<uint32_t>(), shift, n_lo - 1};
    unsigned total = n_lo + n_hi;
    KernelTimer kt(ctx, stark_ctx::CAT_OTHER, 0);
    scale_table_kernel<<<(total + 127) / 128, 128, 0, ctx->stream>>>(ctx->to_mont(base), ctx->to_mont(c0), shift, n_lo, n_hi,
                                                                     out.lo.as<uint32_t>(), out.hi.as<uint32_t>(), ctx->fp);
    ctx->launches++;
    STARK_CUDA(cudaGetLastError());
}

}  // namespace starkb200

// ---- twiddle cache (host builds 2*2^(log_n/2) values once per size) -------------------------------
const starkb200::TwiddleSet& stark_ctx::twiddles(unsigned log_n) {
    using namespace starkb200;
    auto it = tw.find(log_n);
    if (it != tw.end()) return *it->second;
    STARK_REQUIRE(log_n <= two_adicity, "twiddles: 2^log_n does not divide p-1");
    auto ts = std::make_unique<TwiddleSet>();
    unsigned shift = (log_n + 1) / 2;
    size_t n_lo = (size_t)1 << shift, n_hi = (size_t)1 << (log_n - shift);
    ts->shift = shift; ts->mask = (uint32_t)(n_lo - 1);
    uint64_t w = root_of_unity(log_n), wi = h_inv(w, modulus);
    auto fill = [&](uint64_t base, DevBuf& lo, DevBuf& hi) {
        std::vector<uint32_t> a(n_lo), b(n_hi);
        uint64_t acc = 1;
        for (size_t j = 0; j < n_lo; j++) { a[j] = to_mont(acc); acc = h_mul(acc, base, modulus); }
        uint64_t step = h_pow(base, n_lo, modulus); acc = 1;
        for (size_t j = 0; j < n_hi; j++) { b[j] = to_mont(acc); acc = h_mul(acc, step, modulus); }
        lo = DevBuf(n_lo * 4, stream); hi = DevBuf(n_hi * 4, stream);
        STARK_CUDA(cudaMemcpyAsync(lo.p, a.data(), n_lo * 4, cudaMemcpyHostToDevice, stream));
        STARK_CUDA(cudaMemcpyAsync(hi.p, b.data(), n_hi * 4, cudaMemcpyHostToDevice, stream));
        STARK_CUDA(cudaStreamSynchronize(stream));   // a, b die here
    };
    fill(w, ts->fwd_lo, ts->fwd_hi);
    fill(wi, ts->inv_lo, ts->inv_hi);
    auto& ref = *ts;
    tw[log_n] = std::move(ts);
    return ref;
}
