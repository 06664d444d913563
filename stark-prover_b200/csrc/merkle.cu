// merkle.cu — SHA-256 Merkle commitment kernels (integer-pipe bound).
//
// Replaces MerkleTree::<M>::new / root (reference src/merkle/mod.rs:10-26, which delegates to
// rs_merkle 1.4.2 `MerkleTree::<Sha256>::from_leaves` / `root_hex`): leaf = SHA-256(BE8(value)),
// parent = SHA-256(left || right), a node without a right sibling is promoted unchanged.
//
// Layout in HBM.  Leaf VALUES stay where they are (u32, canonical).  Leaf DIGESTS are never stored
// (they are 8x the size of the values and one compression away from them); levels 1..depth are stored
// contiguously as 8 big-endian state words per digest.  An authentication path recomputes its level-0
// sibling digest from the sibling value.
//
// Work decomposition.  Each thread owns 2^S consecutive items of one level and reduces them to one
// digest S levels up entirely in registers (no shared memory, no barriers), writing every
// intermediate digest once.  One launch therefore advances the tree by S levels; a single-CTA kernel
// finishes the last <= 9 levels.  The leaf launch can take its values from the FRI fold of the
// previous layer (fused fold-and-hash, reference src/fri/fri_commit.rs:94-97).
#include "kernels.hpp"
#include "sha256.cuh"

namespace starkb200 {

TreeShape TreeShape::make(size_t n) {
    TreeShape s;
    s.n = n;
    s.len.push_back(n);
    while (s.len.back() > 1) s.len.push_back((s.len.back() + 1) / 2);
    s.depth = (unsigned)(s.len.size() - 1);
    s.off.assign(s.depth + 1, 0);
    size_t o = 0;
    for (unsigned l = 1; l <= s.depth; l++) { s.off[l] = o; o += s.len[l]; }
    s.total = s.depth ? o : 1;   // a one-leaf tree keeps its root (= the leaf digest) in slot 0
    return s;
}

size_t merkle_path_len(size_t n, size_t idx) {
    size_t bytes = 0;
    for (size_t m = n, j = idx; m > 1; m = (m + 1) / 2, j >>= 1)
        if ((j ^ 1) < m) bytes += 32;
    return bytes;
}

constexpr int SUB = 3;            // levels advanced per launch (2^SUB items per thread)
constexpr int TOP_MAX = 512;      // a level this small is finished by the single-CTA kernel
constexpr int SMALL_MAX = 2048;   // a tree this small is built by one CTA in one launch
constexpr int SMALL_THREADS = 512;
constexpr int MERKLE_THREADS = 128;

struct LevelPtrs { uint32_t* p[SUB + 1]; };   // p[l], l = 1..SUB: storage of the l-th level produced by this launch

// One shared copy of the two-block parent hash: every call site in a kernel jumps here, so the code a
// thread walks through stays a few tens of KB (see sha256_compress).
__device__ __noinline__ void sha256_node_fn(const Digest* l, const Digest* r, Digest* out) {
    Digest a = *l, b = *r, o;
    sha256_node(a, b, o);
    *out = o;
}

__device__ __forceinline__ uint32_t fold_one(uint32_t a, uint32_t b, uint32_t s_m, uint32_t inv2_m, const FieldParams& fp) {
    return fadd(mont_mul(fadd(a, b, fp), inv2_m, fp), mont_mul(fsub(a, b, fp), s_m, fp), fp);
}
// e'[i] of the fold described by `src` (also written to src.fold_out)
__device__ __forceinline__ uint32_t fold_at(const LeafSource& src, size_t i, const FieldParams& fp) {
    uint32_t a = src.prev[i], b = src.prev[src.half + i];
    uint32_t s = mont_mul(pow_lookup(src.winv, (uint32_t)i, fp), src.sb_m, fp);
    uint32_t v = fold_one(a, b, s, src.inv2_m, fp);
    src.fold_out[i] = v;
    return v;
}

enum { SRC_VALUES = 0, SRC_FOLD = 1, SRC_DIGESTS = 2 };

// Each thread reduces its 2^SUB consecutive items to one digest SUB levels up with a small stack:
// item i is hashed, then merged upward while its index is odd.  One leaf site and one node site in the
// code; trip counts depend only on i, so a warp never diverges.  A ragged tail (cnt < 2^SUB) is finished
// by the flush loop, which applies rs_merkle's rule: a node without a right sibling is promoted.
#ifndef STARK_MERKLE_MIN_BLOCKS
#define STARK_MERKLE_MIN_BLOCKS 1
#endif
template <int SRC>
__global__ void __launch_bounds__(MERKLE_THREADS, STARK_MERKLE_MIN_BLOCKS)
merkle_subtree_kernel(LeafSource src, const uint32_t* in_digests, size_t n, int nlev, LevelPtrs lv, FieldParams fp,
                      HostResult* result, int last) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t base = t << SUB;
    if (base >= n) return;
    const int cnt = (n - base) < (size_t)(1 << SUB) ? (int)(n - base) : (1 << SUB);
    uint32_t v[1 << SUB];
    if (SRC == SRC_VALUES) {
        if (cnt == (1 << SUB)) {
            const uint4* pv = reinterpret_cast<const uint4*>(src.vals + base);
            uint4 x = pv[0], y = pv[1];
            v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
        } else {
            for (int j = 0; j < (1 << SUB); j++) v[j] = j < cnt ? src.vals[base + j] : 0u;
        }
    } else if (SRC == SRC_FOLD) {
        if (cnt == (1 << SUB) && src.winv.shift >= SUB) {
            const uint4* pa = reinterpret_cast<const uint4*>(src.prev + base);
            const uint4* pb = reinterpret_cast<const uint4*>(src.prev + src.half + base);
            uint32_t a[8], b[8];
            uint4 x = pa[0], y = pa[1];
            a[0] = x.x; a[1] = x.y; a[2] = x.z; a[3] = x.w; a[4] = y.x; a[5] = y.y; a[6] = y.z; a[7] = y.w;
            x = pb[0]; y = pb[1];
            b[0] = x.x; b[1] = x.y; b[2] = x.z; b[3] = x.w; b[4] = y.x; b[5] = y.y; b[6] = y.z; b[7] = y.w;
            // the 8 exponents share their high part: one table product per thread, one per element
            uint32_t hs = mont_mul(__ldg(src.winv.hi + (uint32_t)(base >> src.winv.shift)), src.sb_m, fp);
            uint32_t lo0 = (uint32_t)base & src.winv.mask;
#pragma unroll
            for (int j = 0; j < 8; j++)
                v[j] = fold_one(a[j], b[j], mont_mul(__ldg(src.winv.lo + lo0 + j), hs, fp), src.inv2_m, fp);
            uint4* po = reinterpret_cast<uint4*>(src.fold_out + base);
            po[0] = make_uint4(v[0], v[1], v[2], v[3]);
            po[1] = make_uint4(v[4], v[5], v[6], v[7]);
        } else {
            for (int j = 0; j < (1 << SUB); j++) v[j] = j < cnt ? fold_at(src, base + j, fp) : 0u;
        }
    }
    Digest stk[SUB + 1];
    Digest d;
#pragma unroll 1
    for (int i = 0; i < cnt; i++) {
        if (SRC == SRC_DIGESTS) d = load_digest(in_digests + 8 * (base + i));
        else sha256_leaf(0u, v[i], d);
        int level = 0;
        unsigned idx = (unsigned)i;
        while (level < nlev && (idx & 1u)) {
            sha256_node_fn(&stk[level], &d, &d);
            level++; idx >>= 1;
            store_digest(lv.p[level] + 8 * ((t << (SUB - level)) + idx), d);
        }
        if (level < nlev) stk[level] = d;
    }
    if (nlev == 0) {                                    // one-leaf tree: the root is the leaf digest
        store_digest(lv.p[1], d);
    } else if (cnt != (1 << nlev)) {                    // ragged tail: promote / merge what is pending
        bool have = false;
        for (int level = 0; level < nlev; level++) {
            bool pend = (cnt >> level) & 1;
            if (have && pend) sha256_node_fn(&stk[level], &d, &d);
            else if (pend) { d = stk[level]; have = true; }
            else if (!have) continue;
            store_digest(lv.p[level + 1] + 8 * ((t << (SUB - level - 1)) + (size_t)(cnt >> (level + 1))), d);
        }
    }
    if (last && t == 0 && result)
        for (int i = 0; i < 8; i++) result->root[i] = d.w[i];
}

// Finishes a tree whose current level has <= TOP_MAX nodes: one CTA, one barrier per level.
// Levels are contiguous in the node buffer, so `out` simply advances.
__device__ __forceinline__ void finish_levels(const uint32_t* in, int in_len, uint32_t* out, HostResult* result) {
    while (in_len > 1) {
        int out_len = (in_len + 1) >> 1;
        for (int j = threadIdx.x; j < out_len; j += blockDim.x) {
            Digest l = load_digest(in + 16 * j), o;
            if (2 * j + 1 < in_len) { Digest r = load_digest(in + 16 * j + 8); sha256_node_fn(&l, &r, &o); }
            else o = l;
            store_digest(out + 8 * j, o);
        }
        __syncthreads();
        in = out; out += 8 * (size_t)out_len; in_len = out_len;
    }
    if (threadIdx.x == 0 && result)
        for (int i = 0; i < 8; i++) result->root[i] = in[i];
}
__global__ void __launch_bounds__(TOP_MAX / 2)
merkle_top_kernel(const uint32_t* in, int in_len, uint32_t* out, HostResult* result) {
    finish_levels(in, in_len, out, result);
}

// Whole tree of a small layer (n <= SMALL_MAX) in one launch: the late FRI layers are latency bound
// (one launch + one host round trip each), so spread the leaves over one CTA's threads instead of
// giving 8 of them to each of a few threads.
template <bool FOLD>
__global__ void __launch_bounds__(SMALL_THREADS)
merkle_small_kernel(LeafSource src, int n, uint32_t* nodes, FieldParams fp, HostResult* result) {
    const int len1 = (n + 1) >> 1;
    for (int j = threadIdx.x; j < len1; j += blockDim.x) {
        uint32_t v0 = FOLD ? fold_at(src, 2 * (size_t)j, fp) : src.vals[2 * j];
        Digest l, o;
        sha256_leaf(0u, v0, l);
        if (2 * j + 1 < n) {
            uint32_t v1 = FOLD ? fold_at(src, 2 * (size_t)j + 1, fp) : src.vals[2 * j + 1];
            Digest r;
            sha256_leaf(0u, v1, r);
            sha256_node_fn(&l, &r, &o);
        } else o = l;
        store_digest(nodes + 8 * j, o);          // for n == 1 this is the root slot
    }
    __syncthreads();
    finish_levels(nodes, len1, nodes + 8 * (size_t)len1, result);
}

void merkle_build(stark_ctx* ctx, const LeafSource& src, const TreeShape& shape, uint32_t* nodes, HostResult* result) {
    const size_t n = shape.n;
    STARK_REQUIRE(n >= 1, "merkle: empty tree (MerkleTree::root() would panic on unwrap, merkle/mod.rs:25)");
    const unsigned depth = shape.depth;
    auto level_ptr = [&](unsigned l) { return nodes + 8 * shape.off[l]; };
    const bool fold = src.prev != nullptr;
    // algorithmic int-ops: 1384 per compression; leaf = 1, node = 2
    auto comp_levels = [&](unsigned from, unsigned to) { double c = 0; for (unsigned l = from; l <= to; l++) c += 2.0 * (double)shape.len[l]; return c; };
    if (n <= (size_t)SMALL_MAX) {
        KernelTimer kt(ctx, stark_ctx::CAT_MERKLE_LEAF, 1384.0 * ((double)n + comp_levels(1, depth)));
        if (fold) merkle_small_kernel<true><<<1, SMALL_THREADS, 0, ctx->stream>>>(src, (int)n, nodes, ctx->fp, result);
        else merkle_small_kernel<false><<<1, SMALL_THREADS, 0, ctx->stream>>>(src, (int)n, nodes, ctx->fp, result);
        ctx->launches++;
        STARK_CUDA(cudaGetLastError());
        return;
    }
    unsigned cur = 0;
    {
        int nlev = (int)(depth < (unsigned)SUB ? depth : SUB);
        LevelPtrs lv{};
        for (int l = 1; l <= nlev; l++) lv.p[l] = level_ptr(l);
        size_t threads = (n + (1 << SUB) - 1) >> SUB;
        unsigned blocks = (unsigned)((threads + MERKLE_THREADS - 1) / MERKLE_THREADS);
        int last = (unsigned)nlev == depth;
        KernelTimer kt(ctx, stark_ctx::CAT_MERKLE_LEAF, 1384.0 * ((double)n + comp_levels(1, nlev)));
        if (fold) merkle_subtree_kernel<SRC_FOLD><<<blocks, MERKLE_THREADS, 0, ctx->stream>>>(src, nullptr, n, nlev, lv, ctx->fp, result, last);
        else merkle_subtree_kernel<SRC_VALUES><<<blocks, MERKLE_THREADS, 0, ctx->stream>>>(src, nullptr, n, nlev, lv, ctx->fp, result, last);
        ctx->launches++;
        cur = (unsigned)nlev;
    }
    while (cur < depth) {
        size_t cur_len = shape.len[cur];
        if (cur_len <= (size_t)TOP_MAX) {
            KernelTimer kt(ctx, stark_ctx::CAT_MERKLE_NODE, 1384.0 * comp_levels(cur + 1, depth));
            merkle_top_kernel<<<1, TOP_MAX / 2, 0, ctx->stream>>>(level_ptr(cur), (int)cur_len, level_ptr(cur + 1), result);
            ctx->launches++;
            cur = depth;
            break;
        }
        int nlev = (int)((depth - cur) < (unsigned)SUB ? (depth - cur) : SUB);
        LevelPtrs lv{};
        for (int l = 1; l <= nlev; l++) lv.p[l] = level_ptr(cur + l);
        size_t threads = (cur_len + (1 << SUB) - 1) >> SUB;
        unsigned blocks = (unsigned)((threads + MERKLE_THREADS - 1) / MERKLE_THREADS);
        int last = cur + (unsigned)nlev == depth;
        KernelTimer kt(ctx, stark_ctx::CAT_MERKLE_NODE, 1384.0 * comp_levels(cur + 1, cur + nlev));
        merkle_subtree_kernel<SRC_DIGESTS><<<blocks, MERKLE_THREADS, 0, ctx->stream>>>(LeafSource{}, level_ptr(cur), cur_len, nlev, lv, ctx->fp, result, last);
        ctx->launches++;
        cur += (unsigned)nlev;
    }
    STARK_CUDA(cudaGetLastError());
}

// ---- openings: one warp per record, lane l fetches the sibling at level l ------------------------
__global__ void merkle_open_kernel(const OpenDesc* desc, size_t n_desc, uint8_t* out) {
    size_t w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned lane = threadIdx.x & 31;
    if (w >= n_desc) return;
    OpenDesc d = desc[w];
    unsigned depth = 0;
    for (unsigned long long m = d.n; m > 1; m = (m + 1) >> 1) depth++;
    // level geometry for this lane
    unsigned long long len_l = d.n, off_l = 0;   // off_l: digest offset of level `lane` (levels >= 1)
    for (unsigned l = 0; l < lane && l < depth; l++) {
        if (l >= 1) off_l += len_l;
        len_l = (len_l + 1) >> 1;
    }
    unsigned long long j = (d.idx >> lane) ^ 1ull;
    bool exists = lane < depth && j < len_l;
    unsigned mask = __ballot_sync(0xffffffffu, exists);
    uint32_t* rec = reinterpret_cast<uint32_t*>(out + d.out_off);
    if (lane == 0) {
        rec[0] = 0;                                                  // BE8(value): high word is zero (p < 2^32)
        rec[1] = __byte_perm(d.vals[d.idx], 0, 0x0123);
    }
    if (exists) {
        Digest dg;
        if (lane == 0) sha256_leaf(0u, d.vals[j], dg);
        else dg = load_digest(d.nodes + 8 * (off_l + j));
        unsigned pos = __popc(mask & ((1u << lane) - 1u));
        uint32_t* o = rec + 2 + 8 * pos;
#pragma unroll
        for (int i = 0; i < 8; i++) o[i] = __byte_perm(dg.w[i], 0, 0x0123);
    }
}

void merkle_open(stark_ctx* ctx, const OpenDesc* d_desc, size_t n_desc, uint8_t* d_out) {
    if (n_desc == 0) return;
    unsigned blocks = (unsigned)((n_desc * 32 + 127) / 128);
    KernelTimer kt(ctx, stark_ctx::CAT_OTHER, 0);
    merkle_open_kernel<<<blocks, 128, 0, ctx->stream>>>(d_desc, n_desc, d_out);
    ctx->launches++;
    STARK_CUDA(cudaGetLastError());
}

}  // namespace starkb200
