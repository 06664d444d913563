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
constexpr int MERKLE_THREADS = 128;

struct LevelPtrs { uint32_t* p[SUB + 1]; };   // p[l], l = 1..SUB: storage of the l-th level produced by this launch

// ---- register-resident subtree over leaf values --------------------------------------------------
template <int L, int LI>
__device__ __forceinline__ void leaf_sub_full(const uint32_t (&v)[1 << SUB], size_t t, const LevelPtrs& lv, Digest& out) {
    if constexpr (L == 0) {
        sha256_leaf(0u, v[LI], out);
    } else {
        Digest l, r;
        leaf_sub_full<L - 1, 2 * LI>(v, t, lv, l);
        leaf_sub_full<L - 1, 2 * LI + 1>(v, t, lv, r);
        sha256_node(l, r, out);
        store_digest(lv.p[L] + 8 * ((t << (SUB - L)) + LI), out);
    }
}
template <int L, int LI>
__device__ __forceinline__ void node_sub_full(const uint32_t* in, size_t t, const LevelPtrs& lv, Digest& out) {
    if constexpr (L == 0) {
        out = load_digest(in + 8 * LI);
    } else {
        Digest l, r;
        node_sub_full<L - 1, 2 * LI>(in, t, lv, l);
        node_sub_full<L - 1, 2 * LI + 1>(in, t, lv, r);
        sha256_node(l, r, out);
        store_digest(lv.p[L] + 8 * ((t << (SUB - L)) + LI), out);
    }
}
// Ragged tail (at most one thread per launch): generic loops, local memory, promotion rule.
__device__ __noinline__ void sub_ragged(Digest* buf, int cnt, int nlev, size_t t, const LevelPtrs& lv) {
    int m = cnt;
    for (int L = 1; L <= nlev; L++) {
        int m2 = (m + 1) >> 1;
        for (int j = 0; j < m2; j++) {
            Digest o;
            if (2 * j + 1 < m) sha256_node(buf[2 * j], buf[2 * j + 1], o);
            else o = buf[2 * j];                                   // lone node promoted
            buf[j] = o;
            store_digest(lv.p[L] + 8 * ((t << (SUB - L)) + (size_t)j), o);
        }
        m = m2;
    }
}

__device__ __forceinline__ uint32_t fold_one(uint32_t a, uint32_t b, uint32_t s_m, uint32_t inv2_m, const FieldParams& fp) {
    return fadd(mont_mul(fadd(a, b, fp), inv2_m, fp), mont_mul(fsub(a, b, fp), s_m, fp), fp);
}

template <bool FOLD>
__global__ void __launch_bounds__(MERKLE_THREADS)
merkle_leaf_kernel(LeafSource src, size_t n, int nlev, LevelPtrs lv, FieldParams fp, HostResult* result, int last) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t base = t << SUB;
    if (base >= n) return;
    int cnt = (n - base) < (size_t)(1 << SUB) ? (int)(n - base) : (1 << SUB);
    uint32_t v[1 << SUB];
    if (cnt == (1 << SUB)) {
        if constexpr (FOLD) {
            const uint4* pa = reinterpret_cast<const uint4*>(src.prev + base);
            const uint4* pb = reinterpret_cast<const uint4*>(src.prev + src.half + base);
            uint32_t a[8], b[8];
            uint4 x = pa[0], y = pa[1];
            a[0] = x.x; a[1] = x.y; a[2] = x.z; a[3] = x.w; a[4] = y.x; a[5] = y.y; a[6] = y.z; a[7] = y.w;
            x = pb[0]; y = pb[1];
            b[0] = x.x; b[1] = x.y; b[2] = x.z; b[3] = x.w; b[4] = y.x; b[5] = y.y; b[6] = y.z; b[7] = y.w;
            if (src.winv.shift >= SUB) {
                // the 8 exponents share their high part: one table product per thread, one per element
                uint32_t hs = mont_mul(__ldg(src.winv.hi + (uint32_t)(base >> src.winv.shift)), src.sb_m, fp);
                uint32_t lo0 = (uint32_t)base & src.winv.mask;
#pragma unroll
                for (int j = 0; j < 8; j++)
                    v[j] = fold_one(a[j], b[j], mont_mul(__ldg(src.winv.lo + lo0 + j), hs, fp), src.inv2_m, fp);
            } else {
#pragma unroll
                for (int j = 0; j < 8; j++)
                    v[j] = fold_one(a[j], b[j], mont_mul(pow_lookup(src.winv, (uint32_t)(base + j), fp), src.sb_m, fp), src.inv2_m, fp);
            }
            uint4* po = reinterpret_cast<uint4*>(src.fold_out + base);
            po[0] = make_uint4(v[0], v[1], v[2], v[3]);
            po[1] = make_uint4(v[4], v[5], v[6], v[7]);
        } else {
            const uint4* pv = reinterpret_cast<const uint4*>(src.vals + base);
            uint4 x = pv[0], y = pv[1];
            v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
        }
    } else {
        for (int j = 0; j < cnt; j++) {
            if constexpr (FOLD) {
                uint32_t a = src.prev[base + j], b = src.prev[src.half + base + j];
                uint32_t s = mont_mul(pow_lookup(src.winv, (uint32_t)(base + j), fp), src.sb_m, fp);
                v[j] = fold_one(a, b, s, src.inv2_m, fp);
                src.fold_out[base + j] = v[j];
            } else {
                v[j] = src.vals[base + j];
            }
        }
    }
    if (nlev == 0) {                     // one-leaf tree: root = leaf digest
        Digest d; sha256_leaf(0u, v[0], d);
        store_digest(lv.p[1], d);
        if (result) for (int i = 0; i < 8; i++) result->root[i] = d.w[i];
        return;
    }
    Digest out;
    if (cnt == (1 << SUB) && nlev == SUB) {
        leaf_sub_full<SUB, 0>(v, t, lv, out);
    } else {
        Digest buf[1 << SUB];
        for (int j = 0; j < cnt; j++) sha256_leaf(0u, v[j], buf[j]);
        sub_ragged(buf, cnt, nlev, t, lv);
        out = buf[0];
    }
    if (last && t == 0 && result)
        for (int i = 0; i < 8; i++) result->root[i] = out.w[i];
}

__global__ void __launch_bounds__(MERKLE_THREADS)
merkle_node_kernel(const uint32_t* in, size_t n, int nlev, LevelPtrs lv, HostResult* result, int last) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t base = t << SUB;
    if (base >= n) return;
    int cnt = (n - base) < (size_t)(1 << SUB) ? (int)(n - base) : (1 << SUB);
    Digest out;
    if (cnt == (1 << SUB) && nlev == SUB) {
        node_sub_full<SUB, 0>(in + 8 * base, t, lv, out);
    } else {
        Digest buf[1 << SUB];
        for (int j = 0; j < cnt; j++) buf[j] = load_digest(in + 8 * (base + j));
        sub_ragged(buf, cnt, nlev, t, lv);
        out = buf[0];
    }
    if (last && t == 0 && result)
        for (int i = 0; i < 8; i++) result->root[i] = out.w[i];
}

// Finishes a tree whose current level has <= TOP_MAX nodes: one CTA, one barrier per level.
// Levels are contiguous in the node buffer, so `out` simply advances.
__global__ void __launch_bounds__(TOP_MAX / 2)
merkle_top_kernel(const uint32_t* in, int in_len, uint32_t* out, HostResult* result) {
    while (in_len > 1) {
        int out_len = (in_len + 1) >> 1;
        for (int j = threadIdx.x; j < out_len; j += blockDim.x) {
            Digest l = load_digest(in + 16 * j), o;
            if (2 * j + 1 < in_len) { Digest r = load_digest(in + 16 * j + 8); sha256_node(l, r, o); }
            else o = l;
            store_digest(out + 8 * j, o);
        }
        __syncthreads();
        in = out; out += 8 * (size_t)out_len; in_len = out_len;
    }
    if (threadIdx.x == 0 && result)
        for (int i = 0; i < 8; i++) result->root[i] = in[i];
}

void merkle_build(stark_ctx* ctx, const LeafSource& src, const TreeShape& shape, uint32_t* nodes, HostResult* result) {
    const size_t n = shape.n;
    STARK_REQUIRE(n >= 1, "merkle: empty tree (MerkleTree::root() would panic on unwrap, merkle/mod.rs:25)");
    const unsigned depth = shape.depth;
    auto level_ptr = [&](unsigned l) { return nodes + 8 * shape.off[l]; };
    const bool fold = src.prev != nullptr;
    unsigned cur = 0;
    {
        int nlev = (int)(depth < (unsigned)SUB ? depth : SUB);
        LevelPtrs lv{};
        for (int l = 1; l <= nlev; l++) lv.p[l] = level_ptr(l);
        if (depth == 0) lv.p[1] = nodes;
        size_t threads = (n + (1 << SUB) - 1) >> SUB;
        unsigned blocks = (unsigned)((threads + MERKLE_THREADS - 1) / MERKLE_THREADS);
        int last = (unsigned)nlev == depth;
        // algorithmic int-ops: 1384 per compression; leaf = 1, node = 2; nodes in levels 1..nlev
        double comp = (double)n;
        for (int l = 1; l <= nlev; l++) comp += 2.0 * (double)shape.len[l];
        KernelTimer kt(ctx, stark_ctx::CAT_MERKLE_LEAF, 1384.0 * comp);
        if (fold) merkle_leaf_kernel<true><<<blocks, MERKLE_THREADS, 0, ctx->stream>>>(src, n, nlev, lv, ctx->fp, result, last);
        else merkle_leaf_kernel<false><<<blocks, MERKLE_THREADS, 0, ctx->stream>>>(src, n, nlev, lv, ctx->fp, result, last);
        ctx->launches++;
        cur = (unsigned)nlev;
    }
    while (cur < depth) {
        size_t cur_len = shape.len[cur];
        if (cur_len <= (size_t)TOP_MAX) {
            double comp = 0;
            for (unsigned l = cur + 1; l <= depth; l++) comp += 2.0 * (double)shape.len[l];
            KernelTimer kt(ctx, stark_ctx::CAT_MERKLE_NODE, 1384.0 * comp);
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
        double comp = 0;
        for (int l = 1; l <= nlev; l++) comp += 2.0 * (double)shape.len[cur + l];
        KernelTimer kt(ctx, stark_ctx::CAT_MERKLE_NODE, 1384.0 * comp);
        merkle_node_kernel<<<blocks, MERKLE_THREADS, 0, ctx->stream>>>(level_ptr(cur), cur_len, nlev, lv, result, last);
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
