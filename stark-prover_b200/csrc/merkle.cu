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
// digest S levels up on its own (no barriers; its stack of pending left siblings is a private column of
// shared memory), writing every intermediate digest once.  One launch therefore advances the tree by S
// levels; merkle_tail_kernel finishes the last <= 15 levels in one launch.  The leaf launch can take its values from the FRI fold of the
// previous layer (fused fold-and-hash, reference src/fri/fri_commit.rs:94-97).
#include "kernels.hpp"
#include "sha256.cuh"
#include "coeff_job.cuh"

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
#ifndef STARK_MERKLE_THREADS
#define STARK_MERKLE_THREADS 128
#endif
constexpr int MERKLE_THREADS = STARK_MERKLE_THREADS;

struct LevelPtrs { uint32_t* p[SUB + 1]; };   // p[l], l = 1..SUB: storage of the l-th level produced by this launch
// run-time level -> pointer without indexing the kernel parameter dynamically (that would copy it to local memory)
__device__ __forceinline__ uint32_t* level_storage(const LevelPtrs& lv, int level) {
    static_assert(SUB == 3, "level_storage spells out SUB levels");
    return level == 1 ? lv.p[1] : level == 2 ? lv.p[2] : lv.p[3];
}

// One shared copy of the two-block parent hash: every call site in a kernel jumps here, so the code a
// thread walks through stays a few tens of KB (see sha256_compress).
__device__ __noinline__ Digest sha256_node_fn(Digest l, Digest r) {
    Digest o;
    sha256_node(l, r, o);
    return o;
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
    const unsigned bid = blockIdx.x;
    // the layer's coefficient-space fold: the first job.ctas CTAs do a slice of it each BEFORE they hash (extra CTAs beside an
    // ALU-bound grid that fits one wave cost it 12 %: the 2^20-leaf launch went 176 -> 197 us whatever their number or position)
    if (SRC == SRC_FOLD && bid < src.job.ctas) coeff_job_run(src.job, bid, fp);
    const size_t t = (size_t)bid * blockDim.x + threadIdx.x;
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
    // pending left siblings, one per level, and the thread's 8 values: word-major in shared memory (the loops
    // below index them with run-time levels / positions; in local memory they cost 190 MB of dead write-backs
    // per 2^24-leaf launch)
    __shared__ uint32_t s_stk[SUB][8][MERKLE_THREADS];
    __shared__ uint32_t s_v[1 << SUB][MERKLE_THREADS];
    const int tx = threadIdx.x;
    if (SRC != SRC_DIGESTS) {
#pragma unroll
        for (int j = 0; j < (1 << SUB); j++) s_v[j][tx] = v[j];
    }
    Digest d;
#pragma unroll 1
    for (int i = 0; i < cnt; i++) {
        if (SRC == SRC_DIGESTS) d = load_digest(in_digests + 8 * (base + i));
        else sha256_leaf32(s_v[i][tx], d);
        int level = 0;
        unsigned idx = (unsigned)i;
        while (level < nlev && (idx & 1u)) {
            Digest l;
#pragma unroll
            for (int w = 0; w < 8; w++) l.w[w] = s_stk[level][w][tx];
            d = sha256_node_fn(l, d);
            level++; idx >>= 1;
            store_digest(level_storage(lv, level) + 8 * ((t << (SUB - level)) + idx), d);
        }
        if (level < nlev) {
#pragma unroll
            for (int w = 0; w < 8; w++) s_stk[level][w][tx] = d.w[w];
        }
    }
    if (nlev == 0) {                                    // one-leaf tree: the root is the leaf digest
        store_digest(lv.p[1], d);
    } else if (cnt != (1 << nlev)) {                    // ragged tail: promote / merge what is pending
        bool have = false;
        for (int level = 0; level < nlev; level++) {
            bool pend = (cnt >> level) & 1;
            if (have && pend) {
                Digest l;
#pragma unroll
                for (int w = 0; w < 8; w++) l.w[w] = s_stk[level][w][tx];
                d = sha256_node_fn(l, d);
            } else if (pend) {
#pragma unroll
                for (int w = 0; w < 8; w++) d.w[w] = s_stk[level][w][tx];
                have = true;
            } else if (!have) continue;
            store_digest(level_storage(lv, level + 1) + 8 * ((t << (SUB - level - 1)) + (size_t)(cnt >> (level + 1))), d);
        }
    }
    if (last && t == 0 && result)
        for (int i = 0; i < 8; i++) result->root[i] = d.w[i];
}

// ---- latency-bound part of every tree: one launch, one node per thread per level ------------------------
// Below ~2^17 nodes a level no longer keeps the machine busy for long and what matters is the length of the
// dependency chain: 2 compressions per level.  Giving each thread 8 items (merkle_subtree_kernel) makes that chain
// 14 compressions per 3 levels.  This kernel instead walks the remaining levels with one parent per thread.
// A CTA owns 256 consecutive items and reduces them 8 levels to one node, handing digests from level to
// level through shared memory (word-major rows: a parent reads its two children as one conflict-free 64-bit
// load per word); every node is also written to its place in the stored tree.  The CTA that finishes last
// (an atomic ticket, no CTA ever waits for another) picks the chunk roots out of L2 (up to 512: two levels straight
// from L2, then through shared memory again) and walks the remaining levels the same way.  128-thread CTAs: with up
// to 148 of them every warp gets a scheduler (SM sub-partition) to itself, so the ALU pipe (one warp instruction per
// 2 cycles) is not shared -- two warps per scheduler double the level time.
constexpr int TAIL_THREADS = 128;
// Where this kernel takes over from the 3-levels-per-launch kernel, in CTAs of 256 items.  Measured on the headline
// step (tools/variants_tail.sh): 128 / 512 / 1024 / 2048 / 4096 CTAs -> commit phase 7.24 / 7.10 / 7.18 / 7.33 / 7.53 ms.
#ifndef STARK_TAIL_MAX_CTAS
#define STARK_TAIL_MAX_CTAS 512
#endif
constexpr int TAIL_MAX_CTAS = STARK_TAIL_MAX_CTAS;
constexpr int TAIL_MAX = 2 * TAIL_THREADS * TAIL_MAX_CTAS;      // 2^17 items

__device__ __forceinline__ Digest load_digest_cg(const uint32_t* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldcg(q), b = __ldcg(q + 1);
    Digest d;
    d.w[0] = a.x; d.w[1] = a.y; d.w[2] = a.z; d.w[3] = a.w;
    d.w[4] = b.x; d.w[5] = b.y; d.w[6] = b.z; d.w[7] = b.w;
    return d;
}

struct TailArgs {
    LeafSource src;              // SRC_VALUES / SRC_FOLD: the items are leaf values
    const uint32_t* in_digests;  // SRC_DIGESTS: the items are stored digests
    int n_items;
    uint32_t* out;               // storage of the next level; the levels above follow contiguously
    unsigned* ticket;            // zero between launches (the last CTA resets it)
    HostResult* result;
    FieldParams fp;
    HostTop* top;                // early hand-over of the top levels (common.hpp: HostTop); nullptr = off
    uint32_t seq;
};

// The whole CTA calls this once, with the level it has just produced complete in its threads' registers (thread t holds
// node t, t < len <= HOST_TOP_MAX): nodes, system-scope fence by every writer, barrier, length, fence, sequence number.
__device__ __forceinline__ void publish_top(HostTop* top, uint32_t seq, int len, const Digest& o, int t) {
    if (t < len) {
#pragma unroll
        for (int w = 0; w < 8; w++) top->node[t][w] = o.w[w];
        __threadfence_system();
    }
    __syncthreads();
    if (t == 0) {
        top->top_len = (uint32_t)len;
        __threadfence_system();
        *reinterpret_cast<volatile uint32_t*>(&top->top_seq) = seq;
    }
}

template <int SRC>
__global__ void __launch_bounds__(TAIL_THREADS) merkle_tail_kernel(TailArgs a) {
    __shared__ __align__(8) uint32_t sm[2][8][TAIL_THREADS + 2];
    __shared__ int s_last;
    const int t = threadIdx.x;
    const unsigned bid = blockIdx.x;
    unsigned nblk = gridDim.x;
    if (SRC == SRC_FOLD && a.src.job.ctas) {     // the layer's coefficient-space fold rides in the highest block indices
        nblk -= a.src.job.ctas;
        if (bid >= nblk) { coeff_job_run(a.src.job, bid - nblk, a.fp); return; }
    }
    int len = a.n_items;                         // length of the level being consumed
    uint32_t* out = a.out;                       // storage of the level being produced
    if (len == 1) {                              // one-leaf tree: the root is the leaf digest (slot 0)
        if (bid == 0 && t == 0) {
            Digest d;
            if (SRC == SRC_DIGESTS) d = load_digest_cg(a.in_digests);
            else sha256_leaf32(SRC == SRC_FOLD ? fold_at(a.src, 0, a.fp) : a.src.vals[0], d);
            store_digest(out, d);
            if (a.result) for (int i = 0; i < 8; i++) a.result->root[i] = d.w[i];
            if (a.top) {
                for (int i = 0; i < 8; i++) a.top->node[0][i] = d.w[i];
                a.top->top_len = 1;
                __threadfence_system();
                *reinterpret_cast<volatile uint32_t*>(&a.top->top_seq) = a.seq;
            }
        }
        return;
    }
    bool published = a.top == nullptr;
    int base = (int)bid * 2 * TAIL_THREADS;      // the CTA's window of that level: `width` nodes from `base`
    int width = 2 * TAIL_THREADS;
    int buf = 0;                                 // sm[buf] holds the window (except for the first level)
    bool first = true;
    Digest o;
    for (int phase = 0;; phase++) {
        while (width > 1) {
            const int out_len = (len + 1) >> 1;
            const int j = (base >> 1) + t;       // parent index within its level
            if (t < (width >> 1) && j < out_len) {
                Digest l, r;
                const bool pair = 2 * j + 1 < len;
                if (first) {
                    if (SRC == SRC_DIGESTS) {
                        l = load_digest_cg(a.in_digests + 16 * (size_t)j);
                        if (pair) r = load_digest_cg(a.in_digests + 16 * (size_t)j + 8);
                    } else {
                        sha256_leaf32(SRC == SRC_FOLD ? fold_at(a.src, 2 * (size_t)j, a.fp) : a.src.vals[2 * j], l);
                        if (pair) sha256_leaf32(SRC == SRC_FOLD ? fold_at(a.src, 2 * (size_t)j + 1, a.fp) : a.src.vals[2 * j + 1], r);
                    }
                } else {
#pragma unroll
                    for (int w = 0; w < 8; w++) {
                        uint2 v = *reinterpret_cast<const uint2*>(&sm[buf][w][2 * t]);
                        l.w[w] = v.x; r.w[w] = v.y;
                    }
                }
                if (pair) o = sha256_node_fn(l, r); else o = l;      // lone node promoted
                store_digest(out + 8 * (size_t)j, o);
#pragma unroll
                for (int w = 0; w < 8; w++) sm[buf ^ 1][w][t] = o.w[w];
            }
            first = false;
            out += 8 * (size_t)out_len; len = out_len; base >>= 1; width >>= 1;
            if (!published && len <= HOST_TOP_MAX && (phase > 0 || nblk == 1)) {      // this CTA holds the whole level
                publish_top(a.top, a.seq, len, o, t);
                published = true;
            }
            if (len == 1) {                                            // that was the root (CTA 0, thread 0 holds it)
                if (t == 0 && a.result) for (int i = 0; i < 8; i++) a.result->root[i] = o.w[i];
                return;
            }
            __syncthreads();
            buf ^= 1;
        }
        // the window is reduced to one node and the tree is not finished: more than one CTA took part
        if (phase == 0) {
            if (t == 0) {
                __threadfence();                                       // this CTA's chunk root is visible before the ticket
                unsigned tk = atomicAdd(a.ticket, 1u);
                s_last = tk == nblk - 1;
                if (s_last) { *a.ticket = 0; __threadfence(); }        // every chunk root is in L2; re-arm for the next launch
            }
            __syncthreads();
            if (!s_last) return;
        }
        // the last CTA continues with the level just produced (`len` nodes, stored right below `out`)
        const uint32_t* in = out - 8 * (size_t)len;
        while (len > TAIL_THREADS) {                                   // wider than a window (n_items > TAIL_MAX): via L2
            const int out_len = (len + 1) >> 1;
            for (int j = t; j < out_len; j += TAIL_THREADS) {
                Digest l = load_digest_cg(in + 16 * (size_t)j), q;
                if (2 * j + 1 < len) { Digest r = load_digest_cg(in + 16 * (size_t)j + 8); q = sha256_node_fn(l, r); }
                else q = l;
                store_digest(out + 8 * (size_t)j, q);
            }
            __syncthreads();
            in = out; out += 8 * (size_t)out_len; len = out_len;
        }
        if (t < len) {
            Digest d = load_digest_cg(in + 8 * (size_t)t);
#pragma unroll
            for (int w = 0; w < 8; w++) sm[buf][w][t] = d.w[w];
            o = d;
        }
        __syncthreads();
        if (!published && len <= HOST_TOP_MAX) {                       // few chunk roots: they are the level to hand over
            publish_top(a.top, a.seq, len, o, t);
            published = true;
        }
        base = 0; width = TAIL_THREADS;
    }
}

template <int SRC>
static void launch_tail(stark_ctx* ctx, const LeafSource& src, const uint32_t* in_digests, size_t n_items, uint32_t* out,
                        HostResult* result, uint32_t top_seq) {
    if (!ctx->tail_counter.p) {                      // one ticket per context (= per stream): launches are ordered
        ctx->tail_counter = DevBuf(sizeof(unsigned), ctx->stream);
        STARK_CUDA(cudaMemsetAsync(ctx->tail_counter.p, 0, sizeof(unsigned), ctx->stream));
    }
    int ctas = (int)((n_items + 2 * TAIL_THREADS - 1) / (2 * TAIL_THREADS));
    if (ctas < 1) ctas = 1;
    TailArgs a{src, in_digests, (int)n_items, out, ctx->tail_counter.as<unsigned>(), result, ctx->fp, top_seq ? ctx->d_top : nullptr, top_seq};
    if (SRC == SRC_FOLD) ctas += (int)src.job.ctas;
    merkle_tail_kernel<SRC><<<ctas, TAIL_THREADS, 0, ctx->stream>>>(a);
    ctx->launches++;
}

unsigned merkle_first_launch_threads(size_t n) { return n <= (size_t)TAIL_MAX ? TAIL_THREADS : MERKLE_THREADS; }

void merkle_build(stark_ctx* ctx, const LeafSource& src, const TreeShape& shape, uint32_t* nodes, HostResult* result, uint32_t top_seq) {
    const size_t n = shape.n;
    STARK_REQUIRE(n >= 1, "merkle: empty tree (MerkleTree::root() would panic on unwrap, merkle/mod.rs:25)");
    const unsigned depth = shape.depth;
    auto level_ptr = [&](unsigned l) { return nodes + 8 * shape.off[l]; };
    const bool fold = src.prev != nullptr;
    // algorithmic int-ops: 1384 per compression; leaf = 1, node = 2
    auto comp_levels = [&](unsigned from, unsigned to) { double c = 0; for (unsigned l = from; l <= to; l++) c += 2.0 * (double)shape.len[l]; return c; };
    if (n <= (size_t)TAIL_MAX) {                     // a small layer: the whole tree in one launch
        KernelTimer kt(ctx, stark_ctx::CAT_MERKLE_LEAF, 1384.0 * ((double)n + comp_levels(1, depth)));
        if (fold) launch_tail<SRC_FOLD>(ctx, src, nullptr, n, nodes, result, top_seq);
        else launch_tail<SRC_VALUES>(ctx, src, nullptr, n, nodes, result, top_seq);
        STARK_CUDA(cudaGetLastError());
        return;
    }
    unsigned cur = 0;
    {
        LevelPtrs lv{};
        for (int l = 1; l <= SUB; l++) lv.p[l] = level_ptr(l);        // n > TAIL_MAX: depth > SUB
        size_t threads = (n + (1 << SUB) - 1) >> SUB;
        unsigned blocks = (unsigned)((threads + MERKLE_THREADS - 1) / MERKLE_THREADS);
        KernelTimer kt(ctx, stark_ctx::CAT_MERKLE_LEAF, 1384.0 * ((double)n + comp_levels(1, SUB)));
        if (fold) {
            LeafSource s2 = src;
            if (s2.job.ctas > blocks) s2.job.ctas = blocks;          // grid-stride: any number of CTAs covers the job
            merkle_subtree_kernel<SRC_FOLD><<<blocks, MERKLE_THREADS, 0, ctx->stream>>>(s2, nullptr, n, SUB, lv, ctx->fp, result, 0);
        }
        else merkle_subtree_kernel<SRC_VALUES><<<blocks, MERKLE_THREADS, 0, ctx->stream>>>(src, nullptr, n, SUB, lv, ctx->fp, result, 0);
        ctx->launches++;
        cur = SUB;
    }
    while (shape.len[cur] > (size_t)TAIL_MAX) {      // throughput-bound levels: 3 per launch
        size_t cur_len = shape.len[cur];
        LevelPtrs lv{};
        for (int l = 1; l <= SUB; l++) lv.p[l] = level_ptr(cur + l);
        size_t threads = (cur_len + (1 << SUB) - 1) >> SUB;
        unsigned blocks = (unsigned)((threads + MERKLE_THREADS - 1) / MERKLE_THREADS);
        KernelTimer kt(ctx, stark_ctx::CAT_MERKLE_NODE, 1384.0 * comp_levels(cur + 1, cur + SUB));
        merkle_subtree_kernel<SRC_DIGESTS><<<blocks, MERKLE_THREADS, 0, ctx->stream>>>(LeafSource{}, level_ptr(cur), cur_len, SUB, lv, ctx->fp, result, 0);
        ctx->launches++;
        cur += SUB;
    }
    {
        KernelTimer kt(ctx, stark_ctx::CAT_MERKLE_NODE, 1384.0 * comp_levels(cur + 1, depth));
        launch_tail<SRC_DIGESTS>(ctx, LeafSource{}, level_ptr(cur), shape.len[cur], level_ptr(cur + 1), result, top_seq);
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
        if (lane == 0) sha256_leaf32(d.vals[j], dg);
        else dg = load_digest(d.nodes + 8 * (off_l + j));
        unsigned pos = __popc(mask & ((1u << lane) - 1u));
        uint32_t* o = rec + 2 + 8 * pos;
#pragma unroll
        for (int i = 0; i < 8; i++) o[i] = __byte_perm(dg.w[i], 0, 0x0123);
    }
}

// ---- one query index across the layers of a FRI proof: everything the kernel needs travels in its arguments
// (no descriptor upload), including the byte offset of every record.  Warp w opens record w:
// layer first + w/2, position idx (even w) or its sibling (odd w).
__global__ void fri_open_one_kernel(FriOpenArgs a) {
    const unsigned w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const unsigned n_rec = 2 * (a.n_layers - a.first);
    if (w >= n_rec) return;
    const unsigned before = a.rec_off[w];          // computed by the host together with the total (api.cu: open_one_index)
    const FriLayerDesc& L = a.layers[a.first + w / 2];
    unsigned long long idx = a.index % L.n;
    if (w & 1) idx = (idx + L.n / 2) % L.n;
    unsigned depth = 0;
    for (unsigned long long m = L.n; m > 1; m = (m + 1) >> 1) depth++;
    unsigned long long len_l = L.n, off_l = 0;
    for (unsigned l = 0; l < lane && l < depth; l++) {
        if (l >= 1) off_l += len_l;
        len_l = (len_l + 1) >> 1;
    }
    unsigned long long j = (idx >> lane) ^ 1ull;
    bool exists = lane < depth && j < len_l;
    unsigned mask = __ballot_sync(0xffffffffu, exists);
    uint32_t* rec = reinterpret_cast<uint32_t*>(a.out + before);
    if (lane == 0) {
        rec[0] = 0;
        rec[1] = __byte_perm(L.vals[idx], 0, 0x0123);
    }
    if (exists) {
        Digest dg;
        if (lane == 0) sha256_leaf32(L.vals[j], dg);
        else dg = load_digest(L.nodes + 8 * (off_l + j));
        unsigned pos = __popc(mask & ((1u << lane) - 1u));
        uint32_t* o = rec + 2 + 8 * pos;
#pragma unroll
        for (int i = 0; i < 8; i++) o[i] = __byte_perm(dg.w[i], 0, 0x0123);
    }
}
void fri_open_one(stark_ctx* ctx, const FriOpenArgs& a) {
    unsigned n_rec = 2 * (a.n_layers - a.first);
    if (!n_rec) return;
    KernelTimer kt(ctx, stark_ctx::CAT_OTHER, 0);
    fri_open_one_kernel<<<(n_rec * 32 + 127) / 128, 128, 0, ctx->stream>>>(a);
    ctx->launches++;
    STARK_CUDA(cudaGetLastError());
}

// ---- opening server: decommit_fri draws every query index from the transcript state left by the previous query's
// messages (fri_commit.rs:168-179), so the openings are a chain of host <-> device round trips.  One resident CTA
// replaces the launch + stream synchronisation per query: the host posts (sequence, index) as ONE 64-bit word in mapped
// pinned memory, thread 0 polls it, the CTA writes the records of all layers straight into the mapped output buffer
// and answers with (sequence, bytes) after one system-scope fence.  Geometry per record is computed on the device
// (sibling-exists mask per level, prefix sum of the record sizes); the sibling-leaf digests of all records are hashed
// by neighbouring threads in lock step.  The kernel leaves when the host posts ~0 or nothing arrives for `idle_ns`.
__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__global__ void __launch_bounds__(1024, 1) fri_open_server_kernel(FriOpenArgs a, FriServerBox* box, unsigned long long idle_ns) {
    __shared__ unsigned long long s_req;
    __shared__ uint32_t s_off[2 * FRI_MAX_LAYERS + 1], s_mask[2 * FRI_MAX_LAYERS];
    const unsigned n_rec = 2 * (a.n_layers - a.first);
    for (unsigned long long seq = 1;; seq++) {
        if (threadIdx.x == 0) {
            const unsigned long long t0 = global_timer_ns();
            unsigned long long r;
            for (;;) {
                r = ld_sys_u64(&box->req);
                if ((r >> 32) == seq || r == ~0ull) break;
                if (global_timer_ns() - t0 > idle_ns) { r = ~0ull; break; }
            }
            s_req = r;
        }
        __syncthreads();
        const unsigned long long req = s_req;
        if (req == ~0ull) return;
        const unsigned long long index = req & 0xffffffffull;
        // record geometry: which levels have a sibling, and the record's size
        if (threadIdx.x < n_rec) {
            const FriLayerDesc& L = a.layers[a.first + threadIdx.x / 2];
            unsigned long long idx = index % L.n;
            if (threadIdx.x & 1) idx = (idx + L.n / 2) % L.n;
            uint32_t mask = 0;
            unsigned l = 0;
            for (unsigned long long m = L.n, j = idx; m > 1; m = (m + 1) >> 1, j >>= 1, l++)
                if ((j ^ 1ull) < m) mask |= 1u << l;
            s_mask[threadIdx.x] = mask;
            s_off[threadIdx.x + 1] = 8u + 32u * (unsigned)__popc(mask);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            s_off[0] = 0;
            for (unsigned r = 0; r < n_rec; r++) s_off[r + 1] += s_off[r];
        }
        __syncthreads();
        // value + level-0 sibling digest: one thread per record
        if (threadIdx.x < n_rec) {
            const FriLayerDesc& L = a.layers[a.first + threadIdx.x / 2];
            unsigned long long idx = index % L.n;
            if (threadIdx.x & 1) idx = (idx + L.n / 2) % L.n;
            uint32_t* rec = reinterpret_cast<uint32_t*>(a.out + s_off[threadIdx.x]);
            rec[0] = 0;
            rec[1] = __byte_perm(L.vals[idx], 0, 0x0123);
            if (s_mask[threadIdx.x] & 1u) {
                Digest dg;
                sha256_leaf32(L.vals[idx ^ 1ull], dg);
#pragma unroll
                for (int i = 0; i < 8; i++) rec[2 + i] = __byte_perm(dg.w[i], 0, 0x0123);
            }
        }
        // stored levels: one thread per (record, level >= 1)
        for (unsigned item = threadIdx.x; item < n_rec * 32; item += blockDim.x) {
            const unsigned r = item >> 5, l = item & 31;
            const uint32_t mask = s_mask[r];
            if (l == 0 || !((mask >> l) & 1u)) continue;
            const FriLayerDesc& L = a.layers[a.first + r / 2];
            unsigned long long idx = index % L.n;
            if (r & 1) idx = (idx + L.n / 2) % L.n;
            unsigned long long len_l = L.n, off_l = 0;
            for (unsigned m = 0; m < l; m++) {
                if (m >= 1) off_l += len_l;
                len_l = (len_l + 1) >> 1;
            }
            const unsigned long long j = (idx >> l) ^ 1ull;
            const Digest dg = load_digest(L.nodes + 8 * (off_l + j));
            uint32_t* o = reinterpret_cast<uint32_t*>(a.out + s_off[r]) + 2 + 8 * __popc(mask & ((1u << l) - 1u));
#pragma unroll
            for (int i = 0; i < 8; i++) o[i] = __byte_perm(dg.w[i], 0, 0x0123);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            st_sys_u64(&box->done, (seq << 32) | s_off[n_rec]);
        }
    }
}
void fri_open_server_launch(stark_ctx* ctx, const FriOpenArgs& a, FriServerBox* d_box, unsigned long long idle_ns) {
    fri_open_server_kernel<<<1, 1024, 0, ctx->stream>>>(a, d_box, idle_ns);
    ctx->launches++;
    STARK_CUDA(cudaGetLastError());
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
