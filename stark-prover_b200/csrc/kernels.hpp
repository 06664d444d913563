// kernels.hpp — host-callable launchers for the CUDA kernels (implemented in the *.cu files).
#pragma once
#include "common.hpp"

namespace starkb200 {

// ---------------- Merkle (merkle.cu) ----------------
struct TreeShape {
    size_t n = 0;                      // leaves
    unsigned depth = 0;                // levels above the leaves
    std::vector<size_t> len;           // len[l], l = 0..depth   (len[0] = n)
    std::vector<size_t> off;           // digest offset of level l (l >= 1) inside the node buffer
    size_t total = 0;                  // digests in the node buffer (levels 1..depth, +1 slot when depth == 0)
    static TreeShape make(size_t n);
};

// Coefficient-space fold of a FRI layer (fri_commit.rs:32-50: even + beta * odd, with the exact degree the loop condition
// at :89 needs) riding in the FIRST launch of that layer's tree.  Throughput-bound launches (merkle_subtree_kernel): the
// first `ctas` CTAs do a slice each before they hash.  The latency-bound one-launch trees (merkle_tail_kernel, one warp per
// scheduler): `ctas` EXTRA CTAs at the end of the grid.  It is a few microseconds of memory work beside ALU-bound hashing; as
// a launch of its own it cost ~6-14 us per layer (21 of them per proof).
struct CoeffJob {
    const uint32_t* c = nullptr;       // len coefficients
    uint32_t* out = nullptr;           // out_len = (len + 1) / 2 folded coefficients
    uint32_t len = 0, out_len = 0;
    uint32_t beta_m = 0;               // beta, Montgomery form
    unsigned ctas = 0;                 // 0 = no job
    HostResult* result = nullptr;      // degree + 1 of the folded polynomial lands in result->degree_plus1
    DegScratch* scratch = nullptr;
    HostTop* top = nullptr;            // early hand-over (common.hpp): after degree_plus1, top->deg_seq = seq
    uint32_t seq = 0;
};
constexpr unsigned COEFF_JOB_MAX_CTAS = 96;      // measured: 8 / 32 / 96 / 296 -> commit phase 7.82 / 7.17 / 7.07 / 7.13 ms

// Source of the leaf VALUES of a tree: either an existing layer, or the FRI fold of the previous layer
// computed on the fly (and written to `fold_out`) — the fused fold-and-hash of fri_commit.rs:94-97.
struct LeafSource {
    const uint32_t* vals = nullptr;    // plain: leaf values
    // fold: e'[i] = (a+b)/2 + s_i (a-b),  a = prev[i], b = prev[i+half], s_i = beta/(2*offset) * w^-i
    const uint32_t* prev = nullptr;
    uint32_t* fold_out = nullptr;
    size_t half = 0;
    uint32_t inv2_m = 0;               // 1/2, Montgomery form
    uint32_t sb_m = 0;                 // beta/(2*offset), Montgomery form
    PowTable winv{};                   // w^-i for this layer's size
    CoeffJob job{};                    // fold launches only
};
// threads per CTA of the first launch merkle_build makes for a tree of n leaves (what a CoeffJob CTA will have)
unsigned merkle_first_launch_threads(size_t n);
// fills src.job (fri.cu): c -> out with beta, degree to `result`
void coeff_fold_job(stark_ctx* ctx, const uint32_t* c, size_t len, uint32_t beta_m, uint32_t* out, HostResult* result, unsigned cta_threads, CoeffJob& job);
// the same fold as a launch of its own
void coeff_fold(stark_ctx* ctx, const uint32_t* c, size_t len, uint32_t beta_m, uint32_t* out, HostResult* result);

// Builds levels 1..depth into `nodes` (TreeShape layout); when `result` is non-null the root's 8 state
// words are also written there (mapped host memory).  Leaf digests are not stored.
// top_seq != 0: the launch that finishes the tree also publishes its first level of <= 32 nodes to ctx->d_top under that
// sequence number (HostTop, common.hpp), so that the host need not wait for the kernel to end.
void merkle_build(stark_ctx* ctx, const LeafSource& src, const TreeShape& shape, uint32_t* nodes,
                  HostResult* result, uint32_t top_seq = 0);

// One authentication path / element read per descriptor; see merkle.cu.
struct OpenDesc {
    const uint32_t* vals;      // leaf values of the tree
    const uint32_t* nodes;     // node buffer (levels 1..depth)
    unsigned long long n;      // leaves
    unsigned long long idx;    // leaf index
    unsigned long long out_off;// byte offset of this record in the output: BE8(value) || path bytes
};
void merkle_open(stark_ctx* ctx, const OpenDesc* d_desc, size_t n_desc, uint8_t* d_out);
// One query index across the layers of a FRI proof, descriptors passed by value as kernel arguments.
constexpr int FRI_MAX_LAYERS = 40;
struct FriLayerDesc { const uint32_t* vals; const uint32_t* nodes; unsigned long long n; };
struct FriOpenArgs {
    FriLayerDesc layers[FRI_MAX_LAYERS];
    unsigned n_layers, first;
    unsigned long long index;
    uint8_t* out;            // records back to back: BE8(value) || path, layer by layer, idx then sibling
    uint32_t rec_off[2 * FRI_MAX_LAYERS];   // byte offset of record r = 2*(layer - first) + (0: idx, 1: sibling), from the host
};
void fri_open_one(stark_ctx* ctx, const FriOpenArgs& a);
// Resident opening server for a run of queries on one proof (merkle.cu): `req` = (sequence << 32) | index posted by the
// host, `done` = (sequence << 32) | bytes written by the device; both live in mapped pinned host memory.  a.index and
// a.rec_off are unused (the device computes the record offsets).  Every layer length must divide the first one's.
struct FriServerBox { unsigned long long req; unsigned long long done; };
void fri_open_server_launch(stark_ctx* ctx, const FriOpenArgs& a, FriServerBox* d_box, unsigned long long idle_ns);
// host-side: bytes of the path of leaf idx in a tree of n leaves (32 per level that has a sibling)
size_t merkle_path_len(size_t n, size_t idx);

// ---------------- NTT (ntt.cu) ----------------
// forward (or inverse-root) decimation-in-time: bit-reversed input -> natural output, in `data`.
//   log_pad > 0: the bit-reversed input is the size-2^(log_n-log_pad) array `src`, zero-padded
//   (position q*2^log_pad holds src[q]), and src[q] is first multiplied by scale(bitrev(q)).
void ntt_dit(stark_ctx* ctx, const uint32_t* src, uint32_t* data, unsigned log_n, unsigned log_pad,
             const PowTable* scale, bool inverse_root, size_t batch = 1);
// decimation-in-frequency: natural input -> bit-reversed output, in place.
void ntt_dif(stark_ctx* ctx, uint32_t* data, unsigned log_n, bool inverse_root, size_t batch = 1);
// natural -> natural transform of size 2^log_n >= 2^10 with no permutation sweep (ntt.cu): dst = NTT(src zero-padded
// above src_len), optionally x_j *= in_scale(j) on the way in and X_k *= out_scale(k) on the way out.  `work` is a
// scratch array of 2^log_n words (x batch); src must differ from work and may equal dst (it is consumed by the first pass).
// All pointers 16-byte aligned.
bool ntt_natural_supported(unsigned log_n, const void* src, const void* work, const void* dst);
// batch > 1: `batch` transforms back to back (src_len = batch * 2^log_n, no input scale; out_scale is indexed inside each transform)
void ntt_natural(stark_ctx* ctx, const uint32_t* src, size_t src_len, uint32_t* work, uint32_t* dst, unsigned log_n,
                 bool inverse_root, const PowTable* in_scale, const PowTable* out_scale, size_t batch = 1);
// column-batched transform over data[2^log_n rows][2^col_bits columns] with the strided passes of ntt_natural (log_n >= 10,
// col_bits >= 5): natural rows in, DIGIT-reversed slots out (slot [k_1]..[k_m] holds X[k_1 + R_1 k_2 + ..]), weak values;
// returns the digit widths R_i = 2^bits[i], most significant address digit first
// `from` (optional): the first pass reads element (row, col) from from->src[(row << row_log) + col_off + col] (zero beyond
// src_len) and multiplies it by scale(that index) -- a column slice of a larger row-major array, no staging copy
struct ColumnSource { const uint32_t* src; size_t src_len; unsigned row_log, col_off; const PowTable* scale; };
std::vector<unsigned> ntt_columns_digitrev(stark_ctx* ctx, uint32_t* data, unsigned log_n, unsigned col_bits, bool inverse_root,
                                           const ColumnSource* from = nullptr);
// column-batched DIF over data[2^log_n rows][2^col_bits columns] (col_bits >= 5): natural rows in, bit-reversed out
void ntt_dif_columns(stark_ctx* ctx, uint32_t* data, unsigned log_n, unsigned col_bits, bool inverse_root);
// Blow-up-by-8 forward transform (the LDE / evaluate hot path): dst[8k'+s] = sum_j c_j (base w_N^s)^j w_n^(j k'),
// c_j = c0 * (natural coefficients src[j], or src in bit-reversed order when src_bitrev).  n = 2^log_rows >= 2^10.
bool lde8_supported(unsigned log_rows);
void lde8_forward(stark_ctx* ctx, const uint32_t* src, size_t src_len, bool src_bitrev, uint32_t* dst, unsigned log_rows,
                  uint64_t base, uint64_t c0);
// out[i] = in[bitrev(i)] * scale(scale_on_input_index ? bitrev(i) : i)   (scale optional)
void bitrev_permute(stark_ctx* ctx, const uint32_t* in, uint32_t* out, unsigned log_n, const PowTable* scale,
                    bool scale_by_input_index, size_t batch = 1);
// lo[j] = base^j (j < 2^shift), hi[j] = c0 * base^(j << shift); Montgomery form
struct ScaleTable {
    DevBuf lo, hi;
    PowTable view{};
};
void build_scale_table(stark_ctx* ctx, uint64_t base, uint64_t c0, unsigned log_n, ScaleTable& out);

// ---------------- element-wise / FRI helpers (fri.cu) ----------------
void narrow_u64(stark_ctx* ctx, const uint64_t* in, uint32_t* out, size_t n);      // v % p  (FieldElement::new)
void widen_u32(stark_ctx* ctx, const uint32_t* in, uint64_t* out, size_t n);
void widen_u32_on(stark_ctx* ctx, cudaStream_t s, const uint32_t* in, uint64_t* out, size_t n);
void fill_zero(stark_ctx* ctx, uint32_t* p, size_t n);
// c'[j] = c[2j] + beta*c[2j+1]; result->degree_plus1 = 1 + max{j : c'[j] != 0} (0 for the zero poly)
void poly_degree(stark_ctx* ctx, const uint32_t* c, size_t len, HostResult* result);
// a[i] <- a[i]^-1 (0 stays 0); if num != null: out[i] = num[i] * a[i]^-1
void batch_inverse(stark_ctx* ctx, const uint32_t* a, const uint32_t* num, uint32_t* out, size_t n);
void pointwise_mul(stark_ctx* ctx, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n);
// out[i] = offset * w^i
void coset_domain(stark_ctx* ctx, uint64_t offset, unsigned log_n, uint32_t* out);
// v[i] *= c0 * base^e(i) for a [outer][inner] array: inner = i % inner_len, outer = outer0 + i / inner_len;
//   product mode: e = inner * outer (the four-step twiddle w_N^(n2*k1));  affine: e = inner * inner_stride + outer
void pow_mul(stark_ctx* ctx, uint32_t* v, size_t n, size_t inner_len, size_t outer0, bool product, size_t inner_stride,
             const PowTable& table);
// ---- multi-GPU four-step NTT over peer memory (fourstep.cu) ----
constexpr int MAX_PEERS = 16;
struct PeerPtrs { uint32_t* p[MAX_PEERS]; };
// how a transform left its outputs in its slots: index(slot) = bit reversal (DIF passes), digit reversal (strided
// natural-order passes without the final transposition), or the slot itself (natural order)
struct SlotMap {
    int mode;                // 0: bit reversal, 1: digit reversal, 2: identity
    unsigned total_bits;
    unsigned nd;
    unsigned bits[4];        // digit widths, most significant address digit first
};
// A[n1][n2'] = c[n1*N2 + rank*w + n2'] * offset^(n1*N2 + rank*w + n2')   (zero above `len`)
void fourstep_stage_input(stark_ctx* ctx, const uint32_t* coeffs, size_t len, uint32_t* A, unsigned log_n1, unsigned log_n2,
                          unsigned world, unsigned rank, uint64_t offset);
// rows of A (slot q holds k1 = bitrev(q)) times w_N^(k1*n2), written to their owner: peers[k1 / (N1/G)][k1 % (N1/G)][rank*w + n2']
// (staged: into per-destination chunks [N1/G][w] of a local send buffer instead, for the NCCL transport).
// peer_flags != null: the kernel publishes `epoch` to slot 0 of every peer's flag array when its stores are complete.
void fourstep_twiddle_scatter_rows(stark_ctx* ctx, const uint32_t* A, unsigned log_n1, unsigned log_n2, unsigned world, unsigned rank,
                                   const PeerPtrs& peers, bool staged, void* const* peer_flags, uint32_t epoch, const SlotMap& map);
// X[k1'][slot q2] (k2 = bitrev(q2)) -> natural-order blocks: peers[k2 / (N2/G)][(k2 % (N2/G))*N1 + rank*N1/G + k1']
// (staged: chunks [N2/G][N1/G]); publishes `epoch` to slot 1 when peer_flags != null.
void fourstep_transpose_scatter(stark_ctx* ctx, const uint32_t* X, unsigned log_n1, unsigned log_n2, unsigned world, unsigned rank,
                                const PeerPtrs& peers, bool staged, void* const* peer_flags, uint32_t epoch, const SlotMap& map);
// one-warp kernel that spins until all `world` epoch words of `slot` in this rank's flag array have reached `epoch`
void fourstep_wait(stark_ctx* ctx, const void* own_flags, unsigned slot, unsigned world, uint32_t epoch);
// phase A = stage + column-batched DIF + twiddle/scatter; phase C = row-batched DIF + transpose/scatter
void fourstep_phase_a_launch(stark_ctx* ctx, const uint32_t* coeffs, size_t n_coeffs, unsigned log_n, uint64_t offset, unsigned world,
                             unsigned rank, const PeerPtrs& dst, bool staged, void* const* peer_flags, uint32_t epoch);
void fourstep_phase_c_launch(stark_ctx* ctx, uint32_t* rows, unsigned log_n, unsigned world, unsigned rank, const PeerPtrs& dst,
                             bool staged, void* const* peer_flags, uint32_t epoch);
// plain (unfused) evaluation-space fold of one layer
void fri_fold(stark_ctx* ctx, const LeafSource& src);
void fri_fold_on(stark_ctx* ctx, cudaStream_t s, const LeafSource& src);      // the same on another stream

// STARK-101 FibonacciSq composition on the LDE coset (build-defined)
struct FibSqParams {
    unsigned log_trace, log_blowup;
    uint64_t offset;             // coset offset w
    uint64_t alpha[3];
    uint64_t last_value;         // a[T-2]
};
void fibsq_composition(stark_ctx* ctx, const uint32_t* f_eval, const FibSqParams& prm, uint32_t* cp_eval, size_t start, size_t count);

}  // namespace starkb200
