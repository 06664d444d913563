/*
 * stark_b200.h — C ABI of libstark_b200.so: the B200 (sm_100a) implementation of the data-parallel hot
 * path of RazorClient/Stark-prover (crate `stark-101`).
 *
 * This is the seam a thin Rust FFI shim in src/polynomial, src/merkle and src/fri binds (INTEGRATION.md
 * shows the `extern "C"` block and build.rs).  The reference has no FFI of its own, so every entry
 * point names the reference `pub fn` it stands behind (paths relative to the reference repo).
 *
 * Conventions
 *   - Field elements cross the boundary as the reference stores them: one canonical u64 per
 *     FieldElement<MODULUS> (src/fields/element.rs:7-10).  Inputs are reduced `% modulus` like
 *     FieldElement::new (:13-17).  The modulus must be an odd prime < 2^32 — the reference's own domain
 *     of validity (`pow` multiplies in u64, :45,47).
 *   - Every function returns 0 on success; non-zero = STARK_E_* and stark_last_error() (thread-local)
 *     describes it.  Nothing unwinds across this boundary; the Rust shim maps non-zero to `panic!`,
 *     the reference's error convention (e.g. src/polynomial/ops.rs:143, src/merkle/mod.rs:25).
 *   - Host pointers are caller-owned.  Opaque handles (stark_ctx, stark_vec, stark_tree, stark_fri,
 *     stark_channel) are library-owned and released with their *_destroy function.
 *   - A context is bound to one CUDA device and one modulus and serialises its callers (safe to call
 *     from rayon workers, cf. src/polynomial/interpolation.rs:89-111).
 *   - There is no CPU fallback: without a CUDA device stark_ctx_create fails with STARK_E_CUDA.
 */
#ifndef STARK_B200_H
#define STARK_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STARK_OK 0
#define STARK_E_INVALID 1      /* bad argument (the reference would panic) */
#define STARK_E_CUDA 2         /* CUDA runtime / no device */
#define STARK_E_UNSUPPORTED 3  /* outside the supported domain (modulus >= 2^32, size not dividing p-1, ...) */
#define STARK_E_INTERNAL 4

typedef struct stark_ctx stark_ctx;
typedef struct stark_vec stark_vec;         /* device-resident Vec<FieldElement<M>> */
typedef struct stark_tree stark_tree;       /* MerkleTree<M>            (src/merkle/mod.rs:5-7) */
typedef struct stark_fri stark_fri;         /* FRIProof                 (src/fri/fri_commit.rs:9-13) */
typedef struct stark_channel stark_channel; /* Channel<M>               (src/channel/channel.rs:14-20) */
typedef struct stark_mg stark_mg;           /* a group of ranks, one GPU each (multi-GPU entry points below) */
typedef struct stark_mg_fri stark_mg_fri;   /* FRIProof whose layer 0 is spread over the group */

const char* stark_last_error(void);
const char* stark_version(void);

/* ---- context -------------------------------------------------------------------------------------
 * generator: a generator of F_p^* (5 for p = 3221225473); w_n = generator^((p-1)/n) is the n-th root of
 * unity used for every domain, the convention of SURVEY.md 8(c)/(d).  generator == 0 lets the library
 * pick the smallest one. */
int stark_ctx_create(uint64_t modulus, uint64_t generator, int device, stark_ctx** out);
void stark_ctx_destroy(stark_ctx* ctx);
int stark_ctx_sync(stark_ctx* ctx);
uint64_t stark_ctx_modulus(const stark_ctx* ctx);
uint64_t stark_ctx_generator(const stark_ctx* ctx);
uint64_t stark_ctx_root_of_unity(const stark_ctx* ctx, unsigned log_n);
unsigned stark_ctx_two_adicity(const stark_ctx* ctx);
/* number of kernels launched through this context so far (bench.py's gpu_launches) */
unsigned long long stark_ctx_launch_count(const stark_ctx* ctx);
/* The stream all work of this context is issued on (a cudaStream_t), for CUDA-event timing. */
void* stark_ctx_stream(const stark_ctx* ctx);
/* Measurement aid: while on, every kernel launch is bracketed by a CUDA-event pair on the context's stream.
 * stark_ctx_read_timing syncs, returns per category {0: leaf hashing (incl. the fused FRI fold), 1: node
 * hashing, 2: NTT passes, 3: everything else} the summed kernel milliseconds, the algorithmic work issued
 * (int-ops for 0/1, bytes for 2/3; SURVEY.md 8d figures) and the launch count, and resets the counters. */
int stark_ctx_set_timing(stark_ctx* ctx, int on);
int stark_ctx_read_timing(stark_ctx* ctx, double ms[4], double units[4], unsigned long long launches[4]);
/* Integer-pipe issue peak of this device in 1e12 thread-instructions/s: tops[0] = SHF+LOP3+IADD3 chains (ALU
 * pipe only, what SHA-256's rotates and boolean functions are bound by), tops[1] = the same with one IMAD per
 * three ALU instructions (ALU + FMA pipes).  MEASURED_PEAKS.json has no integer figure. */
int stark_measure_int_peak(stark_ctx* ctx, double tops[2]);
/* Diagnostic: 1e12 chain steps/s of {3 ALU instructions} + {nothing, IMAD, FMA-pipe rotate (mul.wide + mad), mul.hi, 2 IMAD};
 * entries 5..7: three IMAD / IMAD.HI / IMAD.WIDE per step with no ALU-pipe work. */
int stark_measure_pipe_mix(stark_ctx* ctx, double steps_per_s[8]);

/* ---- device vectors ------------------------------------------------------------------------------ */
int stark_vec_upload(stark_ctx* ctx, const uint64_t* host, size_t n, stark_vec** out);
int stark_vec_alloc(stark_ctx* ctx, size_t n, stark_vec** out);                 /* zero-filled */
int stark_vec_download(const stark_vec* v, size_t offset, size_t n, uint64_t* host);
/* copies n canonical u32 values from caller-owned device memory (e.g. the receive buffer of an NCCL exchange) */
int stark_vec_from_device(stark_ctx* ctx, const void* device_u32, size_t n, stark_vec** out);
size_t stark_vec_len(const stark_vec* v);
/* device address of the n canonical u32 values (for NCCL exchanges issued by the caller) */
void* stark_vec_device_ptr(const stark_vec* v);
void stark_vec_destroy(stark_vec* v);

/* ---- polynomial: src/polynomial ------------------------------------------------------------------
 * Domains are power-of-two cosets D[i] = offset * w_n^i in natural order (src/fri/coset_fri.rs:32-36).
 *
 * stark_coset_evaluate   == domain.iter().map(|x| poly.evaluate(x))      ops.rs:76-83 at fri_commit.rs:78
 * stark_coset_interpolate== Polynomial::interpolate(domain, evals)       ops.rs:239-241 -> interpolation.rs:121-152
 *                           (n coefficients, low -> high, NOT trimmed; the shim's Polynomial::new trims)
 * stark_coset_lde        == interpolate on offset_in*<w_n>, evaluate on offset_out*<w_{n*2^log_blowup}>
 * stark_ntt / stark_intt == the same with offset 1, in place
 * stark_batch_inverse    == a.iter().map(|x| x.inverse())                element.rs:54-57 (inverse(0) == 0)
 * stark_quotient_pointwise == num[i] / den[i]                            element.rs:116-122; the evaluation-
 *                           space form of Polynomial::div by a vanishing polynomial (ops.rs:141-191, :412-421)
 */
int stark_ntt(stark_ctx* ctx, uint64_t* inout, unsigned log_n);
int stark_intt(stark_ctx* ctx, uint64_t* inout, unsigned log_n);
int stark_coset_evaluate(stark_ctx* ctx, const uint64_t* coeffs, size_t n_coeffs, unsigned log_n, uint64_t offset,
                         uint64_t* out);
int stark_coset_interpolate(stark_ctx* ctx, const uint64_t* evals, unsigned log_n, uint64_t offset, uint64_t* coeffs_out);
int stark_coset_lde(stark_ctx* ctx, const uint64_t* evals, unsigned log_n, uint64_t offset_in, unsigned log_blowup,
                    uint64_t offset_out, uint64_t* out);
int stark_batch_inverse(stark_ctx* ctx, uint64_t* inout, size_t n);
int stark_quotient_pointwise(stark_ctx* ctx, const uint64_t* num, const uint64_t* den, size_t n, uint64_t* out);
int stark_coset_domain(stark_ctx* ctx, unsigned log_n, uint64_t offset, uint64_t* out);   /* coset_fri.rs:32-36 */
/* device-resident variants (inputs already in HBM; outputs are new vectors) */
int stark_coset_evaluate_dev(stark_ctx* ctx, const stark_vec* coeffs, unsigned log_n, uint64_t offset, stark_vec** out);
int stark_coset_interpolate_dev(stark_ctx* ctx, const stark_vec* evals, uint64_t offset, stark_vec** out);
int stark_coset_lde_dev(stark_ctx* ctx, const stark_vec* evals, uint64_t offset_in, unsigned log_blowup,
                        uint64_t offset_out, stark_vec** out);
int stark_batch_inverse_dev(stark_ctx* ctx, const stark_vec* a, stark_vec** out);
int stark_quotient_pointwise_dev(stark_ctx* ctx, const stark_vec* num, const stark_vec* den, stark_vec** out);

/* Building blocks of the multi-GPU four-step NTT (the all-to-all between them is the caller's
 * NCCL call, see stark-prover_b200/multi_gpu.py):
 *   stark_ntt_batch_dev: len/2^log_m independent size-2^log_m transforms (offset 1, natural order), in place;
 *   stark_pow_mul_dev:   v[i] *= c0 * base^e(i) over a [outer][inner] array with inner = i % inner_len and
 *                        outer = outer0 + i / inner_len;  product != 0: e = inner*outer (the twiddle w_N^(n2*k1)),
 *                        else e = inner*inner_stride + outer (coset scaling of a column-distributed array);
 *                        every e must be < 2^log_table. */
int stark_ntt_batch_dev(stark_ctx* ctx, stark_vec* v, unsigned log_m, int inverse);
int stark_pow_mul_dev(stark_ctx* ctx, stark_vec* v, size_t inner_len, size_t outer0, int product, size_t inner_stride,
                      uint64_t base, uint64_t c0, unsigned log_table);

/* The same four-step NTT with the exchanges written straight into PEER memory over NVLink (fourstep.cu):
 *   stark_peer_alloc   a cudaMalloc'ed, zero-filled vector plus its 64-byte CUDA IPC handle (send it to the peers)
 *   stark_peer_open    maps a peer's handle into this process; the pointer is valid in this context's kernels
 *   stark_fourstep_phase_a  coefficients -> column-batched transforms -> twiddle -> row stores into every peer's
 *                           [N1/G][N2] buffer (peer_rows[s] = rank s's buffer; peer_rows[rank] = own buffer)
 *   stark_fourstep_phase_c  row transforms on the own buffer -> 32x32 transpose -> 128-byte stores into every peer's
 *                           natural-order block (peer_blocks[t])
 * Hand-over between the phases.  peer_flags == NULL: both calls return after their stores are complete and the caller
 * places one barrier between them.  peer_flags[s] = rank s's flag array (stark_peer_alloc of 2 * 16 words, zero-filled,
 * opened on every rank): nothing is synchronised on the host -- phase A publishes `epoch` (1, 2, 3, ... per transform) to
 * every peer when its stores are complete, phase C starts with a device-side wait for all peers' phase-A epochs and
 * publishes its own, and stark_fourstep_wait(own flags, 1, ...) enqueues the wait for the block. */
int stark_peer_alloc(stark_ctx* ctx, size_t n, stark_vec** out, uint8_t handle[64]);
int stark_peer_open(stark_ctx* ctx, const uint8_t handle[64], void** dptr);
int stark_peer_close(stark_ctx* ctx, void* dptr);
int stark_fourstep_phase_a(stark_ctx* ctx, const stark_vec* coeffs, unsigned log_n, uint64_t offset, unsigned world,
                           unsigned rank, void* const* peer_rows, void* const* peer_flags, uint32_t epoch);
int stark_fourstep_phase_c(stark_ctx* ctx, stark_vec* rows, unsigned log_n, unsigned world, unsigned rank,
                           void* const* peer_blocks, void* const* peer_flags, uint32_t epoch);
int stark_fourstep_wait(stark_ctx* ctx, const void* own_flags, unsigned slot, unsigned world, uint32_t epoch);

/* ---- merkle: src/merkle/mod.rs -------------------------------------------------------------------
 * stark_merkle_commit == MerkleTree::new(data)   :10-22   leaf = SHA-256(value.to_be_bytes()), rs_merkle tree
 * stark_merkle_root_hex == MerkleTree::root()    :24-26   64 lowercase hex chars + NUL
 * stark_merkle_open   == the `get_authentication_path(idx)` that src/fri/fri_commit.rs:157 calls and the
 *                        reference never defines: sibling digests bottom -> top, 32 bytes each
 *                        (rs_merkle MerkleProof::to_bytes for one leaf).
 * n == 0 is STARK_E_INVALID (root() would panic on unwrap, :25). */
int stark_merkle_commit(stark_ctx* ctx, const uint64_t* leaves, size_t n, stark_tree** out);
int stark_merkle_commit_dev(stark_ctx* ctx, const stark_vec* leaves, stark_tree** out);
int stark_merkle_root(const stark_tree* t, uint8_t root[32]);
int stark_merkle_root_hex(const stark_tree* t, char out[65]);
size_t stark_merkle_num_leaves(const stark_tree* t);
size_t stark_merkle_depth(const stark_tree* t);
int stark_merkle_open(const stark_tree* t, size_t idx, uint8_t* path, size_t cap, size_t* path_len);
/* digest j of level l >= 1 (level 0 digests are not stored; they are SHA-256 of the leaf) */
int stark_merkle_node(const stark_tree* t, size_t level, size_t j, uint8_t out[32]);
void stark_tree_destroy(stark_tree* t);

/* ---- channel: src/channel/channel.rs (host side; kept here so C/C++ callers have the transcript) -- */
int stark_channel_new(uint64_t modulus, stark_channel** out);                                  /* :24-30 */
void stark_channel_destroy(stark_channel* ch);
int stark_channel_send(stark_channel* ch, const uint8_t* msg, size_t len);                       /* :35-44 */
int stark_channel_receive_random_field_element(stark_channel* ch, uint64_t* out);               /* :47-55 */
int stark_channel_receive_random_int(stark_channel* ch, uint64_t min, uint64_t max, int show_in_proof, uint64_t* out); /* :58-84 */
size_t stark_channel_proof_size(const stark_channel* ch);                                       /* :88-90 */
size_t stark_channel_compressed_proof_size(const stark_channel* ch);                            /* :93-95 */
const char* stark_channel_state(const stark_channel* ch);
size_t stark_channel_proof_len(const stark_channel* ch);
/* message i: returns its length, *data points into the channel's log and stays valid until the next send/receive */
size_t stark_channel_proof_msg(const stark_channel* ch, size_t i, const uint8_t** data);
size_t stark_channel_compressed_len(const stark_channel* ch);
size_t stark_channel_compressed_msg(const stark_channel* ch, size_t i, const uint8_t** data);
/* all proof messages as  u32-LE length || bytes  records; returns total size (out may be NULL) */
size_t stark_channel_proof_flat(const stark_channel* ch, uint8_t* out);

/* ---- FRI: src/fri/fri_commit.rs ------------------------------------------------------------------
 * Step API (the Rust shim keeps its own Channel and drives these):
 *   stark_fri_begin  == lines :78-86   evaluate on the coset, build the tree; root returned
 *   stark_fri_degree == poly.degree    (exact, ops.rs:19-37), the loop condition of :89
 *   stark_fri_fold   == lines :91-103  given beta: next_fri_layer (:53-65) fused with MerkleTree::new (:97)
 *   stark_fri_final  == lines :109-113 the constant that is sent last
 * Whole-loop API with the library's Channel:
 *   stark_fri_commit          == fri_commit(poly, domain, &mut channel)        :72-122
 *   stark_decommit_fri_layers == decommit_fri_layers(index, ..., &mut channel) :137-165
 *   stark_decommit_fri        == decommit_fri(num_queries, max_index, ...)     :168-179
 * The domain is offset*<w_{2^log_n}>; root messages are the 64 ASCII bytes of the hex root
 * (src/fri/fri_verify.rs:24-25). */
int stark_fri_begin(stark_ctx* ctx, const uint64_t* coeffs, size_t n_coeffs, unsigned log_n, uint64_t offset,
                    stark_fri** out, uint8_t root[32]);
int stark_fri_begin_dev(stark_ctx* ctx, const stark_vec* coeffs, unsigned log_n, uint64_t offset, stark_fri** out,
                        uint8_t root[32]);
int stark_fri_degree(const stark_fri* f, long long* degree);
int stark_fri_fold(stark_fri* f, uint64_t beta, uint8_t root[32]);
int stark_fri_final(const stark_fri* f, uint64_t* value, size_t* final_poly_len);
size_t stark_fri_num_layers(const stark_fri* f);
size_t stark_fri_layer_len(const stark_fri* f, size_t k);
int stark_fri_layer_read(const stark_fri* f, size_t k, size_t offset, size_t n, uint64_t* out);
const stark_tree* stark_fri_layer_tree(const stark_fri* f, size_t k);           /* borrowed */
/* Openings of `n_idx` query indices across all layers in one launch.  For each query, for each layer k:
 *   BE8(evals[idx]) || path(idx) || BE8(evals[sib]) || path(sib),  idx = index % len_k, sib = (idx+len_k/2) % len_k
 * written back to back.  *len receives the total; call with out == NULL to size. */
int stark_fri_open(const stark_fri* f, const uint64_t* indices, size_t n_idx, uint8_t* out, size_t cap, size_t* len);
/* Multi-GPU layer 0: the evaluations were produced by the four-step NTT and hashed in leaf
 * ranges on several GPUs.  stark_fri_begin_external adopts the gathered layer and its combined root (no
 * evaluation, no hashing); the folds and later layers are stark_fri_fold as usual.  Layer-0 openings are made
 * on the owning ranks (stark_merkle_open on their subtree + the top levels); stark_fri_open_layers returns the
 * records of layers >= first_layer in the stark_fri_open format. */
int stark_fri_begin_external(stark_ctx* ctx, const stark_vec* coeffs, unsigned log_n, uint64_t offset,
                             const stark_vec* layer0, const uint8_t root0[32], stark_fri** out);
int stark_fri_open_layers(const stark_fri* f, size_t first_layer, const uint64_t* indices, size_t n_idx, uint8_t* out,
                          size_t cap, size_t* len);
void stark_fri_destroy(stark_fri* f);
/* FRIProof.fri_layers BY VALUE (fri_commit.rs:9-13, :117-121 return Vec<Vec<FieldElement>>) without paying for the copy
 * after the commit.  stark_fri_begin_to_host is stark_fri_begin with a host destination: layer 0, and every layer a later
 * stark_fri_fold produces, is also written to layers_out as u64 -- layer k at element offset
 * stark_fri_layer_host_offset(f, k) = the sum of the earlier layers' lengths -- by a second, high-priority stream (widening
 * kernel + copy engine) while the main stream hashes the following layers.  cap = capacity of layers_out in elements
 * (2^(log_n+1) hold every layer).  A layer's copy is enqueued while the main stream works on the next one (the last layer's
 * with stark_fri_final); stark_fri_layers_wait blocks until every layer produced so far has landed;
 * stark_fri_destroy waits as well.  With pinned memory (cudaHostAlloc / cudaHostRegister) the copies are asynchronous;
 * pageable memory works but blocks the calling thread for each copy.  stark_fri_commit_to_host = the whole loop with the
 * library's Channel, complete on return; stark_fri_commit_to_host_async returns as soon as the transcript is complete, the
 * last copies may still be in flight (they finish under the caller's query phase: stark_fri_layers_wait before reading). */
int stark_fri_begin_to_host(stark_ctx* ctx, const uint64_t* coeffs, size_t n_coeffs, unsigned log_n, uint64_t offset,
                            uint64_t* layers_out, size_t cap, stark_fri** out, uint8_t root[32]);
int stark_fri_layers_wait(const stark_fri* f);
size_t stark_fri_layer_host_offset(const stark_fri* f, size_t k);               /* (size_t)-1: no such layer / no host copy */
int stark_fri_commit_to_host(stark_ctx* ctx, const uint64_t* coeffs, size_t n_coeffs, unsigned log_n, uint64_t offset,
                             stark_channel* ch, uint64_t* layers_out, size_t cap, stark_fri** out);
int stark_fri_commit_to_host_async(stark_ctx* ctx, const uint64_t* coeffs, size_t n_coeffs, unsigned log_n, uint64_t offset,
                                   stark_channel* ch, uint64_t* layers_out, size_t cap, stark_fri** out);

int stark_fri_commit(stark_ctx* ctx, const uint64_t* coeffs, size_t n_coeffs, unsigned log_n, uint64_t offset,
                     stark_channel* ch, stark_fri** out);
int stark_fri_commit_dev(stark_ctx* ctx, const stark_vec* coeffs, unsigned log_n, uint64_t offset,
                         stark_channel* ch, stark_fri** out);
int stark_decommit_fri_layers(const stark_fri* f, size_t index, stark_channel* ch);
int stark_decommit_fri(const stark_fri* f, size_t num_queries, size_t max_index, stark_channel* ch);

/* ---- multi-GPU: one process per GPU, one stark_mg per process ------------------------------------
 * What shards: independent trace columns (cfg4), and one big column through the four-step NTT + contiguous leaf ranges
 * (cfg5) -- layer 0 of a FRI proof and, on more than one GPU, every later layer of >= 2^20 leaves (each rank folds the
 * whole layer itself, hashes its leaf range, 32-byte subtree roots are gathered); the smaller trees and the channel
 * stay on rank 0, and no data is partitioned except for leaf hashing (north-star: "nothing else is partitioned").  The group owns an NCCL communicator (bound at run time, libnccl.so.2): rank 0 calls
 * stark_mg_unique_id, hands the 128 bytes to the other ranks by whatever means the host program has (MPI, a file, its
 * own RPC), and every rank calls stark_mg_create; stark_mg_adopt takes an existing ncclComm_t instead.  Calls on a group
 * are collective: every rank makes the same calls in the same order.  world must be a power of two <= 16 for the
 * four-step entry points. */
int stark_mg_unique_id(uint8_t id[128]);
int stark_mg_create(stark_ctx* ctx, const uint8_t id[128], unsigned rank, unsigned world, stark_mg** out);
int stark_mg_adopt(stark_ctx* ctx, void* nccl_comm, unsigned rank, unsigned world, stark_mg** out);
void stark_mg_destroy(stark_mg* mg);
unsigned stark_mg_rank(const stark_mg* mg);
unsigned stark_mg_world(const stark_mg* mg);
int stark_mg_barrier(stark_mg* mg);
/* cfg4.  Column c = columns[c][0 .. 2^log_rows) (host memory, read only on rank c % world; pinned memory lets the next
 * upload run under the current column's hashing): trace column on offset_in*<g> -> coset LDE on offset_out*<h>
 * (interpolate + evaluate, ops.rs:239 / :76) -> MerkleTree::new (merkle/mod.rs:10).  roots: n_cols * 32 bytes, all
 * columns, identical on every rank.  ldes / trees (optional arrays of n_cols entries): the owner's device-resident LDE
 * and tree of each of its columns (NULL for the others), to be destroyed by the caller. */
int stark_mg_commit_columns(stark_mg* mg, size_t n_cols, const uint64_t* const* columns, unsigned log_rows, uint64_t offset_in,
                            unsigned log_blowup, uint64_t offset_out, uint8_t* roots, stark_vec** ldes, stark_tree** trees);
/* cfg5.  Evaluations of `coeffs` (held by every rank) on offset*<w_{2^log_n}> through the four-step NTT:
 * transport 0 = all-to-all over NCCL, 1 = the producing kernels store into peer memory over NVLink with device-side epoch
 * flags between the phases (no host synchronisation inside the transform).  *block = this rank's natural-order range
 * [rank*N/G, (rank+1)*N/G); it aliases a buffer of the group that the next transform of the same size overwrites. */
int stark_mg_fourstep_lde(stark_mg* mg, const stark_vec* coeffs, unsigned log_n, uint64_t offset, int transport, stark_vec** block);
/* MerkleTree::new over a column held in contiguous leaf ranges: each rank hashes its range (an exact subtree), the
 * subtree roots are gathered and the top levels finished on every rank.  subtree_roots: optional, world * 32 bytes. */
int stark_mg_commit_leaf_ranges(stark_mg* mg, const stark_vec* block, stark_tree** subtree, uint8_t root[32], uint8_t* subtree_roots);
/* fri_commit (fri_commit.rs:72-122) / decommit_fri (:168-179) with layer 0 -- and the later large layers -- hashed in
 * leaf ranges; `ch` is used on rank 0 only and receives the single-GPU transcript byte for byte.  Environment:
 * STARK_MG_FRI_SHARD=0 leaves the layers >= 1 on rank 0, STARK_MG_FRI_SHARD_MIN_LOG (default 20) moves the threshold. */
int stark_mg_fri_commit(stark_mg* mg, const stark_vec* coeffs, unsigned log_n, uint64_t offset, int transport, stark_channel* ch,
                        stark_mg_fri** out);
int stark_mg_decommit_fri(stark_mg_fri* f, size_t num_queries, size_t max_index, stark_channel* ch);
/* stark101_prove over the group (BASELINE cfg5, "end-to-end prove ... using four-step NTT"): trace LDE through the
 * four-step transform, commitments of f and of the composition polynomial in leaf ranges, the composition polynomial on
 * each rank's own range (a halo of 2 * blowup values from the next rank), the FRI layers as in stark_mg_fri_commit.
 * The transcript in `ch` (rank 0) is stark101_prove's, byte for byte. */
int stark_mg_stark101_prove(stark_mg* mg, uint64_t a1, unsigned log_trace, unsigned log_blowup, size_t num_queries, int transport,
                            stark_channel* ch);
const stark_fri* stark_mg_fri_proof(const stark_mg_fri* f);      /* rank 0: every layer's values, the small layers' trees (borrowed); elsewhere NULL or the replicated large layers */
const stark_tree* stark_mg_fri_subtree(const stark_mg_fri* f);   /* this rank's subtree of layer 0 (borrowed) */
void stark_mg_fri_destroy(stark_mg_fri* f);

/* ---- verification (host side; completes what src/fri/fri_verify.rs sketches) ------------------------
 * stark_merkle_verify == the `MerkleTree::validate` that fri_verify.rs:109,137 calls and the reference never defines
 *                        (rs_merkle MerkleProof::verify for one leaf).
 * stark_fri_verify    == verify_fri (fri_verify.rs:12-177) with the fold-consistency check that is commented out there
 *                        (:153-170): replays the flattened proof (stark_channel_proof_flat) against a fresh channel.
 * Both set *ok to 1/0 and return STARK_OK unless an argument is unusable; `reason` (>= 160 bytes, optional) gets
 * the first failure. */
int stark_merkle_verify(const uint8_t root[32], size_t n_leaves, size_t idx, uint64_t value, const uint8_t* path,
                        size_t path_len, int* ok);
int stark_fri_verify(const uint8_t* proof_flat, size_t proof_len, uint64_t modulus, uint64_t generator, unsigned log_n,
                     uint64_t offset, size_t num_queries, size_t max_index, unsigned log_degree_bound, int* ok, char* reason);
/* log_degree_bound: the committed polynomial is claimed to have at most 2^log_degree_bound coefficients (<= 2^log_n);
 * a proof with more than log_degree_bound folds is rejected -- the layer count is the prover's choice, and without this
 * cap any function folds down to a one-point layer that trivially "equals the final constant"
 * (`expected_num_layers`, fri_verify.rs:15).
 *
 * Verifier of stark101_prove's transcript; public input = the claimed a_{T-2}.  The transcript opens with the statement
 * (modulus, generator, log_trace, log_blowup, num_queries, claimed a_{T-2}: 8 big-endian bytes each) so that every
 * challenge depends on it.  Adds to the FRI checks: the degree bound 2^log_trace that the statement implies for the
 * composition polynomial, the three trace openings per query authenticate against the trace root, and layer 0 of the
 * FRI equals the composition polynomial computed from them. */
int stark101_verify(const uint8_t* proof_flat, size_t proof_len, uint64_t modulus, uint64_t generator, uint64_t claimed_last,
                    unsigned log_trace, unsigned log_blowup, size_t num_queries, int* ok, char* reason);

/* ---- prover (build-defined: src/prover, src/trace, src/composition are empty in the reference) ----
 * STARK-101 FibonacciSq statement a0 = 1, a1 = `a1`, a_{n+2} = a_{n+1}^2 + a_n^2 over 2^log_trace - 1 rows;
 * protocol and transcript order in DESIGN.md "cfg1". */
int stark101_prove(stark_ctx* ctx, uint64_t a1, unsigned log_trace, unsigned log_blowup, size_t num_queries,
                   stark_channel* ch);
/* Pieces of the same prover for callers that spread it over several GPUs (stark-prover_b200/multi_gpu.py,
 * stark101_prove_multi): the T coefficients of the trace polynomial f (the top one is zero) together with a_{T-2};
 * and the composition polynomial on the points start .. start+count-1 of the coset generator*<h>, where f_block
 * holds f on that range followed by the next 2*blowup points (count == N: the whole coset, no halo). */
int stark101_trace_poly(stark_ctx* ctx, uint64_t a1, unsigned log_trace, stark_vec** coeffs, uint64_t* last_value);
int stark101_composition_range(stark_ctx* ctx, const stark_vec* f_block, size_t start, size_t count, const uint64_t alpha[3],
                               uint64_t last_value, unsigned log_trace, unsigned log_blowup, stark_vec** cp_block);

#ifdef __cplusplus
}
#endif
#endif
