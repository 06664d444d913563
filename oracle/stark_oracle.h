/*
 * stark_oracle.h — CPU oracle for the STARK hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  Nothing under stark-prover_b200/ links, imports or calls it.
 *
 * It restates, in plain C, the arithmetic of RazorClient/Stark-prover (crate `stark-101`):
 *   field       src/fields/element.rs:13-17,38-61,72-147
 *   polynomial  src/polynomial/ops.rs:19-37,47-60,76-83,114-138,141-191
 *   lagrange    src/polynomial/interpolation.rs:9-23,46-78,121-152
 *   merkle      src/merkle/mod.rs:10-26            (+ rs_merkle 1.4.2 tree rule, sha2 0.10.8)
 *   channel     src/channel/channel.rs:24-95       (+ sha256 1.5.0, ruint 1.12.3 U256 semantics)
 *   fri         src/fri/fri_commit.rs:18-24,32-65,72-122,137-179, src/fri/coset_fri.rs:32-36
 *
 * Two tiers with bit-identical results:
 *   literal  — the reference's own algorithms (Horner per point, long division, Lagrange,
 *              Fermat inverse, coefficient-space fold).  O(N*d): small sizes only.
 *   fast     — radix-2 NTT, Montgomery-trick inverse, evaluation-space fold, OpenMP.
 *              Checked against literal wherever literal finishes; used for full-size parity and
 *              as the timed CPU baseline.
 *
 * PINNING.  The reference is Rust (nightly) and cannot be built in this image, and the crates
 * that decide the hash/tree/transcript rules (rs_merkle 1.4.2, sha2 0.10.8, sha256 1.5.0,
 * alloy 0.11.1 / ruint 1.12.3) are not vendored under /root/reference.  Field and polynomial
 * results are pinned by the reference's own mod-7 known-answer tests (tests/golden/ref_kat.json,
 * each with its file:line).  SHA-256 is pinned by the FIPS 180-4 vectors and Python hashlib.
 * The rs_merkle tree rule (pairwise SHA-256, leaves not re-hashed, lone node promoted) is additionally checked
 * against the root rs_merkle publishes for the leaves "a".."f" (tests/golden/spec_anchors.json).
 * Merkle roots over field elements, channel states, FRI layers and openings have no reference test or
 * fixture: for those rows PARITY IS UNPINNED — the oracle is a restatement of the published rules.
 */
#ifndef STARK_ORACLE_H
#define STARK_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------- field: element.rs ---------------- */
uint64_t or_fe_new(uint64_t v, uint64_t M);                 /* :13-17  */
uint64_t or_fe_add(uint64_t a, uint64_t b, uint64_t M);     /* :72-78  */
uint64_t or_fe_sub(uint64_t a, uint64_t b, uint64_t M);     /* :86-92  */
uint64_t or_fe_mul(uint64_t a, uint64_t b, uint64_t M);     /* :102-108 */
uint64_t or_fe_neg(uint64_t a, uint64_t M);                 /* :130-136 */
uint64_t or_fe_pow(uint64_t a, uint64_t e, uint64_t M);     /* :38-51 (u64 products, as written) */
uint64_t or_fe_inverse(uint64_t a, uint64_t M);             /* :54-57  inverse(0) == 0 */
uint64_t or_fe_div(uint64_t a, uint64_t b, uint64_t M);     /* :116-122 */
void     or_fe_to_bytes(uint64_t a, uint8_t out[8]);        /* :59-61 big-endian */
uint64_t or_fe_from_i128(int64_t hi_sign_ext_lo, uint64_t M); /* :138-147 (i64 range is enough) */

/* ---------------- polynomial: ops.rs / interpolation.rs (literal tier) ---------------- */
/* Coefficients low->high.  Every function returns the trimmed length (degree+1; 0 = zero poly). */
size_t   or_poly_trim(const uint64_t* c, size_t len);                                  /* ops.rs:19-37 */
uint64_t or_poly_evaluate(const uint64_t* c, size_t len, uint64_t x, uint64_t M);      /* ops.rs:76-83 */
size_t   or_poly_add(const uint64_t* a, size_t al, const uint64_t* b, size_t bl, uint64_t* out, uint64_t M);
size_t   or_poly_sub(const uint64_t* a, size_t al, const uint64_t* b, size_t bl, uint64_t* out, uint64_t M);
size_t   or_poly_mul(const uint64_t* a, size_t al, const uint64_t* b, size_t bl, uint64_t* out, uint64_t M); /* :114-138; out has al+bl-1 slots */
/* ops.rs:141-191.  q has max(al-bl+1,1) slots, r has al slots.  returns 0, or -1 for division by zero poly */
int      or_poly_div_rem(const uint64_t* a, size_t al, const uint64_t* b, size_t bl,
                         uint64_t* q, size_t* ql, uint64_t* r, size_t* rl, uint64_t M);
size_t   or_poly_from_roots(const uint64_t* roots, size_t n, uint64_t* out, uint64_t M);   /* interpolation.rs:9-23; out has n+1 slots */
/* interpolation.rs:121-152 via :46-78.  out has n slots.  returns trimmed len, or (size_t)-1 on error */
size_t   or_poly_interpolate(const uint64_t* xs, const uint64_t* ys, size_t n, uint64_t* out, uint64_t M);
/* L_i for i in [0,n): out is n*n row-major, row i = L_i padded with zeros (interpolation.rs:46-78) */
int      or_lagrange_basis(const uint64_t* xs, size_t n, uint64_t* out, uint64_t M);

/* ---------------- SHA-256 (FIPS 180-4; sha2 0.10.8 / sha256 1.5.0) ---------------- */
void or_sha256(const uint8_t* msg, size_t len, uint8_t out[32]);
/* 0 = portable C rounds, 1 = x86 SHA-NI when the CPU has it (what sha2 0.10.8 does at run time) */
void or_sha256_set_accel(int on);
int  or_sha256_accel_active(void);

/* ---------------- merkle: merkle/mod.rs + rs_merkle 1.4.2 ---------------- */
typedef struct or_tree or_tree;
or_tree* or_merkle_new(const uint64_t* leaves, size_t n);     /* mod.rs:10-22; n==0 -> NULL (root() would panic) */
void     or_merkle_free(or_tree* t);
size_t   or_merkle_num_leaves(const or_tree* t);
size_t   or_merkle_depth(const or_tree* t);                    /* number of levels above the leaves */
void     or_merkle_root(const or_tree* t, uint8_t out[32]);
void     or_merkle_root_hex(const or_tree* t, char out[65]);   /* mod.rs:24-26 lowercase hex */
/* level 0 = leaf digests; node j of level l */
int      or_merkle_node(const or_tree* t, size_t level, size_t j, uint8_t out[32]);
/* authentication path of leaf idx: sibling digests bottom->top, concatenated (rs_merkle
 * MerkleProof::to_bytes for one leaf; a level where the node has no sibling contributes nothing).
 * Build-defined: fri_commit.rs:157 calls get_authentication_path, which the reference never defines. */
size_t   or_merkle_path(const or_tree* t, size_t idx, uint8_t* out /* >= 32*depth */);
/* verify a path produced above (rs_merkle MerkleProof::verify for one leaf) */
int      or_merkle_verify(const uint8_t root[32], size_t n_leaves, size_t idx, uint64_t value,
                          const uint8_t* path, size_t path_len);
/* the tree rule alone over caller-supplied leaf digests (rs_merkle from_leaves + root), for published vectors */
void     or_merkle_root_from_digests(const uint8_t* digests, size_t n, uint8_t out[32]);
/* root only, multi-threaded, no retained tree (CPU baseline) */
void     or_merkle_root_only(const uint64_t* leaves, size_t n, uint8_t out[32]);

/* ---------------- channel: channel.rs ---------------- */
typedef struct or_channel or_channel;
or_channel* or_channel_new(uint64_t M);                                        /* :24-30 */
void        or_channel_free(or_channel* c);
void        or_channel_send(or_channel* c, const uint8_t* msg, size_t len);    /* :35-44 */
uint64_t    or_channel_receive_random_field_element(or_channel* c);            /* :47-55 */
uint64_t    or_channel_receive_random_int(or_channel* c, uint64_t min, uint64_t max, int show); /* :58-84 */
size_t      or_channel_proof_size(const or_channel* c);                        /* :88-90 */
size_t      or_channel_compressed_proof_size(const or_channel* c);             /* :93-95 */
const char* or_channel_state(const or_channel* c);
size_t      or_channel_proof_len(const or_channel* c);                         /* number of messages */
size_t      or_channel_proof_msg(const or_channel* c, size_t i, const uint8_t** data);
size_t      or_channel_compressed_len(const or_channel* c);
size_t      or_channel_compressed_msg(const or_channel* c, size_t i, const uint8_t** data);
/* all proof messages as  u32-LE length || bytes  records; returns total size (call with out=NULL to size) */
size_t      or_channel_proof_flat(const or_channel* c, uint8_t* out);

/* ---------------- domains: coset_fri.rs ---------------- */
void or_coset_domain(uint64_t offset, uint64_t omega, size_t n, uint64_t* out, uint64_t M);  /* :32-36 */
void or_next_fri_domain(const uint64_t* d, size_t n, uint64_t* out, uint64_t M);             /* fri_commit.rs:18-24 */
size_t or_next_fri_polynomial(const uint64_t* c, size_t len, uint64_t beta, uint64_t* out, uint64_t M); /* :32-50 */

/* ---------------- fast tier primitives ---------------- */
/* omega_n = g^((M-1)/n) for a generator g of F_M^* */
uint64_t or_root_of_unity(uint64_t generator, unsigned log_n, uint64_t M);
/* natural order in, natural order out, in place; inverse includes the 1/n scaling */
int  or_ntt(uint64_t* a, unsigned log_n, uint64_t omega, uint64_t M);
int  or_intt(uint64_t* a, unsigned log_n, uint64_t omega, uint64_t M);
/* evaluate coeffs (len <= 2^log_n) on offset*<omega>, natural order == map(evaluate) over the coset */
int  or_coset_evaluate(const uint64_t* c, size_t len, unsigned log_n, uint64_t offset, uint64_t omega,
                       uint64_t* out, uint64_t M);
/* coefficients of the interpolant through (offset*omega^i, evals[i]); == Polynomial::interpolate */
int  or_coset_interpolate(const uint64_t* evals, unsigned log_n, uint64_t offset, uint64_t omega,
                          uint64_t* out, uint64_t M);
/* Montgomery-trick inverse; zeros map to zero (element.rs:54-57 semantics) */
void or_batch_inverse(uint64_t* a, size_t n, uint64_t M);
/* evaluation-space fold of one layer: out has n/2 slots */
void or_fri_fold_evals(const uint64_t* e, size_t n, uint64_t beta, uint64_t offset, uint64_t omega,
                       uint64_t* out, uint64_t M);

/* ---------------- FRI commit / decommit: fri_commit.rs ---------------- */
typedef struct or_fri_proof or_fri_proof;   /* FRIProof :9-13 */
/* literal: Horner over the explicit domain, coefficient fold (fri_commit.rs:72-122) */
or_fri_proof* or_fri_commit_literal(const uint64_t* coeffs, size_t len, const uint64_t* domain, size_t n,
                                    or_channel* ch, uint64_t M);
/* fast: same transcript via NTT + evaluation-space fold; domain = offset*<omega>, n = 2^log_n */
or_fri_proof* or_fri_commit_fast(const uint64_t* coeffs, size_t len, unsigned log_n, uint64_t offset,
                                 uint64_t omega, or_channel* ch, uint64_t M);
/* as fast but builds no retained trees (root-only hashing): the timed CPU baseline for the commit phase */
int           or_fri_commit_fast_rootonly(const uint64_t* coeffs, size_t len, unsigned log_n, uint64_t offset,
                                 uint64_t omega, or_channel* ch, uint64_t M);
void     or_fri_free(or_fri_proof* p);
size_t   or_fri_num_layers(const or_fri_proof* p);
size_t   or_fri_layer_len(const or_fri_proof* p, size_t k);
const uint64_t* or_fri_layer(const or_fri_proof* p, size_t k);
const or_tree*  or_fri_tree(const or_fri_proof* p, size_t k);
size_t   or_fri_final_poly(const or_fri_proof* p, uint64_t* out /* 1 slot */);   /* returns trimmed len (0 or 1) */
/* fri_commit.rs:137-165 / :168-179 (auth path = or_merkle_path) */
void or_decommit_fri_layers(size_t index, const or_fri_proof* p, or_channel* ch);
void or_decommit_fri(size_t num_queries, size_t max_index, const or_fri_proof* p, or_channel* ch);

/* ---------------- STARK-101 FibonacciSq prover (build-defined rows of SURVEY 8f) ---------------- */
/* trace a0=1, a1=x, a_{n+2}=a_{n+1}^2+a_n^2 */
void or_fibsq_trace(uint64_t a1, size_t rows, uint64_t* out, uint64_t M);
/* Full transcript; see DESIGN.md "cfg1".  literal!=0 uses Horner/Lagrange/long-division everywhere. */
int  or_stark101_prove(uint64_t a1, unsigned log_trace /*10*/, unsigned log_blowup /*3*/, uint64_t generator,
                       size_t num_queries, int literal, or_channel* ch, uint64_t M);

/* threads the fast tier will use (OpenMP) */
int  or_num_threads(void);
void or_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
