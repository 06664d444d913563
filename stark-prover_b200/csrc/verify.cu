// verify.cu — host-side verification of what the device produced.
//
// The reference calls `MerkleTree::validate` (src/fri/fri_verify.rs:109,137) and sketches `verify_fri`
// (src/fri/fri_verify.rs:12-177) but defines neither: the Merkle check does not exist and the fold-consistency
// check is commented out (:153-170).  Both are completed here on the CPU (a verifier is sequential hashing and a
// few field operations per query; nothing in it is data parallel), reading the transcript in the order
// fri_commit / decommit_fri wrote it (src/fri/fri_commit.rs:72-179):
//
//   proof = root_0, [beta_k, root_k] for k = 1..L-1, final(8 B), then per query:
//           index(8 B), and per layer k: elem(idx) , path(idx), elem(sib), path(sib)
//
// Roots travel as 64 ASCII hex bytes (fri_verify.rs:24-25), elements as 8 big-endian bytes (:58,99), the
// sibling of idx in a layer of n values is idx + n/2 (:136).
#include <string.h>

#include <array>
#include <functional>

#include "../../include/stark_b200.h"
#include "handles.hpp"

using namespace starkb200;
namespace starkb200 { void api_set_error(const std::string& s); }

namespace {

struct Fp {
    uint64_t p;
    uint64_t add(uint64_t a, uint64_t b) const { return (a + b) % p; }
    uint64_t sub(uint64_t a, uint64_t b) const { return (p + a - b) % p; }
    uint64_t mul(uint64_t a, uint64_t b) const { return h_mul(a, b, p); }
    uint64_t inv(uint64_t a) const { return h_inv(a, p); }
};

void leaf_digest(uint64_t v, uint8_t out[32]) {
    uint8_t b[8];
    be8(v, b);
    HostSha256::digest(b, 8, out);
}
void node_digest(const uint8_t* l, const uint8_t* r, uint8_t out[32]) {
    uint8_t cat[64];
    memcpy(cat, l, 32); memcpy(cat + 32, r, 32);
    HostSha256::digest(cat, 64, out);
}
// rs_merkle MerkleProof::verify for one leaf: sibling digests bottom -> top; levels without a sibling are skipped
bool merkle_verify(const uint8_t root[32], size_t n_leaves, size_t idx, uint64_t value, const uint8_t* path, size_t path_len) {
    if (n_leaves == 0 || idx >= n_leaves) return false;
    uint8_t cur[32], nxt[32];
    leaf_digest(value, cur);
    size_t j = idx, m = n_leaves, used = 0;
    while (m > 1) {
        if ((j ^ 1) < m) {
            if (used + 32 > path_len) return false;
            if (j & 1) node_digest(path + used, cur, nxt); else node_digest(cur, path + used, nxt);
            memcpy(cur, nxt, 32); used += 32;
        }
        j >>= 1; m = (m + 1) / 2;
    }
    return used == path_len && memcmp(cur, root, 32) == 0;
}
bool parse_hex_root(const std::vector<uint8_t>& m, uint8_t out[32]) {
    if (m.size() != 64) return false;
    auto nib = [](uint8_t c, int* v) { if (c >= '0' && c <= '9') { *v = c - '0'; return true; } if (c >= 'a' && c <= 'f') { *v = c - 'a' + 10; return true; } return false; };
    for (int i = 0; i < 32; i++) {
        int a, b;
        if (!nib(m[2 * i], &a) || !nib(m[2 * i + 1], &b)) return false;
        out[i] = (uint8_t)(a * 16 + b);
    }
    return true;
}
bool parse_be8(const std::vector<uint8_t>& m, uint64_t* v) {
    if (m.size() != 8) return false;
    *v = 0;
    for (int i = 0; i < 8; i++) *v = (*v << 8) | m[i];
    return true;
}

using Msgs = std::vector<std::vector<uint8_t>>;
// called once per query, after the index is drawn and before the FRI layers are read: may consume messages
// (trace openings) and returns the value layer 0 must show at `idx` (or no constraint)
using QueryHook = std::function<bool(size_t idx, const Msgs& msgs, size_t& pos, Channel& ch, uint64_t* expect0, bool* has_expect, std::string* why)>;

bool split_records(const uint8_t* flat, size_t len, Msgs& msgs, std::string* why) {
    for (size_t o = 0; o < len;) {
        if (o + 4 > len) { *why = "truncated record header"; return false; }
        size_t n = (size_t)flat[o] | ((size_t)flat[o + 1] << 8) | ((size_t)flat[o + 2] << 16) | ((size_t)flat[o + 3] << 24);
        if (o + 4 + n > len) { *why = "truncated record"; return false; }
        msgs.emplace_back(flat + o + 4, flat + o + 4 + n);
        o += 4 + n;
    }
    return true;
}

// fri_commit + decommit_fri as seen by a verifier (fri_verify.rs:12-177, completed); starts at msgs[pos] with the
// channel in the state the prover had when it called fri_commit
//
// Degree bound.  The number of layers is the prover's choice (fri_commit.rs:89 folds until the polynomial is constant),
// so it is the VERIFIER that has to cap it: a polynomial with at most 2^log_degree_bound coefficients is constant after
// log_degree_bound folds.  Without the cap a prover may fold any function down to a one-point layer, where "the last
// layer equals the final constant" is vacuous, and every check passes (the reference's draft takes
// `expected_num_layers`, fri_verify.rs:15, for the same reason).  With it the last layer keeps
// 2^(log_n - log_degree_bound) points -- the blow-up -- that must all show the same constant.
bool verify_fri_core(const Msgs& msgs, size_t& pos, Channel& ch, const Fp& F, uint64_t generator, unsigned log_n, uint64_t offset,
                     size_t num_queries, size_t max_index, unsigned log_degree_bound, const QueryHook& hook, std::string* why) {
    auto fail = [&](const std::string& w) { *why = w; return false; };
    if (log_n > 40 || ((F.p - 1) & ((((uint64_t)1) << log_n) - 1)) != 0) return fail("2^log_n does not divide p-1");
    if (offset % F.p == 0) return fail("zero coset offset");
    if (log_degree_bound > log_n) return fail("degree bound larger than the domain");
    const size_t n0 = (size_t)1 << log_n;
    auto next = [&]() -> const std::vector<uint8_t>* { return pos < msgs.size() ? &msgs[pos++] : nullptr; };

    // ---- commit phase: roots and betas (fri_commit.rs:86-103), then the final constant (:109-114)
    std::vector<std::array<uint8_t, 32>> roots;
    std::vector<uint64_t> betas;
    const std::vector<uint8_t>* m = next();
    std::array<uint8_t, 32> r{};
    if (!m || !parse_hex_root(*m, r.data())) return fail("first FRI message is not a 64-character hex root");
    roots.push_back(r);
    ch.send(m->data(), m->size());
    uint64_t final_value = 0;
    for (;;) {
        m = next();
        if (!m) return fail("transcript ends inside the commit phase");
        if (m->size() == 8 && pos < msgs.size() && msgs[pos].size() == 64) {           // beta_k followed by root_k
            uint64_t beta, want;
            parse_be8(*m, &beta);
            if (!ch.receive_random_field_element(&want) || want != beta % F.p) return fail("beta does not follow from the transcript");
            betas.push_back(want);
            m = next();
            if (!parse_hex_root(*m, r.data())) return fail("malformed layer root");
            roots.push_back(r);
            ch.send(m->data(), m->size());
        } else if (m->size() == 8) {                                                    // the final constant
            parse_be8(*m, &final_value);
            ch.send(m->data(), 8);
            break;
        } else {
            return fail("unexpected message in the commit phase");
        }
    }
    const size_t L = roots.size();
    if (L - 1 > log_degree_bound) return fail("more FRI layers than the degree bound allows (" + std::to_string(L - 1) + " folds, bound " +
                                              std::to_string(log_degree_bound) + ")");
    if (final_value >= F.p) return fail("final constant is not a canonical field element");

    // ---- query phase (fri_commit.rs:137-179)
    const uint64_t inv2 = F.inv(2 % F.p);
    const uint64_t w0 = h_pow(generator % F.p, (F.p - 1) >> log_n, F.p);
    for (size_t q = 0; q < num_queries; q++) {
        uint64_t idx_want;
        if (!ch.receive_random_int(0, max_index, true, &idx_want)) return fail("cannot draw a query index");
        m = next();
        uint64_t idx_msg;
        if (!m || !parse_be8(*m, &idx_msg) || idx_msg != idx_want) return fail("query index does not follow from the transcript");
        uint64_t expect0 = 0;
        bool has_expect = false;
        if (hook && !hook((size_t)idx_want, msgs, pos, ch, &expect0, &has_expect, why)) return false;
        uint64_t prev_a = 0, prev_b = 0;      // opened pair of the previous layer, ordered (j, j + n/2)
        size_t prev_j = 0;
        uint64_t off_k = offset % F.p, w_k = w0;
        for (size_t k = 0; k < L; k++) {
            const size_t n = n0 >> k;
            if (n == 0) return fail("layer of size zero");
            const size_t i = (size_t)idx_want % n, s = (i + n / 2) % n;
            uint64_t lone = 0;
            if (n == 1) {                                                               // :147-149 sends the constant first
                m = next();
                if (!m || !parse_be8(*m, &lone)) return fail("missing the length-1 layer element");
                ch.send(m->data(), 8);
            }
            uint64_t val[2];
            const size_t which[2] = {i, s};
            for (int t = 0; t < 2; t++) {
                const std::vector<uint8_t>* me = next();
                const std::vector<uint8_t>* mp = next();
                if (!me || !mp || !parse_be8(*me, &val[t])) return fail("truncated opening");
                if (val[t] >= F.p) return fail("opened value is not canonical");
                if (!merkle_verify(roots[k].data(), n, which[t], val[t], mp->data(), mp->size()))
                    return fail("authentication path does not lead to the layer root (layer " + std::to_string(k) + ")");
                ch.send(me->data(), me->size());
                ch.send(mp->data(), mp->size());
            }
            if (n == 1 && lone != val[0]) return fail("the length-1 layer element differs from the opened value");
            if (k == 0 && has_expect && val[0] != expect0) return fail("layer 0 does not equal the composition polynomial at the queried point");
            const size_t j = i % (n / 2 ? n / 2 : 1);
            const uint64_t a = (i < n / 2 || n == 1) ? val[0] : val[1], b = (i < n / 2 || n == 1) ? val[1] : val[0];
            if (k > 0) {
                // e_k[j'] == (a+b)/2 + beta_k (a-b) / (2 D_{k-1}[j]),  D_{k-1}[j] = off * w^j   (fri_verify.rs:153-170, completed)
                uint64_t d = F.mul(off_k, h_pow(w_k, prev_j, F.p));
                uint64_t folded = F.add(F.mul(F.add(prev_a, prev_b), inv2),
                                        F.mul(F.mul(betas[k - 1], F.sub(prev_a, prev_b)), F.inv(F.mul(2 % F.p, d))));
                if (folded != val[0]) return fail("layer " + std::to_string(k) + " is not the fold of layer " + std::to_string(k - 1) + " at the queried point");
                off_k = F.mul(off_k, off_k); w_k = F.mul(w_k, w_k);
            }
            prev_a = a; prev_b = b; prev_j = j;
            if (k + 1 == L && (val[0] != final_value || val[1] != final_value)) return fail("last layer does not equal the final constant");
        }
    }
    return true;
}

}  // namespace

extern "C" int stark_merkle_verify(const uint8_t root[32], size_t n_leaves, size_t idx, uint64_t value, const uint8_t* path,
                                   size_t path_len, int* ok) {
    if (!root || !ok || (!path && path_len)) { api_set_error("merkle_verify: null argument"); return ST_INVALID; }
    *ok = merkle_verify(root, n_leaves, idx, value, path, path_len) ? 1 : 0;
    return ST_OK;
}

static int finish(bool good, const std::string& why, int* ok, char* reason) {
    *ok = good ? 1 : 0;
    if (reason) { strncpy(reason, good ? "" : why.c_str(), 159); reason[159] = 0; }
    return ST_OK;
}

// Replays a proof — the `proof` messages of a Channel that ran fri_commit + decommit_fri, flattened as
// u32-LE length || bytes records (stark_channel_proof_flat) — against a fresh channel.
// *ok = 1 iff every root/beta/index is the one the transcript dictates, every opened value authenticates
// against its layer root, every layer is the fold of the previous one at the queried points, and the last
// layer equals the final constant.  `reason` (optional, >= 160 bytes) receives the first failure.
extern "C" int stark_fri_verify(const uint8_t* proof_flat, size_t proof_len, uint64_t modulus, uint64_t generator, unsigned log_n,
                                uint64_t offset, size_t num_queries, size_t max_index, unsigned log_degree_bound, int* ok, char* reason) {
    if (!proof_flat || !ok || modulus < 3) { api_set_error("fri_verify: bad argument"); return ST_INVALID; }
    Msgs msgs;
    std::string why;
    if (!split_records(proof_flat, proof_len, msgs, &why)) return finish(false, why, ok, reason);
    const Fp F{modulus};
    Channel ch(F.p);
    size_t pos = 0;
    bool good = verify_fri_core(msgs, pos, ch, F, generator, log_n, offset, num_queries, max_index, log_degree_bound, nullptr, &why);
    if (good && pos != msgs.size()) { good = false; why = "trailing messages after the last query"; }
    return finish(good, why, ok, reason);
}

// Verifier of the build-defined FibonacciSq STARK (stark101_prove).  Public input: the claimed
// a_{T-2} (`claimed_last`); the witness a1 is not needed.  On top of the FRI checks it authenticates f(x), f(gx),
// f(g^2 x) against the trace root and requires layer 0 of the FRI to equal alpha0 p0 + alpha1 p1 + alpha2 p2
// computed from them — the link between the trace commitment and the low-degree test.
extern "C" int stark101_verify(const uint8_t* proof_flat, size_t proof_len, uint64_t modulus, uint64_t generator, uint64_t claimed_last,
                               unsigned log_trace, unsigned log_blowup, size_t num_queries, int* ok, char* reason) {
    if (!proof_flat || !ok || modulus < 3) { api_set_error("stark101_verify: bad argument"); return ST_INVALID; }
    Msgs msgs;
    std::string why;
    if (!split_records(proof_flat, proof_len, msgs, &why)) return finish(false, why, ok, reason);
    const Fp F{modulus};
    const unsigned log_N = log_trace + log_blowup;
    if (log_N > 40 || ((F.p - 1) & ((((uint64_t)1) << log_N) - 1)) != 0) return finish(false, "domain does not divide p-1", ok, reason);
    const size_t T = (size_t)1 << log_trace, N = (size_t)1 << log_N, blow = (size_t)1 << log_blowup;
    const uint64_t g = h_pow(generator % F.p, (F.p - 1) >> log_trace, F.p), h = h_pow(generator % F.p, (F.p - 1) >> log_N, F.p);
    const uint64_t w = generator % F.p;
    Channel ch(F.p);
    size_t pos = 0;
    // the statement is part of the transcript: every challenge depends on it (stark101_statement, host_channel.hpp)
    uint8_t want_stmt[STARK101_STATEMENT_BYTES];
    stark101_statement(F.p, generator % F.p, log_trace, log_blowup, num_queries, claimed_last % F.p, want_stmt);
    if (msgs.empty() || msgs[0].size() != sizeof want_stmt || memcmp(msgs[0].data(), want_stmt, sizeof want_stmt) != 0)
        return finish(false, "first message is not this statement (modulus, generator, sizes, queries, claimed a_{T-2})", ok, reason);
    ch.send(msgs[0].data(), msgs[0].size());
    std::array<uint8_t, 32> f_root{};
    if (msgs.size() < 2 || !parse_hex_root(msgs[1], f_root.data())) return finish(false, "second message is not the trace root", ok, reason);
    ch.send(msgs[1].data(), msgs[1].size());
    pos = 2;
    uint64_t alpha[3];
    for (int k = 0; k < 3; k++) {
        uint64_t rec, want;
        if (pos >= msgs.size() || !parse_be8(msgs[pos], &rec) || !ch.receive_random_field_element(&want) || want != rec % F.p)
            return finish(false, "alpha does not follow from the transcript", ok, reason);
        alpha[k] = want; pos++;
    }
    const uint64_t x_last = h_pow(g, T - 2, F.p), ex0 = h_pow(g, T - 3, F.p), ex2 = h_pow(g, T - 1, F.p);
    QueryHook hook = [&](size_t idx, const Msgs& ms, size_t& p, Channel& c, uint64_t* expect0, bool* has, std::string* wy) {
        uint64_t fv[3];
        for (int t = 0; t < 3; t++) {                                       // f(x), f(g x), f(g^2 x): g = h^blow
            if (p + 2 > ms.size() || !parse_be8(ms[p], &fv[t]) || fv[t] >= F.p) { *wy = "truncated trace opening"; return false; }
            if (!merkle_verify(f_root.data(), N, idx + t * blow, fv[t], ms[p + 1].data(), ms[p + 1].size())) { *wy = "trace opening does not authenticate against the trace root"; return false; }
            c.send(ms[p].data(), ms[p].size()); c.send(ms[p + 1].data(), ms[p + 1].size());
            p += 2;
        }
        const uint64_t x = F.mul(w, h_pow(h, idx, F.p));
        const uint64_t p0 = F.mul(F.sub(fv[0], 1 % F.p), F.inv(F.sub(x, 1 % F.p)));
        const uint64_t p1 = F.mul(F.sub(fv[0], claimed_last % F.p), F.inv(F.sub(x, x_last)));
        const uint64_t num = F.sub(F.sub(fv[2], F.mul(fv[1], fv[1])), F.mul(fv[0], fv[0]));
        const uint64_t e3 = F.mul(F.mul(F.sub(x, ex0), F.sub(x, x_last)), F.sub(x, ex2));
        const uint64_t p2 = F.mul(F.mul(num, e3), F.inv(F.sub(h_pow(x, T, F.p), 1 % F.p)));
        *expect0 = F.add(F.add(F.mul(alpha[0], p0), F.mul(alpha[1], p1)), F.mul(alpha[2], p2));
        *has = true;
        return true;
    };
    // deg CP <= T - 1 (p0, p1: deg f - 1 <= T - 3; p2: 2 (T - 2) + 3 - T): at most T coefficients, constant after log_trace folds
    bool good = verify_fri_core(msgs, pos, ch, F, generator, log_N, w, num_queries, N - 1 - 2 * blow, log_trace, hook, &why);
    if (good && pos != msgs.size()) { good = false; why = "trailing messages after the last query"; }
    return finish(good, why, ok, reason);
}
