// C++ caller of the drop-in boundary through include/stark101.hpp, written like the reference's own unit
// tests (src/fields/element.rs:149-290, src/polynomial/ops.rs:551-1089 use GF(7)).
//   test_stark101 host                      -> host-side mirror only (no GPU needed)
//   test_stark101 gpu LOG_N LOG_DEG SEED Q  -> fri_commit + decommit_fri on the device; prints the transcript
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "stark101.hpp"

using namespace stark101;
constexpr uint64_t P = 3221225473ull;

#define EXPECT(c) do { if (!(c)) { fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

static uint64_t splitmix64_at(uint64_t seed, uint64_t k) {      // k-th output (k >= 1), SURVEY.md 8(d)
    uint64_t z = seed + k * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static int host_tests() {
    using F = FieldElement<7>;
    EXPECT(F(3).inverse() == F(5));                               // element.rs:186-190
    EXPECT(F(3).pow(3) == F(6));                                  // :193-197
    EXPECT(F(1) / F(3) == F(5));                                  // :178-183
    EXPECT(F(0).inverse() == F(0));                               // 0^(M-2): no panic
    EXPECT(-F(0) == F(0) && F(2) - F(5) == F(4) && F(10) == F(3));
    EXPECT(F(6).to_bytes() == (std::vector<uint8_t>{0, 0, 0, 0, 0, 0, 0, 6}));
    Polynomial<7> p({F(1), F(2), F(0), F(0)});                    // ops.rs:19-37: trailing zeros trimmed
    EXPECT(p.degree == 1 && p.coefficients.size() == 2);
    EXPECT(p.evaluate(F(3)) == F(0));                             // 1 + 2*3 = 7 = 0
    EXPECT(Polynomial<7>::zero().is_zero() && Polynomial<7>({F(0)}).degree == -1);
    printf("host ok\n");
    return 0;
}

static int gpu_tests(unsigned log_n, unsigned log_deg, uint64_t seed, size_t queries) {
    using F = FieldElement<P>;
    std::vector<F> c((size_t)1 << log_deg);
    for (size_t i = 0; i < c.size(); i++) c[i] = F(splitmix64_at(seed, i + 1));
    if (c.back() == F::zero()) c.back() = F::one();
    Polynomial<P> poly(c);
    CosetFri<P> domain = CosetFri<P>::with_size(F(5), log_n);
    // MerkleTree::new / root (merkle/mod.rs) and the anchors of SURVEY 8(c)
    std::vector<F> v8;
    for (uint64_t i = 0; i < 8; i++) v8.push_back(F(i));
    EXPECT(MerkleTree<P>::create(v8).root() == "8bad90db1d14c89a4efad7446090ec18ebae364b0faeab226c6f9b16ecec53b0");
    bool panicked = false;
    try { MerkleTree<P>::create({}); } catch (const Panic&) { panicked = true; }                   // root() unwrap on None
    EXPECT(panicked);
    // evaluate over the domain == Horner at sampled points; interpolate inverts it
    std::vector<F> ev = poly.evaluate_domain(domain);
    std::vector<F> dom = domain.generate_coset_domain();
    for (size_t i : {(size_t)0, (size_t)1, dom.size() / 2, dom.size() - 1}) EXPECT(ev[i] == poly.evaluate(dom[i]));
    EXPECT(dom[1] == domain.offset * domain.omega);
    Polynomial<P> back = Polynomial<P>::interpolate(domain, ev);
    EXPECT(back.degree == poly.degree && back.coefficients == poly.coefficients);
    // fri_commit / decommit_fri (fri_commit.rs:72-179)
    Channel<P> channel;
    FRIProof<P> proof = fri_commit(poly, domain, channel);
    EXPECT(proof.num_layers() == (size_t)log_deg + 1);
    EXPECT(proof.fri_layer(0) == ev);
    EXPECT(MerkleTree<P>::create(ev).root() == proof.fri_merkle(0).root());
    printf("state_after_commit %s\n", channel.state().c_str());
    decommit_fri(queries, ((size_t)1 << log_n) - 1, proof, channel);
    printf("final_state %s\nproof_size %zu\ncompressed_proof_size %zu\nnum_layers %zu\n", channel.state().c_str(),
           channel.proof_size(), channel.compressed_proof_size(), proof.num_layers());
    for (size_t k = 0; k < proof.num_layers(); k++) printf("root %zu %s\n", k, proof.fri_merkle(k).root().c_str());
    Polynomial<P> fin = proof.final_poly();
    printf("final_poly_len %zu value %" PRIu64 "\n", fin.coefficients.size(), fin.is_zero() ? 0 : fin.coefficients[0].value());
    return 0;
}

int main(int argc, char** argv) {
    try {
        if (argc >= 2 && !strcmp(argv[1], "host")) return host_tests();
        if (argc >= 6 && !strcmp(argv[1], "gpu"))
            return gpu_tests((unsigned)atoi(argv[2]), (unsigned)atoi(argv[3]), strtoull(argv[4], nullptr, 10), (size_t)atoi(argv[5]));
    } catch (const Panic& e) {
        fprintf(stderr, "panic: %s\n", e.what());
        return 2;
    }
    fprintf(stderr, "usage: test_stark101 host | gpu LOG_N LOG_DEG SEED QUERIES\n");
    return 64;
}
