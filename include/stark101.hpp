// stark101.hpp — header-only C++ mirror of the reference crate's public types on top of the C ABI
// (include/stark_b200.h).  The reference is Rust; this image has no Rust toolchain, so the host side above
// the ABI is written in C++ with the reference's names, argument meaning and error behaviour:
//
//   FieldElement<M>   src/fields/element.rs:7-147      (host-side scalar arithmetic, `%`-based like the reference)
//   Polynomial<M>     src/polynomial/ops.rs:9-241      (new/trim, degree, evaluate; evaluate over a CosetFri domain
//                                                       and interpolate on a coset go to the device)
//   MerkleTree<M>     src/merkle/mod.rs:5-27           (new, root; + get_authentication_path, fri_commit.rs:157)
//   Channel<M>        src/channel/channel.rs:14-95
//   CosetFri<M>       src/fri/coset_fri.rs:9-51
//   fri_commit / decommit_fri_layers / decommit_fri    src/fri/fri_commit.rs:72-179
//
// Errors: the reference panics (ops.rs:143, merkle/mod.rs:25, channel.rs:65); here every failing ABI call
// throws stark101::Panic carrying stark_last_error().
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "stark_b200.h"

namespace stark101 {

struct Panic : std::runtime_error {
    using std::runtime_error::runtime_error;
};
inline void check(int rc) {
    if (rc != STARK_OK) throw Panic(stark_last_error());
}

// One context per (MODULUS, device); generator 0 = smallest generator of F_p^* (5 for p = 3221225473).
template <uint64_t MODULUS>
inline stark_ctx* context(int device = 0) {
    static stark_ctx* ctx = [device] {
        stark_ctx* c = nullptr;
        check(stark_ctx_create(MODULUS, 0, device, &c));
        return c;
    }();
    return ctx;
}

// ---------------------------------------------------------------- element.rs
template <uint64_t MODULUS>
class FieldElement {
    uint64_t value_;

  public:
    FieldElement() : value_(0) {}
    explicit FieldElement(uint64_t v) : value_(v % MODULUS) {}                       // new, :13-17
    static FieldElement zero() { return FieldElement(0); }
    static FieldElement one() { return FieldElement(1); }
    uint64_t value() const { return value_; }
    FieldElement pow(uint64_t e) const {                                             // :38-51 (u64 products, as written)
        uint64_t r = 1, b = value_;
        while (e) { if (e & 1) r = (r * b) % MODULUS; b = (b * b) % MODULUS; e >>= 1; }
        FieldElement out; out.value_ = r; return out;
    }
    FieldElement inverse() const { return pow(MODULUS - 2); }                         // :54-57, inverse(0) == 0
    std::vector<uint8_t> to_bytes() const {                                          // :59-61 big-endian
        std::vector<uint8_t> b(8);
        for (int i = 0; i < 8; i++) b[i] = (uint8_t)(value_ >> (56 - 8 * i));
        return b;
    }
    bool operator==(const FieldElement& o) const { return value_ == o.value_; }
    bool operator!=(const FieldElement& o) const { return value_ != o.value_; }
    FieldElement operator+(FieldElement o) const { return FieldElement(value_ + o.value_); }              // :72-78
    FieldElement operator-(FieldElement o) const { return FieldElement((MODULUS + value_ - o.value_) % MODULUS); }   // :86-92
    FieldElement operator*(FieldElement o) const { return FieldElement((uint64_t)((unsigned __int128)value_ * o.value_ % MODULUS)); }  // :102-108
    FieldElement operator/(FieldElement o) const { return *this * o.inverse(); }                          // :116-122
    FieldElement operator-() const { return FieldElement(MODULUS - value_); }                             // :130-136
};
static_assert(sizeof(FieldElement<7>) == sizeof(uint64_t), "FieldElement must be layout-compatible with u64 (repr(transparent))");

// ---------------------------------------------------------------- coset_fri.rs
template <uint64_t M>
struct CosetFri {
    FieldElement<M> offset, omega;
    size_t domain_size;
    CosetFri(FieldElement<M> off, FieldElement<M> w, size_t n) : offset(off), omega(w), domain_size(n) {}   // new, :22-29
    // the library's convention: omega = generator^((p-1)/n)
    static CosetFri with_size(FieldElement<M> off, unsigned log_n) {
        return CosetFri(off, FieldElement<M>(stark_ctx_root_of_unity(context<M>(), log_n)), (size_t)1 << log_n);
    }
    unsigned log_size() const { unsigned l = 0; while (((size_t)1 << l) < domain_size) l++; return l; }
    std::vector<FieldElement<M>> generate_coset_domain() const {                                           // :32-36
        check_omega();
        std::vector<FieldElement<M>> d(domain_size);
        check(stark_coset_domain(context<M>(), log_size(), offset.value(), reinterpret_cast<uint64_t*>(d.data())));
        return d;
    }
    void check_omega() const {
        if ((domain_size & (domain_size - 1)) != 0 || omega.value() != stark_ctx_root_of_unity(context<M>(), log_size()))
            throw Panic("CosetFri: the device path needs a power-of-two domain generated by generator^((p-1)/n)");
    }
};

// ---------------------------------------------------------------- ops.rs
template <uint64_t M>
struct Polynomial {
    std::vector<FieldElement<M>> coefficients;   // low -> high, trailing zeros trimmed
    long degree;                                 // -1 for the zero polynomial
    explicit Polynomial(std::vector<FieldElement<M>> c) : coefficients(std::move(c)) {                     // new, :19-37
        while (!coefficients.empty() && coefficients.back() == FieldElement<M>::zero()) coefficients.pop_back();
        degree = (long)coefficients.size() - 1;
    }
    static Polynomial zero() { return Polynomial({}); }
    bool is_zero() const { return degree == -1; }
    FieldElement<M> evaluate(FieldElement<M> x) const {                                                    // :76-83 (one point: Horner on the host)
        FieldElement<M> r = FieldElement<M>::zero();
        for (size_t i = coefficients.size(); i-- > 0;) r = r * x + coefficients[i];
        return r;
    }
    // domain.iter().map(|x| self.evaluate(*x)) (fri_commit.rs:78) on the device
    std::vector<FieldElement<M>> evaluate_domain(const CosetFri<M>& d) const {
        d.check_omega();
        std::vector<FieldElement<M>> out(d.domain_size);
        check(stark_coset_evaluate(context<M>(), reinterpret_cast<const uint64_t*>(coefficients.data()), coefficients.size(),
                                   d.log_size(), d.offset.value(), reinterpret_cast<uint64_t*>(out.data())));
        return out;
    }
    // Polynomial::interpolate(xs, ys) (:239-241) for xs = d.generate_coset_domain()
    static Polynomial interpolate(const CosetFri<M>& d, const std::vector<FieldElement<M>>& ys) {
        d.check_omega();
        if (ys.size() != d.domain_size) throw Panic("Mismatched x and y lengths");                         // interpolation.rs:126-132
        std::vector<FieldElement<M>> c(d.domain_size);
        check(stark_coset_interpolate(context<M>(), reinterpret_cast<const uint64_t*>(ys.data()), d.log_size(), d.offset.value(),
                                      reinterpret_cast<uint64_t*>(c.data())));
        return Polynomial(std::move(c));
    }
};

// ---------------------------------------------------------------- merkle/mod.rs
template <uint64_t M>
class MerkleTree {
    std::shared_ptr<stark_tree> inner_;
    const stark_tree* borrowed_ = nullptr;

  public:
    static MerkleTree create(const std::vector<FieldElement<M>>& data) {                                   // new, :10-22
        stark_tree* t = nullptr;
        check(stark_merkle_commit(context<M>(), reinterpret_cast<const uint64_t*>(data.data()), data.size(), &t));
        MerkleTree m;
        m.inner_.reset(t, stark_tree_destroy);
        return m;
    }
    static MerkleTree borrow(const stark_tree* t) { MerkleTree m; m.borrowed_ = t; return m; }
    const stark_tree* raw() const { return inner_ ? inner_.get() : borrowed_; }
    std::string root() const {                                                                             // :24-26
        char buf[65];
        check(stark_merkle_root_hex(raw(), buf));
        return std::string(buf);
    }
    std::vector<uint8_t> get_authentication_path(size_t idx) const {                                       // fri_commit.rs:157
        size_t n = 0;
        check(stark_merkle_open(raw(), idx, nullptr, 0, &n));
        std::vector<uint8_t> p(n);
        if (n) check(stark_merkle_open(raw(), idx, p.data(), p.size(), &n));
        return p;
    }
};

// ---------------------------------------------------------------- channel.rs
template <uint64_t M>
class Channel {
    std::shared_ptr<stark_channel> ch_;

  public:
    Channel() {                                                                                            // new, :24-30
        stark_channel* c = nullptr;
        check(stark_channel_new(M, &c));
        ch_.reset(c, stark_channel_destroy);
    }
    stark_channel* raw() const { return ch_.get(); }
    void send(const std::vector<uint8_t>& m) { check(stark_channel_send(raw(), m.data(), m.size())); }     // :35-44
    void send(const std::string& s) { check(stark_channel_send(raw(), reinterpret_cast<const uint8_t*>(s.data()), s.size())); }
    FieldElement<M> receive_random_field_element() {                                                       // :47-55
        uint64_t v = 0;
        check(stark_channel_receive_random_field_element(raw(), &v));
        return FieldElement<M>(v);
    }
    size_t receive_random_int(size_t min, size_t max, bool show_in_proof) {                                // :58-84
        uint64_t v = 0;
        check(stark_channel_receive_random_int(raw(), min, max, show_in_proof ? 1 : 0, &v));
        return (size_t)v;
    }
    size_t proof_size() const { return stark_channel_proof_size(raw()); }                                  // :88-90
    size_t compressed_proof_size() const { return stark_channel_compressed_proof_size(raw()); }            // :93-95
    std::string state() const { return stark_channel_state(raw()); }
    std::vector<std::vector<uint8_t>> proof() const {
        std::vector<std::vector<uint8_t>> out;
        for (size_t i = 0, n = stark_channel_proof_len(raw()); i < n; i++) {
            const uint8_t* d = nullptr;
            size_t len = stark_channel_proof_msg(raw(), i, &d);
            out.emplace_back(d, d + len);
        }
        return out;
    }
};

// ---------------------------------------------------------------- fri_commit.rs
template <uint64_t M>
class FRIProof {                                                                                           // :9-13
    std::shared_ptr<stark_fri> f_;

  public:
    explicit FRIProof(stark_fri* f) : f_(f, stark_fri_destroy) {}
    stark_fri* raw() const { return f_.get(); }
    size_t num_layers() const { return stark_fri_num_layers(raw()); }
    std::vector<FieldElement<M>> fri_layer(size_t k) const {                                               // fri_layers[k]
        std::vector<FieldElement<M>> v(stark_fri_layer_len(raw(), k));
        check(stark_fri_layer_read(raw(), k, 0, v.size(), reinterpret_cast<uint64_t*>(v.data())));
        return v;
    }
    MerkleTree<M> fri_merkle(size_t k) const { return MerkleTree<M>::borrow(stark_fri_layer_tree(raw(), k)); }   // fri_merkles[k]
    Polynomial<M> final_poly() const {
        uint64_t v = 0; size_t n = 0;
        check(stark_fri_final(raw(), &v, &n));
        return Polynomial<M>(n ? std::vector<FieldElement<M>>{FieldElement<M>(v)} : std::vector<FieldElement<M>>{});
    }
};

// fri_commit(poly, domain, &mut channel) -> FRIProof, :72-122 — the loop is spelled out with the step API so
// that it reads like the reference (the one-call form is stark_fri_commit).
template <uint64_t M>
FRIProof<M> fri_commit(const Polynomial<M>& poly, const CosetFri<M>& domain, Channel<M>& channel) {
    domain.check_omega();
    stark_fri* f = nullptr;
    uint8_t root[32];
    auto hex = [](const uint8_t* r) {
        static const char* d = "0123456789abcdef";
        std::string s(64, '0');
        for (int i = 0; i < 32; i++) { s[2 * i] = d[r[i] >> 4]; s[2 * i + 1] = d[r[i] & 15]; }
        return s;
    };
    check(stark_fri_begin(context<M>(), reinterpret_cast<const uint64_t*>(poly.coefficients.data()), poly.coefficients.size(),
                          domain.log_size(), domain.offset.value(), &f, root));                           // :78-79
    FRIProof<M> proof(f);
    channel.send(hex(root));                                                                               // :86
    for (;;) {
        long long degree = 0;
        check(stark_fri_degree(f, &degree));
        if (degree < 1) break;                                                                             // while poly.degree >= 1, :89
        FieldElement<M> beta = channel.receive_random_field_element();                                     // :91
        check(stark_fri_fold(f, beta.value(), root));                                                      // :94-97
        channel.send(hex(root));                                                                           // :100
    }
    uint64_t fv = 0; size_t fl = 0;
    check(stark_fri_final(f, &fv, &fl));
    channel.send(FieldElement<M>(fv).to_bytes());                                                          // :109-114
    return proof;
}
template <uint64_t M>
void decommit_fri_layers(size_t index, const FRIProof<M>& proof, Channel<M>& channel) {                    // :137-165
    check(stark_decommit_fri_layers(proof.raw(), index, channel.raw()));
}
template <uint64_t M>
void decommit_fri(size_t num_queries, size_t max_index, const FRIProof<M>& proof, Channel<M>& channel) {   // :168-179
    for (size_t q = 0; q < num_queries; q++) {
        size_t idx = channel.receive_random_int(0, max_index, true);
        decommit_fri_layers(idx, proof, channel);
    }
}

}  // namespace stark101
