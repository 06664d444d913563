// host_channel.hpp — the Fiat-Shamir transcript, host side (it is sequential string hashing; it stays on
// the CPU exactly as in the reference, src/channel/channel.rs:14-95).  Product code: this is the
// library's own Channel, not the test oracle.
#pragma once
#include <stdint.h>
#include <string.h>

#include <mutex>
#include <string>
#include <vector>
#if defined(__x86_64__)
#include <cpuid.h>
#include <immintrin.h>
#endif

namespace starkb200 {

// FIPS 180-4 SHA-256 (what sha256 1.5.0 `digest` computes for the transcript strings).
class HostSha256 {
  public:
    static void digest(const uint8_t* msg, size_t len, uint8_t out[32]) {
        uint32_t st[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
        size_t off = 0;
        const bool ni = have_sha_ni();
        if (ni) { size_t nb = len / 64; blocks_ni(st, msg, nb); off = nb * 64; }
        else for (; off + 64 <= len; off += 64) block(st, msg + off);
        uint8_t tail[128];
        size_t r = len - off;
        memset(tail, 0, sizeof tail);
        memcpy(tail, msg + off, r);
        tail[r] = 0x80;
        size_t tl = (r + 9 <= 64) ? 64 : 128;
        uint64_t bits = (uint64_t)len * 8;
        for (int i = 0; i < 8; i++) tail[tl - 1 - i] = (uint8_t)(bits >> (8 * i));
        ni ? block_ni(st, tail) : block(st, tail);
        if (tl == 128) ni ? block_ni(st, tail + 64) : block(st, tail + 64);
        for (int i = 0; i < 8; i++) {
            out[4 * i] = (uint8_t)(st[i] >> 24); out[4 * i + 1] = (uint8_t)(st[i] >> 16);
            out[4 * i + 2] = (uint8_t)(st[i] >> 8); out[4 * i + 3] = (uint8_t)st[i];
        }
    }
    // nblk whole blocks that already carry their padding (Channel::send builds state || hex(msg) || padding in one
    // buffer): one multi-block call, so the state never leaves the registers between the blocks of a message
    static void digest_padded(const uint8_t* blocks, size_t nblk, uint8_t out[32]) {
        uint32_t st[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
        if (have_sha_ni()) blocks_ni(st, blocks, nblk);
        else for (size_t i = 0; i < nblk; i++) block(st, blocks + 64 * i);
        for (int i = 0; i < 8; i++) {
            out[4 * i] = (uint8_t)(st[i] >> 24); out[4 * i + 1] = (uint8_t)(st[i] >> 16);
            out[4 * i + 2] = (uint8_t)(st[i] >> 8); out[4 * i + 3] = (uint8_t)st[i];
        }
    }
    // Streaming form used by Channel::send: H(state || hex(msg)) without materialising the concatenation.
    struct Stream {
        uint32_t st[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
        uint8_t buf[64];
        size_t fill = 0;
        uint64_t total = 0;
        bool ni = have_sha_ni();
        void blocks(const uint8_t* p, size_t nblk) {
            if (ni) blocks_ni(st, p, nblk);
            else for (size_t i = 0; i < nblk; i++) block(st, p + 64 * i);
        }
        void update(const uint8_t* p, size_t n) {
            total += n;
            if (fill) {
                size_t take = 64 - fill < n ? 64 - fill : n;
                memcpy(buf + fill, p, take); fill += take; p += take; n -= take;
                if (fill == 64) { blocks(buf, 1); fill = 0; }
            }
            if (n >= 64) { blocks(p, n / 64); p += (n / 64) * 64; n %= 64; }
            if (n) { memcpy(buf, p, n); fill = n; }
        }
        void finish(uint8_t out[32]) {
            uint8_t tail[128];
            memset(tail, 0, sizeof tail);
            memcpy(tail, buf, fill);
            tail[fill] = 0x80;
            size_t tl = (fill + 9 <= 64) ? 64 : 128;
            uint64_t bits = total * 8;
            for (int i = 0; i < 8; i++) tail[tl - 1 - i] = (uint8_t)(bits >> (8 * i));
            blocks(tail, tl / 64);
            for (int i = 0; i < 8; i++) {
                out[4 * i] = (uint8_t)(st[i] >> 24); out[4 * i + 1] = (uint8_t)(st[i] >> 16);
                out[4 * i + 2] = (uint8_t)(st[i] >> 8); out[4 * i + 3] = (uint8_t)st[i];
            }
        }
    };
    // lowercase hex of n bytes into out[2n] (no terminator); 16 bytes per step with SSSE3 when available
    static void hex_into(const uint8_t* b, size_t n, char* out) {
        size_t i = 0;
#if defined(__x86_64__)
        if (have_ssse3()) i = hex_ssse3(b, n, out);
#endif
        static const char* d = "0123456789abcdef";
        for (; i < n; i++) { out[2 * i] = d[b[i] >> 4]; out[2 * i + 1] = d[b[i] & 15]; }
    }
    static std::string hex(const uint8_t* b, size_t n) {       // lowercase, like const-hex / rs_merkle root_hex
        static const char* d = "0123456789abcdef";
        std::string s(2 * n, '0');
        for (size_t i = 0; i < n; i++) { s[2 * i] = d[b[i] >> 4]; s[2 * i + 1] = d[b[i] & 15]; }
        return s;
    }
    static std::string hex_digest(const std::string& s) {
        uint8_t dg[32];
        digest(reinterpret_cast<const uint8_t*>(s.data()), s.size(), dg);
        return hex(dg, 32);
    }

  private:
#if defined(__x86_64__)
    static bool have_ssse3() {
        static const bool ok = [] { unsigned a, b, c, d; return __get_cpuid(1, &a, &b, &c, &d) && ((c >> 9) & 1); }();
        return ok;
    }
    __attribute__((target("ssse3"))) static size_t hex_ssse3(const uint8_t* b, size_t n, char* out) {
        const __m128i lut = _mm_setr_epi8('0', '1', '2', '3', '4', '5', '6', '7', '8', '9', 'a', 'b', 'c', 'd', 'e', 'f');
        const __m128i m4 = _mm_set1_epi8(0x0f);
        size_t i = 0;
        for (; i + 16 <= n; i += 16) {
            __m128i v = _mm_loadu_si128((const __m128i*)(b + i));
            __m128i hi = _mm_shuffle_epi8(lut, _mm_and_si128(_mm_srli_epi16(v, 4), m4));
            __m128i lo = _mm_shuffle_epi8(lut, _mm_and_si128(v, m4));
            _mm_storeu_si128((__m128i*)(out + 2 * i), _mm_unpacklo_epi8(hi, lo));
            _mm_storeu_si128((__m128i*)(out + 2 * i + 16), _mm_unpackhi_epi8(hi, lo));
        }
        return i;
    }
#endif
    // The transcript hashes ~40 KB of hex text per query; like sha2 0.10.8 in the reference, use the x86 SHA
    // extensions when the CPU has them (checked once against the portable rounds).
    static bool have_sha_ni() {
#if defined(__x86_64__)
        static const bool ok = [] {
            unsigned a, b, c, d;
            if (!__get_cpuid_count(7, 0, &a, &b, &c, &d) || !((b >> 29) & 1)) return false;
            uint8_t blk[64];
            for (int i = 0; i < 64; i++) blk[i] = (uint8_t)(i * 29 + 5);
            uint32_t x[8] = {1, 2, 3, 4, 5, 6, 7, 8}, y[8] = {1, 2, 3, 4, 5, 6, 7, 8};
            block(x, blk); block_ni(y, blk);
            return memcmp(x, y, 32) == 0;
        }();
        return ok;
#else
        return false;
#endif
    }
#if defined(__x86_64__)
    // nblk consecutive blocks; the state stays in the (ABEF, CDGH) register layout of sha256rnds2 from the first block
    // to the last, so the layout shuffles and the store/reload of st[] are off the per-block dependency chain
    __attribute__((target("sha,sse4.1,ssse3"))) static void blocks_ni(uint32_t st[8], const uint8_t* p, size_t nblk) {
        alignas(16) static const uint32_t K[64] = {
            0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
            0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
            0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
            0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
            0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
            0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
            0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
            0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
        const __m128i MASK = _mm_set_epi64x(0x0c0d0e0f08090a0bULL, 0x0405060700010203ULL);
        __m128i TMP = _mm_loadu_si128((const __m128i*)&st[0]);
        __m128i S1 = _mm_loadu_si128((const __m128i*)&st[4]);
        TMP = _mm_shuffle_epi32(TMP, 0xB1);
        S1 = _mm_shuffle_epi32(S1, 0x1B);
        __m128i S0 = _mm_alignr_epi8(TMP, S1, 8);
        S1 = _mm_blend_epi16(S1, TMP, 0xF0);
        for (size_t blk = 0; blk < nblk; blk++, p += 64) {
            const __m128i S0_SAVE = S0, S1_SAVE = S1;
            __m128i M[4];
            for (int i = 0; i < 4; i++) M[i] = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i*)(p + 16 * i)), MASK);
#pragma GCC unroll 16
            for (int r = 0; r < 16; r++) {
                __m128i MSG = _mm_add_epi32(M[r & 3], _mm_load_si128((const __m128i*)&K[4 * r]));
                S1 = _mm_sha256rnds2_epu32(S1, S0, MSG);
                if (r >= 3 && r < 15) {
                    __m128i T = _mm_alignr_epi8(M[r & 3], M[(r + 3) & 3], 4);
                    M[(r + 1) & 3] = _mm_sha256msg2_epu32(_mm_add_epi32(M[(r + 1) & 3], T), M[r & 3]);
                }
                MSG = _mm_shuffle_epi32(MSG, 0x0E);
                S0 = _mm_sha256rnds2_epu32(S0, S1, MSG);
                if (r >= 1 && r < 13) M[(r + 3) & 3] = _mm_sha256msg1_epu32(M[(r + 3) & 3], M[r & 3]);
            }
            S0 = _mm_add_epi32(S0, S0_SAVE);
            S1 = _mm_add_epi32(S1, S1_SAVE);
        }
        TMP = _mm_shuffle_epi32(S0, 0x1B);
        S1 = _mm_shuffle_epi32(S1, 0xB1);
        S0 = _mm_blend_epi16(TMP, S1, 0xF0);
        S1 = _mm_alignr_epi8(S1, TMP, 8);
        _mm_storeu_si128((__m128i*)&st[0], S0);
        _mm_storeu_si128((__m128i*)&st[4], S1);
    }
    static void block_ni(uint32_t st[8], const uint8_t* p) { blocks_ni(st, p, 1); }
#else
    static void block_ni(uint32_t st[8], const uint8_t* p) { block(st, p); }
    static void blocks_ni(uint32_t st[8], const uint8_t* p, size_t nblk) { for (size_t i = 0; i < nblk; i++) block(st, p + 64 * i); }
#endif
    static uint32_t ror(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
    static void block(uint32_t st[8], const uint8_t* p) {
        static const uint32_t K[64] = {
            0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
            0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
            0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
            0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
            0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
            0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
            0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
            0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
        uint32_t w[64];
        for (int i = 0; i < 16; i++)
            w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
        for (int i = 16; i < 64; i++) {
            uint32_t s0 = ror(w[i - 15], 7) ^ ror(w[i - 15], 18) ^ (w[i - 15] >> 3);
            uint32_t s1 = ror(w[i - 2], 17) ^ ror(w[i - 2], 19) ^ (w[i - 2] >> 10);
            w[i] = w[i - 16] + s0 + w[i - 7] + s1;
        }
        uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
        for (int i = 0; i < 64; i++) {
            uint32_t t1 = h + (ror(e, 6) ^ ror(e, 11) ^ ror(e, 25)) + ((e & f) ^ (~e & g)) + K[i] + w[i];
            uint32_t t2 = (ror(a, 2) ^ ror(a, 13) ^ ror(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
            h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
    }
};

inline void be8(uint64_t v, uint8_t out[8]) {            // FieldElement::to_bytes, element.rs:59-61
    for (int i = 0; i < 8; i++) out[i] = (uint8_t)(v >> (56 - 8 * i));
}

// The statement of the build-defined FibonacciSq STARK as the FIRST transcript message: every
// Fiat-Shamir challenge then depends on the modulus, the generator, the sizes, the query count and the claimed a_{T-2}.
constexpr size_t STARK101_STATEMENT_BYTES = 48;
inline void stark101_statement(uint64_t modulus, uint64_t generator, unsigned log_trace, unsigned log_blowup, size_t num_queries,
                               uint64_t claimed_last, uint8_t out[STARK101_STATEMENT_BYTES]) {
    be8(modulus, out); be8(generator, out + 8); be8(log_trace, out + 16); be8(log_blowup, out + 24);
    be8((uint64_t)num_queries, out + 32); be8(claimed_last, out + 40);
}

// Vec<Vec<u8>> as one byte arena plus spans: a query appends ~90 messages, and a heap block per message costs more
// than hashing the short ones.
// The arena of a finished transcript goes back to a small pool and the next Channel starts on it: a proof logs ~0.6 MB,
// and first-touch page faults on a fresh block every proof cost more than copying the bytes.
struct MessageLog {
    std::vector<uint8_t> bytes;
    std::vector<std::pair<size_t, size_t>> spans;         // (offset, length)
    MessageLog() = default;
    MessageLog(const MessageLog&) = default;
    MessageLog& operator=(const MessageLog&) = default;
    MessageLog(MessageLog&&) = default;
    MessageLog& operator=(MessageLog&&) = default;
    ~MessageLog() { recycle(&bytes, false); }
    // take == true: hands out a pooled arena (empty, capacity kept) if there is one; false: gives one back
    static void recycle(std::vector<uint8_t>* v, bool take) {
        // leaked on purpose: a Channel with static storage duration may be destroyed after any function-local static
        static std::mutex& mu = *new std::mutex;
        static std::vector<std::vector<uint8_t>>& pool = *new std::vector<std::vector<uint8_t>>;
        std::lock_guard<std::mutex> g(mu);
        if (take) {
            if (!pool.empty()) { *v = std::move(pool.back()); pool.pop_back(); v->clear(); }
        } else if (v->capacity() >= ((size_t)1 << 16) && v->capacity() <= ((size_t)8 << 20) && pool.size() < 8) {
            pool.push_back(std::move(*v));
        }
    }
    void push(const uint8_t* p, size_t n) {
        // the ABI hands out pointers into this arena (stark_channel_proof_msg): a caller re-sending one of those messages
        // passes a pointer the reserve below may free -- remember where it points and re-derive it afterwards
        const bool inside = n && !bytes.empty() && p >= bytes.data() && p < bytes.data() + bytes.size();
        const size_t self_off = inside ? (size_t)(p - bytes.data()) : 0;
        if (bytes.capacity() == 0) recycle(&bytes, true);
        if (spans.size() == spans.capacity()) spans.reserve(spans.empty() ? 4096 : 2 * spans.size());
        if (bytes.size() + n > bytes.capacity()) {           // a query appends ~20 KB: grow in big steps, few re-copies
            size_t want = 2 * bytes.capacity();
            if (want < bytes.size() + n) want = bytes.size() + n;
            if (want < ((size_t)1 << 20)) want = (size_t)1 << 20;
            bytes.reserve(want);
        }
        spans.emplace_back(bytes.size(), n);
        if (inside) {                                        // capacity is already there: resize never reallocates here
            const size_t at = bytes.size();
            bytes.resize(at + n);
            memmove(bytes.data() + at, bytes.data() + self_off, n);
        } else {
            bytes.insert(bytes.end(), p, p + n);
        }
    }
    size_t size() const { return spans.size(); }
    const uint8_t* data(size_t i) const { return bytes.data() + spans[i].first; }
    size_t len(size_t i) const { return spans[i].second; }
    size_t total() const { return bytes.size(); }
};

// Channel<MODULUS> — channel.rs:14-95, field for field.
struct Channel {
    MessageLog proof;                                      // :16
    std::vector<size_t> compressed_idx;                    // :17 compressed_proof: the same bytes as proof[i], stored once
    std::string state;                                     // :19, "" initially (:24-30)
    uint64_t modulus;
    std::vector<uint8_t> scratch;                          // state || hex(message) || padding of the send in flight

    explicit Channel(uint64_t m) : modulus(m) {}

    void send(const uint8_t* msg, size_t len) {            // :35-44
        // state = sha256::digest(old_state + hex::encode(message)): the text and its padding are laid out in one scratch
        // buffer and hashed by ONE multi-block call, so the hash state stays in registers from the first block of a
        // message to its last (a query sends ~90 messages of 2 .. 26 blocks; per-message set-up was a third of the time)
        const size_t sl = state.size();                    // 0 before the first send, 64 afterwards
        const size_t body = sl + 2 * len;
        const size_t total = ((body + 9 + 63) / 64) * 64;
        if (scratch.size() < total) scratch.resize(total < 4096 ? 4096 : total);
        uint8_t* b = scratch.data();
        memcpy(b, state.data(), sl);
        HostSha256::hex_into(msg, len, reinterpret_cast<char*>(b + sl));
        memset(b + body, 0, total - body);
        b[body] = 0x80;
        const uint64_t bits = (uint64_t)body * 8;
        for (int i = 0; i < 8; i++) b[total - 1 - i] = (uint8_t)(bits >> (8 * i));
        uint8_t dg[32];
        HostSha256::digest_padded(b, total / 64, dg);
        state.resize(64);
        HostSha256::hex_into(dg, 32, &state[0]);
        proof.push(msg, len);
        compressed_idx.push_back(proof.size() - 1);      // compressed_proof.push(message.to_vec()), kept as an index
    }
    // :58-84.  Returns false where the reference would panic ("Channel state is not valid hex" on "").
    bool receive_random_int(uint64_t min, uint64_t max, bool show_in_proof, uint64_t* out) {
        if (state.empty() || max < min) return false;
        const uint64_t range = (max - min) + 1;           // :68
        if (range == 0) return false;                     // usize overflow in the reference
        // (U256::from_str_radix(state,16) + min) % range  (:72), digit by digit
        unsigned __int128 acc = 0;
        for (char ch : state) {
            unsigned d = (ch >= '0' && ch <= '9') ? (unsigned)(ch - '0') : (unsigned)(ch - 'a' + 10);
            acc = (acc * 16 + d) % range;
        }
        uint64_t num = (uint64_t)((acc + (unsigned __int128)(min % range)) % range);
        state = HostSha256::hex_digest(state);            // :75-76
        if (show_in_proof) { uint8_t b[8]; be8(num, b); proof.push(b, 8); }   // :78-80
        *out = num;                                        // :83
        return true;
    }
    bool receive_random_field_element(uint64_t* out) {     // :47-55
        uint64_t num;
        if (!receive_random_int(0, modulus - 1, false, &num)) return false;
        uint8_t b[8]; be8(num, b);
        proof.push(b, 8);
        *out = num % modulus;
        return true;
    }
    size_t proof_size() const { return proof.total(); }                                                               // :88-90
    size_t compressed_proof_size() const { size_t s = 0; for (size_t i : compressed_idx) s += proof.len(i); return s; } // :93-95
};

}  // namespace starkb200
