// common.hpp — context, device buffers and error plumbing shared by the C-ABI and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "field.cuh"

namespace starkb200 {

enum : int { ST_OK = 0, ST_INVALID = 1, ST_CUDA = 2, ST_UNSUPPORTED = 3, ST_INTERNAL = 4 };

struct StarkError : std::runtime_error {
    int code;
    StarkError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define STARK_CUDA(expr)                                                                              \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess)                                                                        \
            throw ::starkb200::StarkError(::starkb200::ST_CUDA, std::string(#expr) + ": " +           \
                                                                    cudaGetErrorString(_e));          \
    } while (0)

#define STARK_REQUIRE(cond, msg)                                                                      \
    do {                                                                                              \
        if (!(cond)) throw ::starkb200::StarkError(::starkb200::ST_INVALID, msg);                     \
    } while (0)

// Stream-ordered device allocation.  Blocks of >= 4 KiB are recycled through an exact-size free list per
// stream before falling back to the cudaMallocAsync pool (small ones too since round 2: three cudaMallocAsync calls per FRI
// layer sat between a layer's root and the next launch, ~0.05 ms per proof; their number is capped so that a workload of
// ever-changing sizes cannot pile up entries): a prover allocates the same few sizes over and over
// (layer, tree levels, staging), and reusing the very same block on the very same stream is always ordered
// correctly and never depends on how the driver's pool splits and coalesces (measured: without the list a
// 64-column commit showed 10-200 ms allocation spikes once three sizes interleaved).
struct BlockCache {
    static std::mutex& mu() { static std::mutex m; return m; }
    struct PerStream { std::multimap<size_t, void*> blocks; size_t held = 0; int device = -1; };
    static std::map<cudaStream_t, PerStream>& lists() { static std::map<cudaStream_t, PerStream> l; return l; }
    static constexpr size_t kMinBytes = (size_t)1 << 12;          // smaller blocks: the driver pool is as fast as these lists
    static constexpr size_t kMaxHeld = (size_t)24 << 30;          // per stream
    static constexpr size_t kMaxBlocks = 4096;                    // per stream
    static void* take(cudaStream_t s, size_t bytes) {
        if (bytes < kMinBytes) return nullptr;
        std::lock_guard<std::mutex> g(mu());
        auto& l = lists()[s];
        auto it = l.blocks.find(bytes);
        if (it == l.blocks.end()) return nullptr;
        void* p = it->second;
        l.blocks.erase(it);
        l.held -= bytes;
        return p;
    }
    static bool give(cudaStream_t s, size_t bytes, void* p) {
        if (bytes < kMinBytes) return false;
        std::lock_guard<std::mutex> g(mu());
        auto& l = lists()[s];
        if (l.held + bytes > kMaxHeld || l.blocks.size() >= kMaxBlocks) return false;
        if (l.device < 0) cudaGetDevice(&l.device);
        l.blocks.emplace(bytes, p);
        l.held += bytes;
        return true;
    }
    static void flush(cudaStream_t s) {                           // context teardown
        std::lock_guard<std::mutex> g(mu());
        auto it = lists().find(s);
        if (it == lists().end()) return;
        for (auto& kv : it->second.blocks) cudaFreeAsync(kv.second, s);
        lists().erase(it);
    }
    // An allocation failed: hand every parked block of EVERY stream on the current device back to the driver's pool
    // (other contexts' lists may be holding the memory this one needs), wait for the frees, and let the caller retry.
    static void flush_device() {
        int dev = -1;
        cudaGetDevice(&dev);
        {
            std::lock_guard<std::mutex> g(mu());
            for (auto& kv : lists()) {
                if (kv.second.device != dev) continue;
                for (auto& b : kv.second.blocks) cudaFreeAsync(b.second, kv.first);
                kv.second.blocks.clear();
                kv.second.held = 0;
            }
        }
        cudaDeviceSynchronize();
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
    }
};

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    cudaStream_t stream = nullptr;
    bool plain = false;          // cudaMalloc'ed (exportable over CUDA IPC) instead of pool-allocated
    DevBuf() = default;
    DevBuf(size_t b, cudaStream_t s) : bytes(b), stream(s) {
        if (!b) return;
        p = BlockCache::take(s, b);
        if (!p) {
            cudaError_t e = cudaMallocAsync(&p, b, s);
            if (e == cudaErrorMemoryAllocation) {                 // free memory may be sitting in the recycling lists
                cudaGetLastError();
                BlockCache::flush_device();
                e = cudaMallocAsync(&p, b, s);
            }
            if (e != cudaSuccess) { p = nullptr; throw StarkError(ST_CUDA, std::string("cudaMallocAsync(") + std::to_string(b) + " bytes): " + cudaGetErrorString(e)); }
        }
    }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), bytes(o.bytes), stream(o.stream), plain(o.plain) { o.p = nullptr; o.bytes = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; bytes = o.bytes; stream = o.stream; plain = o.plain; o.p = nullptr; o.bytes = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void release() {
        if (p) {
            if (plain) { cudaStreamSynchronize(stream); cudaFree(p); }
            else if (!BlockCache::give(stream, bytes, p)) cudaFreeAsync(p, stream);
        }
        p = nullptr; bytes = 0;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};
using DevBufPtr = std::shared_ptr<DevBuf>;
inline DevBufPtr make_buf(size_t bytes, cudaStream_t s) { return std::make_shared<DevBuf>(bytes, s); }

// Two-level twiddle tables for one transform size: w^e = lo[e & mask] * hi[e >> shift].
struct TwiddleSet {
    DevBuf fwd_lo, fwd_hi, inv_lo, inv_hi;
    uint32_t shift = 0, mask = 0;
    PowTable fwd() const { return PowTable{fwd_lo.as<uint32_t>(), fwd_hi.as<uint32_t>(), shift, mask}; }
    PowTable inv() const { return PowTable{inv_lo.as<uint32_t>(), inv_hi.as<uint32_t>(), shift, mask}; }
};

// Growable pinned, device-mapped host buffer: kernels read descriptors from / write small results to it
// directly over PCIe, so a query opening costs one launch and one stream sync, no copy calls.
struct PinnedBuf {
    void* h = nullptr;
    void* d = nullptr;
    size_t cap = 0;
    void ensure(size_t n) {
        if (n <= cap) return;
        release();
        size_t c = 1 << 16;
        while (c < n) c <<= 1;
        STARK_CUDA(cudaHostAlloc(&h, c, cudaHostAllocMapped));
        STARK_CUDA(cudaHostGetDevicePointer(&d, h, 0));
        cap = c;
    }
    void release() {
        if (h) cudaFreeHost(h);
        h = d = nullptr; cap = 0;
    }
};

struct DegScratch { int maxv; unsigned ticket; };     // cross-block reduction state of the coefficient fold job / poly_degree_kernel

// What the host reads after each commit without issuing a copy (mapped pinned memory).
struct HostResult {
    uint32_t root[8];
    int32_t degree_plus1;
    uint32_t flag;
};

// Early hand-over of the top of a FRI layer's tree (merkle.cu: tail kernel, api.cu: wait_tree_top).  Between two layers of a
// commit the GPU idles while the host learns the root, feeds the transcript, draws beta and launches the next fold: ~7.5 us,
// 22 times per 2^24 proof.  The last CTA of the tail kernel therefore publishes the first level of <= 32 nodes it produces
// here (mapped pinned memory) and carries on with the remaining ~5 levels for the stored tree, one dependent parent hash
// (2.7 us) each; the host finishes the same levels itself with SHA-NI (31 hashes, ~3 us), draws beta and has the next launch
// queued before the kernel is done.
struct HostTop {
    uint32_t top_seq;          // written last (system-scope fence before it): which launch the nodes below belong to
    uint32_t top_len;          // 1 .. 32 nodes of one level, left to right
    uint32_t deg_seq;          // the coefficient fold job of that launch has published degree_plus1 (HostResult)
    uint32_t pad;
    uint32_t node[32][8];      // digests as the eight big-endian state words
};
constexpr int HOST_TOP_MAX = 32;

// ---- host-side modular helpers (u128; setup only, never on the data path) ----
inline uint64_t h_mul(uint64_t a, uint64_t b, uint64_t m) { return (uint64_t)((unsigned __int128)a * b % m); }
inline uint64_t h_pow(uint64_t a, uint64_t e, uint64_t m) {
    uint64_t r = 1 % m; a %= m;
    while (e) { if (e & 1) r = h_mul(r, a, m); a = h_mul(a, a, m); e >>= 1; }
    return r;
}
inline uint64_t h_inv(uint64_t a, uint64_t m) { return h_pow(a, m - 2, m); }

}  // namespace starkb200

struct stark_ctx;
void stark_ctx_teardown(stark_ctx* ctx);        // api.cu: what stark_ctx_destroy does once no handle is left
namespace starkb200 {
// The context pointer inside a handle: counts the handle in, and out again when the handle dies.
struct CtxRef {
    stark_ctx* p = nullptr;
    CtxRef() = default;
    CtxRef(const CtxRef&) = delete;
    CtxRef& operator=(const CtxRef&) = delete;
    CtxRef& operator=(stark_ctx* c);
    ~CtxRef() { *this = nullptr; }
    stark_ctx* operator->() const { return p; }
    operator stark_ctx*() const { return p; }
};
// RAII: brackets the launches issued in its scope with a CUDA-event pair when ctx->timing is on.
struct KernelTimer {
    stark_ctx* ctx; int cat; cudaEvent_t e0 = nullptr, e1 = nullptr;
    KernelTimer(stark_ctx* c, int category, double units = 0);
    ~KernelTimer();
};
}  // namespace starkb200

// The opaque C-ABI context (one per device x modulus).
struct stark_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;         // created on first use (ensure_copy_stream): FRI layers on their way to the host
    cudaEvent_t copy_event = nullptr;           //   while `stream` hashes the following layers (stark_fri_begin_to_host)
    uint64_t modulus = 0, generator = 0;
    unsigned two_adicity = 0;       // largest k with 2^k | p-1
    unsigned small_log = 0;         // size (log2) of the in-tile twiddle table
    starkb200::FieldParams fp{};
    std::recursive_mutex mu;
    starkb200::DevBuf small_fwd, small_inv;    // w_{2^small_log}^k, k < 2^(small_log-1), Montgomery form
    std::map<unsigned, std::unique_ptr<starkb200::TwiddleSet>> tw;
    starkb200::HostResult* h_result = nullptr;  // pinned + mapped
    starkb200::HostResult* d_result = nullptr;  // device alias of h_result
    starkb200::HostTop* h_top = nullptr;        // pinned + mapped: early hand-over of a tree's top (see HostTop)
    starkb200::HostTop* d_top = nullptr;
    uint32_t top_seq = 0;                       // launches that published there so far
    starkb200::PinnedBuf pin_desc, pin_out;     // opening descriptors in, opening records out
    starkb200::PinnedBuf pin_stage;             // host-produced columns on their way to HBM (the FibonacciSq trace)
    starkb200::DevBuf deg_scratch;              // DegScratch of the coefficient fold job (coeff_job.cuh)
    starkb200::DevBuf tail_counter;             // "last CTA" ticket of merkle_tail_kernel (zero between launches)
    int sm_count = 148;
    unsigned long long launches = 0;            // kernels launched through this context
    // Handles (vectors, trees, FRI proofs, groups) keep their context alive: stark_ctx_destroy with handles outstanding
    // only marks the context, and the last handle to go tears it down (a C or Rust caller has no _children list).
    std::atomic<long> handles{0};
    std::atomic<bool> destroy_requested{false};

    // optional per-category kernel timing (CUDA events on `stream`); see KernelTimer
    enum { CAT_MERKLE_LEAF = 0, CAT_MERKLE_NODE = 1, CAT_NTT = 2, CAT_OTHER = 3, CAT_COUNT = 4 };
    bool timing = false;
    std::vector<cudaEvent_t> ev_free;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_used[CAT_COUNT];
    double algo_units[CAT_COUNT] = {0, 0, 0, 0};   // algorithmic work issued per category while timing (int-ops or bytes)

    uint32_t to_mont(uint64_t a) const { return (uint32_t)starkb200::h_mul(a % modulus, (uint64_t)1 << 32, modulus); }
    uint64_t root_of_unity(unsigned log_n) const { return starkb200::h_pow(generator, (modulus - 1) >> log_n, modulus); }
    const starkb200::TwiddleSet& twiddles(unsigned log_n);
};

inline starkb200::CtxRef& starkb200::CtxRef::operator=(stark_ctx* c) {
    if (p == c) return *this;
    stark_ctx* old = p;
    p = c;
    if (p) p->handles.fetch_add(1);
    if (old && old->handles.fetch_sub(1) == 1 && old->destroy_requested.load()) stark_ctx_teardown(old);
    return *this;
}
