// api.cu — the extern "C" boundary declared in include/stark_b200.h.
// Everything here is host orchestration: argument checks, HBM allocation, kernel sequencing on the
// context's stream, and the one sync per commitment that hands a root to the (host) Channel.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>

#include "../../include/stark_b200.h"
#include "handles.hpp"

using namespace starkb200;

namespace {
thread_local std::string g_last_error;
void set_error(const std::string& s) { g_last_error = s; }
}  // namespace

#define STARK_API_GUARD_NULL(cond)                      \
    if (!(cond)) { set_error("null argument"); return ST_INVALID; }
#define API_BEGIN try {
#define API_END                                                                        \
    }                                                                                  \
    catch (const StarkError& e) { set_error(e.what()); return e.code; }                \
    catch (const std::bad_alloc&) { set_error("host out of memory"); return ST_INTERNAL; } \
    catch (const std::exception& e) { set_error(e.what()); return ST_INTERNAL; }       \
    return ST_OK;

struct CtxGuard {
    std::lock_guard<std::recursive_mutex> lk;
    explicit CtxGuard(stark_ctx* c) : lk(c->mu) { STARK_CUDA(cudaSetDevice(c->device)); }
};

extern "C" const char* stark_last_error(void) { return g_last_error.c_str(); }
extern "C" const char* stark_version(void) { return "stark-b200 0.1 (sm_100a)"; }

// ======================================================================================= context
static bool is_prime_u64(uint64_t n) {
    if (n < 2) return false;
    for (uint64_t q : {2ull, 3ull, 5ull, 7ull, 11ull, 13ull, 17ull, 19ull, 23ull, 29ull, 31ull, 37ull}) {
        if (n % q == 0) return n == q;
    }
    uint64_t d = n - 1; int s = 0;
    while ((d & 1) == 0) { d >>= 1; s++; }
    for (uint64_t a : {2ull, 3ull, 5ull, 7ull, 11ull, 13ull, 17ull, 19ull, 23ull, 29ull, 31ull, 37ull}) {
        uint64_t x = h_pow(a, d, n);
        if (x == 1 || x == n - 1) continue;
        bool comp = true;
        for (int i = 1; i < s && comp; i++) { x = h_mul(x, x, n); if (x == n - 1) comp = false; }
        if (comp) return false;
    }
    return true;
}
static std::vector<uint64_t> prime_factors(uint64_t n) {
    std::vector<uint64_t> f;
    for (uint64_t q = 2; q * q <= n; q += (q == 2 ? 1 : 2))
        if (n % q == 0) { f.push_back(q); while (n % q == 0) n /= q; }
    if (n > 1) f.push_back(n);
    return f;
}
static bool is_generator(uint64_t g, uint64_t p, const std::vector<uint64_t>& fac) {
    if (g % p == 0) return false;
    for (uint64_t q : fac) if (h_pow(g, (p - 1) / q, p) == 1) return false;
    return true;
}

extern "C" int stark_ctx_create(uint64_t modulus, uint64_t generator, int device, stark_ctx** out) {
    API_BEGIN
    STARK_REQUIRE(out != nullptr, "ctx_create: out is null");
    *out = nullptr;
    if (modulus >= ((uint64_t)1 << 32) || modulus < 3 || (modulus & 1) == 0 || !is_prime_u64(modulus))
        throw StarkError(ST_UNSUPPORTED, "modulus must be an odd prime < 2^32 (the reference's pow() multiplies in u64, element.rs:45,47)");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        throw StarkError(ST_CUDA, std::string("no CUDA device: this library has no CPU fallback (") + cudaGetErrorString(ce) + ")");
    STARK_REQUIRE(device >= 0 && device < ndev, "ctx_create: bad device ordinal");
    auto fac = prime_factors(modulus - 1);
    if (generator == 0) { for (generator = 2; !is_generator(generator, modulus, fac); generator++) {} }
    STARK_REQUIRE(is_generator(generator, modulus, fac), "ctx_create: `generator` does not generate F_p^*");
    std::unique_ptr<stark_ctx> c(new stark_ctx());
    c->device = device; c->modulus = modulus; c->generator = generator % modulus;
    for (uint64_t m = modulus - 1; (m & 1) == 0; m >>= 1) c->two_adicity++;
    c->small_log = c->two_adicity < 12 ? c->two_adicity : 12;
    c->fp.p = (uint32_t)modulus;
    uint32_t inv = c->fp.p;
    for (int i = 0; i < 5; i++) inv *= 2u - c->fp.p * inv;
    c->fp.pinv = inv;
    c->fp.one = (uint32_t)(((uint64_t)1 << 32) % modulus);
    c->fp.r2 = (uint32_t)h_mul(c->fp.one, c->fp.one, modulus);
    STARK_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop{};
    STARK_CUDA(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    STARK_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    cudaMemPool_t pool;
    STARK_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t thr = UINT64_MAX;
    STARK_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    STARK_CUDA(cudaHostAlloc((void**)&c->h_result, sizeof(HostResult), cudaHostAllocMapped));
    memset(c->h_result, 0, sizeof(HostResult));
    STARK_CUDA(cudaHostGetDevicePointer((void**)&c->d_result, c->h_result, 0));
    STARK_CUDA(cudaHostAlloc((void**)&c->h_top, sizeof(HostTop), cudaHostAllocMapped));
    memset(c->h_top, 0, sizeof(HostTop));
    STARK_CUDA(cudaHostGetDevicePointer((void**)&c->d_top, c->h_top, 0));
    // in-tile twiddles: w_{2^small_log}^k, k < 2^(small_log-1)
    size_t ns = c->small_log ? ((size_t)1 << (c->small_log - 1)) : 1;
    std::vector<uint32_t> f(ns), b(ns);
    uint64_t w = c->root_of_unity(c->small_log), wi = h_inv(w, modulus), af = 1, ab = 1;
    for (size_t k = 0; k < ns; k++) { f[k] = c->to_mont(af); b[k] = c->to_mont(ab); af = h_mul(af, w, modulus); ab = h_mul(ab, wi, modulus); }
    c->small_fwd = DevBuf(ns * 4, c->stream); c->small_inv = DevBuf(ns * 4, c->stream);
    STARK_CUDA(cudaMemcpyAsync(c->small_fwd.p, f.data(), ns * 4, cudaMemcpyHostToDevice, c->stream));
    STARK_CUDA(cudaMemcpyAsync(c->small_inv.p, b.data(), ns * 4, cudaMemcpyHostToDevice, c->stream));
    STARK_CUDA(cudaStreamSynchronize(c->stream));
    *out = c.release();
    API_END
}
extern "C" void stark_ctx_destroy(stark_ctx* ctx) {
    if (!ctx) return;
    ctx->destroy_requested.store(true);
    if (ctx->handles.load() > 0) return;         // the last handle released tears the context down (CtxRef, common.hpp)
    stark_ctx_teardown(ctx);
}
void stark_ctx_teardown(stark_ctx* ctx) {
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->tw.clear();
    ctx->small_fwd.release(); ctx->small_inv.release(); ctx->tail_counter.release(); ctx->deg_scratch.release();
    BlockCache::flush(ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream) {
        cudaStreamSynchronize(ctx->copy_stream);
        BlockCache::flush(ctx->copy_stream);
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamDestroy(ctx->copy_stream);
        cudaEventDestroy(ctx->copy_event);
    }
    for (int c = 0; c < stark_ctx::CAT_COUNT; c++) for (auto& pr : ctx->ev_used[c]) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    for (auto e : ctx->ev_free) cudaEventDestroy(e);
    ctx->pin_desc.release(); ctx->pin_out.release(); ctx->pin_stage.release();
    if (ctx->h_result) cudaFreeHost(ctx->h_result);
    if (ctx->h_top) cudaFreeHost(ctx->h_top);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}
extern "C" int stark_ctx_sync(stark_ctx* ctx) {
    API_BEGIN
    STARK_REQUIRE(ctx, "null ctx");
    CtxGuard g(ctx);
    STARK_CUDA(cudaStreamSynchronize(ctx->stream));
    API_END
}
extern "C" uint64_t stark_ctx_modulus(const stark_ctx* ctx) { return ctx ? ctx->modulus : 0; }
extern "C" uint64_t stark_ctx_generator(const stark_ctx* ctx) { return ctx ? ctx->generator : 0; }
extern "C" uint64_t stark_ctx_root_of_unity(const stark_ctx* ctx, unsigned log_n) {
    return (ctx && log_n <= ctx->two_adicity) ? ctx->root_of_unity(log_n) : 0;
}
extern "C" unsigned stark_ctx_two_adicity(const stark_ctx* ctx) { return ctx ? ctx->two_adicity : 0; }
extern "C" unsigned long long stark_ctx_launch_count(const stark_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" void* stark_ctx_stream(const stark_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

extern "C" int stark_ctx_set_timing(stark_ctx* ctx, int on) {
    API_BEGIN
    STARK_REQUIRE(ctx, "null ctx");
    CtxGuard g(ctx);
    ctx->timing = on != 0;
    API_END
}
extern "C" int stark_ctx_read_timing(stark_ctx* ctx, double ms[4], double units[4], unsigned long long launches[4]) {
    API_BEGIN
    STARK_REQUIRE(ctx && ms && units && launches, "read_timing: null argument");
    CtxGuard g(ctx);
    STARK_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int c = 0; c < stark_ctx::CAT_COUNT; c++) {
        double tot = 0;
        for (auto& pr : ctx->ev_used[c]) {
            float t = 0;
            STARK_CUDA(cudaEventElapsedTime(&t, pr.first, pr.second));
            tot += t;
            ctx->ev_free.push_back(pr.first); ctx->ev_free.push_back(pr.second);
        }
        ms[c] = tot; units[c] = ctx->algo_units[c]; launches[c] = ctx->ev_used[c].size();
        ctx->ev_used[c].clear(); ctx->algo_units[c] = 0;
    }
    API_END
}

// ======================================================================================= vectors
static DevBufPtr upload_u64(stark_ctx* ctx, const uint64_t* host, size_t n, size_t alloc_n = 0) {
    if (alloc_n < n) alloc_n = n;
    DevBufPtr buf = make_buf(std::max<size_t>(alloc_n, 1) * 4, ctx->stream);
    if (alloc_n > n) STARK_CUDA(cudaMemsetAsync(buf->as<uint32_t>() + n, 0, (alloc_n - n) * 4, ctx->stream));
    if (n) {
        DevBuf stage(n * 8, ctx->stream);
        STARK_CUDA(cudaMemcpyAsync(stage.p, host, n * 8, cudaMemcpyHostToDevice, ctx->stream));
        narrow_u64(ctx, stage.as<uint64_t>(), buf->as<uint32_t>(), n);
    }
    return buf;
}
static void download_u64(stark_ctx* ctx, const uint32_t* dev, size_t n, uint64_t* host) {
    if (!n) return;
    DevBuf stage(n * 8, ctx->stream);
    widen_u32(ctx, dev, stage.as<uint64_t>(), n);
    STARK_CUDA(cudaMemcpyAsync(host, stage.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    STARK_CUDA(cudaStreamSynchronize(ctx->stream));
}
static stark_vec* new_vec(stark_ctx* ctx, DevBufPtr b, size_t n) {
    stark_vec* v = new stark_vec(); v->ctx = ctx; v->buf = std::move(b); v->n = n; return v;
}

extern "C" int stark_vec_upload(stark_ctx* ctx, const uint64_t* host, size_t n, stark_vec** out) {
    API_BEGIN
    STARK_REQUIRE(ctx && out && (host || n == 0), "vec_upload: null argument");
    CtxGuard g(ctx);
    *out = new_vec(ctx, upload_u64(ctx, host, n), n);
    API_END
}
extern "C" int stark_vec_alloc(stark_ctx* ctx, size_t n, stark_vec** out) {
    API_BEGIN
    STARK_REQUIRE(ctx && out, "vec_alloc: null argument");
    CtxGuard g(ctx);
    DevBufPtr b = make_buf(std::max<size_t>(n, 1) * 4, ctx->stream);
    fill_zero(ctx, b->as<uint32_t>(), n);
    *out = new_vec(ctx, b, n);
    API_END
}
extern "C" int stark_vec_download(const stark_vec* v, size_t offset, size_t n, uint64_t* host) {
    API_BEGIN
    STARK_REQUIRE(v && (host || n == 0), "vec_download: null argument");
    STARK_REQUIRE(offset <= v->n && n <= v->n - offset, "vec_download: range out of bounds");
    CtxGuard g(v->ctx);
    download_u64(v->ctx, v->buf->as<uint32_t>() + offset, n, host);
    API_END
}
extern "C" int stark_vec_from_device(stark_ctx* ctx, const void* device_u32, size_t n, stark_vec** out) {
    API_BEGIN
    STARK_REQUIRE(ctx && out && (device_u32 || n == 0), "vec_from_device: null argument");
    CtxGuard g(ctx);
    DevBufPtr b = make_buf(std::max<size_t>(n, 1) * 4, ctx->stream);
    if (n) STARK_CUDA(cudaMemcpyAsync(b->p, device_u32, n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    STARK_CUDA(cudaStreamSynchronize(ctx->stream));      // the source belongs to the caller (another stream)
    *out = new_vec(ctx, b, n);
    API_END
}
extern "C" size_t stark_vec_len(const stark_vec* v) { return v ? v->n : 0; }
extern "C" void* stark_vec_device_ptr(const stark_vec* v) { return v ? v->buf->p : nullptr; }
extern "C" void stark_vec_destroy(stark_vec* v) {
    if (!v) return;
    cudaSetDevice(v->ctx->device);
    delete v;
}

// ======================================================================================= polynomial
static unsigned ceil_log2(size_t n) { unsigned l = 0; while (((size_t)1 << l) < n) l++; return l; }
static void check_offset(stark_ctx* ctx, uint64_t offset) {
    STARK_REQUIRE(offset % ctx->modulus != 0, "coset offset must be non-zero");
}

// coefficients (natural, length len <= 2^log_m, buffer of 2^log_m zero-padded) -> evaluations on offset*<w_{2^log_n}>
static DevBufPtr evaluate_on_coset(stark_ctx* ctx, const uint32_t* coeffs_padded, unsigned log_m, unsigned log_n, uint64_t offset) {
    STARK_REQUIRE(log_m <= log_n, "evaluate: more coefficients than domain points");
    STARK_REQUIRE(log_n <= ctx->two_adicity, "evaluate: 2^log_n does not divide p-1");
    size_t m = (size_t)1 << log_m, n = (size_t)1 << log_n;
    if (log_n >= 3 && log_m <= log_n - 3 && lde8_supported(log_n - 3)) {
        // blow-up >= 8: eight interleaved size-(n/8) transforms that share all their twiddles (ntt.cu, lde8)
        DevBufPtr out = make_buf(n * 4, ctx->stream);
        lde8_forward(ctx, coeffs_padded, m, false, out->as<uint32_t>(), log_n - 3, offset, 1);
        return out;
    }
    bool unit = (offset % ctx->modulus) == 1;
    ScaleTable st;
    if (!unit) build_scale_table(ctx, offset, 1, log_m, st);
    if (ntt_natural_supported(log_n, coeffs_padded, nullptr, nullptr)) {     // pool blocks are always 16-byte aligned
        // natural -> natural without a permutation sweep; the zero padding above m is never read
        DevBuf work(n * 4, ctx->stream);
        DevBufPtr out = make_buf(n * 4, ctx->stream);
        ntt_natural(ctx, coeffs_padded, m, work.as<uint32_t>(), out->as<uint32_t>(), log_n, false, unit ? nullptr : &st.view, nullptr);
        return out;
    }
    DevBuf tmp(m * 4, ctx->stream);
    bitrev_permute(ctx, coeffs_padded, tmp.as<uint32_t>(), log_m, unit ? nullptr : &st.view, true);
    DevBufPtr out = make_buf(n * 4, ctx->stream);
    ntt_dit(ctx, tmp.as<uint32_t>(), out->as<uint32_t>(), log_n, log_n - log_m, nullptr, false);
    return out;
}
// evaluations on offset*<w_n> (natural) -> n coefficients (natural)
static DevBufPtr interpolate_on_coset(stark_ctx* ctx, const uint32_t* evals, unsigned log_n, uint64_t offset) {
    STARK_REQUIRE(log_n <= ctx->two_adicity, "interpolate: 2^log_n does not divide p-1");
    size_t n = (size_t)1 << log_n;
    DevBuf tmp(n * 4, ctx->stream);
    ScaleTable st;   // c_j = n^-1 * offset^-j * raw_j
    build_scale_table(ctx, h_inv(offset % ctx->modulus, ctx->modulus), h_inv(n % ctx->modulus, ctx->modulus), log_n, st);
    DevBufPtr out = make_buf(n * 4, ctx->stream);
    if (ntt_natural_supported(log_n, evals, tmp.p, out->p)) {      // no staging copy, no permutation sweep
        ntt_natural(ctx, evals, n, tmp.as<uint32_t>(), out->as<uint32_t>(), log_n, true, nullptr, &st.view);
        return out;
    }
    STARK_CUDA(cudaMemcpyAsync(tmp.p, evals, n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    ntt_dif(ctx, tmp.as<uint32_t>(), log_n, true);
    bitrev_permute(ctx, tmp.as<uint32_t>(), out->as<uint32_t>(), log_n, &st.view, false);
    return out;
}
// evaluations on offset_in*<w_n> -> evaluations on offset_out*<w_{n*2^b}>; no permutation anywhere
static DevBufPtr lde_on_coset(stark_ctx* ctx, const uint32_t* evals, unsigned log_n, uint64_t offset_in, unsigned log_blowup, uint64_t offset_out) {
    unsigned log_N = log_n + log_blowup;
    STARK_REQUIRE(log_N <= ctx->two_adicity, "lde: 2^(log_n+log_blowup) does not divide p-1");
    size_t n = (size_t)1 << log_n, N = (size_t)1 << log_N;
    uint64_t p = ctx->modulus;
    DevBuf tmp(n * 4, ctx->stream);
    if (log_blowup == 3 && lde8_supported(log_n) && ntt_natural_supported(log_n, evals, tmp.p, nullptr)) {
        {
            DevBuf coef(n * 4, ctx->stream);
            // unscaled coefficients in natural order (no staging copy), gathered by the blow-up-by-8 transform's first pass
            ntt_natural(ctx, evals, n, tmp.as<uint32_t>(), coef.as<uint32_t>(), log_n, true, nullptr, nullptr);
            DevBufPtr out8 = make_buf(N * 4, ctx->stream);
            lde8_forward(ctx, coef.as<uint32_t>(), n, false, out8->as<uint32_t>(), log_n,
                         h_mul(offset_out % p, h_inv(offset_in % p, p), p), h_inv(n % p, p));
            return out8;
        }
    }
    STARK_CUDA(cudaMemcpyAsync(tmp.p, evals, n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    ntt_dif(ctx, tmp.as<uint32_t>(), log_n, true);                       // bit-reversed, unscaled coefficients
    if (log_blowup == 3 && lde8_supported(log_n)) {
        DevBufPtr out8 = make_buf(N * 4, ctx->stream);
        lde8_forward(ctx, tmp.as<uint32_t>(), n, true, out8->as<uint32_t>(), log_n,
                     h_mul(offset_out % p, h_inv(offset_in % p, p), p), h_inv(n % p, p));
        return out8;
    }
    ScaleTable st;   // c_j * n^-1 * (offset_out/offset_in)^j, applied on the way into the DIT
    build_scale_table(ctx, h_mul(offset_out % p, h_inv(offset_in % p, p), p), h_inv(n % p, p), log_n, st);
    DevBufPtr out = make_buf(N * 4, ctx->stream);
    ntt_dit(ctx, tmp.as<uint32_t>(), out->as<uint32_t>(), log_N, log_blowup, &st.view, false);
    return out;
}

extern "C" int stark_ntt(stark_ctx* ctx, uint64_t* inout, unsigned log_n) {
    API_BEGIN
    STARK_REQUIRE(ctx && inout, "ntt: null argument");
    STARK_REQUIRE(log_n <= 30 && log_n <= ctx->two_adicity, "ntt: 2^log_n does not divide p-1 (or log_n > 30)");
    CtxGuard g(ctx);
    size_t n = (size_t)1 << log_n;
    DevBufPtr c = upload_u64(ctx, inout, n);
    DevBufPtr e = evaluate_on_coset(ctx, c->as<uint32_t>(), log_n, log_n, 1);
    download_u64(ctx, e->as<uint32_t>(), n, inout);
    API_END
}
extern "C" int stark_intt(stark_ctx* ctx, uint64_t* inout, unsigned log_n) {
    API_BEGIN
    STARK_REQUIRE(ctx && inout, "intt: null argument");
    STARK_REQUIRE(log_n <= 30 && log_n <= ctx->two_adicity, "intt: 2^log_n does not divide p-1 (or log_n > 30)");
    CtxGuard g(ctx);
    size_t n = (size_t)1 << log_n;
    DevBufPtr e = upload_u64(ctx, inout, n);
    DevBufPtr c = interpolate_on_coset(ctx, e->as<uint32_t>(), log_n, 1);
    download_u64(ctx, c->as<uint32_t>(), n, inout);
    API_END
}
extern "C" int stark_coset_evaluate(stark_ctx* ctx, const uint64_t* coeffs, size_t n_coeffs, unsigned log_n, uint64_t offset, uint64_t* out) {
    API_BEGIN
    STARK_REQUIRE(ctx && out && (coeffs || n_coeffs == 0), "coset_evaluate: null argument");
    CtxGuard g(ctx);
    check_offset(ctx, offset);
    STARK_REQUIRE(log_n <= 30 && log_n <= ctx->two_adicity, "coset_evaluate: 2^log_n does not divide p-1 (or log_n > 30)");
    STARK_REQUIRE(n_coeffs <= ((size_t)1 << log_n), "coset_evaluate: more coefficients than domain points");
    unsigned log_m = ceil_log2(std::max<size_t>(n_coeffs, 1));
    DevBufPtr c = upload_u64(ctx, coeffs, n_coeffs, (size_t)1 << log_m);
    DevBufPtr e = evaluate_on_coset(ctx, c->as<uint32_t>(), log_m, log_n, offset);
    download_u64(ctx, e->as<uint32_t>(), (size_t)1 << log_n, out);
    API_END
}
extern "C" int stark_coset_interpolate(stark_ctx* ctx, const uint64_t* evals, unsigned log_n, uint64_t offset, uint64_t* coeffs_out) {
    API_BEGIN
    STARK_REQUIRE(ctx && evals && coeffs_out, "coset_interpolate: null argument");
    CtxGuard g(ctx);
    check_offset(ctx, offset);
    STARK_REQUIRE(log_n <= 30 && log_n <= ctx->two_adicity, "coset_interpolate: 2^log_n does not divide p-1 (or log_n > 30)");
    size_t n = (size_t)1 << log_n;
    DevBufPtr e = upload_u64(ctx, evals, n);
    DevBufPtr c = interpolate_on_coset(ctx, e->as<uint32_t>(), log_n, offset);
    download_u64(ctx, c->as<uint32_t>(), n, coeffs_out);
    API_END
}
extern "C" int stark_coset_lde(stark_ctx* ctx, const uint64_t* evals, unsigned log_n, uint64_t offset_in, unsigned log_blowup,
                               uint64_t offset_out, uint64_t* out) {
    API_BEGIN
    STARK_REQUIRE(ctx && evals && out, "coset_lde: null argument");
    CtxGuard g(ctx);
    check_offset(ctx, offset_in); check_offset(ctx, offset_out);
    STARK_REQUIRE(log_n <= 30 && log_blowup <= 30 && log_n + log_blowup <= 30 && log_n + log_blowup <= ctx->two_adicity,
                  "coset_lde: 2^(log_n+log_blowup) does not divide p-1 (or exceeds 2^30)");
    size_t n = (size_t)1 << log_n;
    DevBufPtr e = upload_u64(ctx, evals, n);
    DevBufPtr r = lde_on_coset(ctx, e->as<uint32_t>(), log_n, offset_in, log_blowup, offset_out);
    download_u64(ctx, r->as<uint32_t>(), n << log_blowup, out);
    API_END
}
extern "C" int stark_batch_inverse(stark_ctx* ctx, uint64_t* inout, size_t n) {
    API_BEGIN
    STARK_REQUIRE(ctx && (inout || n == 0), "batch_inverse: null argument");
    CtxGuard g(ctx);
    DevBufPtr a = upload_u64(ctx, inout, n);
    batch_inverse(ctx, a->as<uint32_t>(), nullptr, a->as<uint32_t>(), n);
    download_u64(ctx, a->as<uint32_t>(), n, inout);
    API_END
}
extern "C" int stark_quotient_pointwise(stark_ctx* ctx, const uint64_t* num, const uint64_t* den, size_t n, uint64_t* out) {
    API_BEGIN
    STARK_REQUIRE(ctx && ((num && den && out) || n == 0), "quotient_pointwise: null argument");
    CtxGuard g(ctx);
    DevBufPtr a = upload_u64(ctx, num, n), b = upload_u64(ctx, den, n);
    batch_inverse(ctx, b->as<uint32_t>(), a->as<uint32_t>(), b->as<uint32_t>(), n);
    download_u64(ctx, b->as<uint32_t>(), n, out);
    API_END
}
extern "C" int stark_coset_domain(stark_ctx* ctx, unsigned log_n, uint64_t offset, uint64_t* out) {
    API_BEGIN
    STARK_REQUIRE(ctx && out, "coset_domain: null argument");
    CtxGuard g(ctx);
    STARK_REQUIRE(log_n <= 30 && log_n <= ctx->two_adicity, "coset_domain: 2^log_n does not divide p-1");
    size_t n = (size_t)1 << log_n;
    DevBuf d(n * 4, ctx->stream);
    coset_domain(ctx, offset, log_n, d.as<uint32_t>());
    download_u64(ctx, d.as<uint32_t>(), n, out);
    API_END
}
static unsigned exact_log2(size_t n, const char* what) {
    STARK_REQUIRE(n >= 1 && (n & (n - 1)) == 0, std::string(what) + ": length must be a power of two");
    return ceil_log2(n);
}
extern "C" int stark_coset_evaluate_dev(stark_ctx* ctx, const stark_vec* coeffs, unsigned log_n, uint64_t offset, stark_vec** out) {
    API_BEGIN
    STARK_REQUIRE(ctx && coeffs && out && coeffs->ctx == ctx, "coset_evaluate_dev: bad argument");
    CtxGuard g(ctx);
    check_offset(ctx, offset);
    STARK_REQUIRE(log_n <= 30 && coeffs->n <= ((size_t)1 << log_n), "coset_evaluate_dev: more coefficients than domain points");
    unsigned log_m = ceil_log2(std::max<size_t>(coeffs->n, 1));
    size_t m = (size_t)1 << log_m;
    DevBufPtr c = coeffs->buf;
    if (m != coeffs->n) {
        c = make_buf(m * 4, ctx->stream);
        STARK_CUDA(cudaMemsetAsync(c->p, 0, m * 4, ctx->stream));
        STARK_CUDA(cudaMemcpyAsync(c->p, coeffs->buf->p, coeffs->n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    *out = new_vec(ctx, evaluate_on_coset(ctx, c->as<uint32_t>(), log_m, log_n, offset), (size_t)1 << log_n);
    API_END
}
extern "C" int stark_coset_interpolate_dev(stark_ctx* ctx, const stark_vec* evals, uint64_t offset, stark_vec** out) {
    API_BEGIN
    STARK_REQUIRE(ctx && evals && out && evals->ctx == ctx, "coset_interpolate_dev: bad argument");
    CtxGuard g(ctx);
    check_offset(ctx, offset);
    unsigned log_n = exact_log2(evals->n, "coset_interpolate_dev");
    *out = new_vec(ctx, interpolate_on_coset(ctx, evals->buf->as<uint32_t>(), log_n, offset), evals->n);
    API_END
}
extern "C" int stark_coset_lde_dev(stark_ctx* ctx, const stark_vec* evals, uint64_t offset_in, unsigned log_blowup, uint64_t offset_out, stark_vec** out) {
    API_BEGIN
    STARK_REQUIRE(ctx && evals && out && evals->ctx == ctx, "coset_lde_dev: bad argument");
    CtxGuard g(ctx);
    check_offset(ctx, offset_in); check_offset(ctx, offset_out);
    unsigned log_n = exact_log2(evals->n, "coset_lde_dev");
    STARK_REQUIRE(log_n + log_blowup <= 30, "coset_lde_dev: domain too large");
    *out = new_vec(ctx, lde_on_coset(ctx, evals->buf->as<uint32_t>(), log_n, offset_in, log_blowup, offset_out), evals->n << log_blowup);
    API_END
}
extern "C" int stark_batch_inverse_dev(stark_ctx* ctx, const stark_vec* a, stark_vec** out) {
    API_BEGIN
    STARK_REQUIRE(ctx && a && out && a->ctx == ctx, "batch_inverse_dev: bad argument");
    CtxGuard g(ctx);
    DevBufPtr r = make_buf(std::max<size_t>(a->n, 1) * 4, ctx->stream);
    batch_inverse(ctx, a->buf->as<uint32_t>(), nullptr, r->as<uint32_t>(), a->n);
    *out = new_vec(ctx, r, a->n);
    API_END
}
extern "C" int stark_quotient_pointwise_dev(stark_ctx* ctx, const stark_vec* num, const stark_vec* den, stark_vec** out) {
    API_BEGIN
    STARK_REQUIRE(ctx && num && den && out && num->ctx == ctx && den->ctx == ctx, "quotient_pointwise_dev: bad argument");
    STARK_REQUIRE(num->n == den->n, "quotient_pointwise_dev: length mismatch");
    CtxGuard g(ctx);
    DevBufPtr r = make_buf(std::max<size_t>(num->n, 1) * 4, ctx->stream);
    batch_inverse(ctx, den->buf->as<uint32_t>(), num->buf->as<uint32_t>(), r->as<uint32_t>(), num->n);
    *out = new_vec(ctx, r, num->n);
    API_END
}

// ---- building blocks of the multi-GPU four-step NTT; the exchange itself is the caller's NCCL call
extern "C" int stark_ntt_batch_dev(stark_ctx* ctx, stark_vec* v, unsigned log_m, int inverse) {
    API_BEGIN
    STARK_REQUIRE(ctx && v && v->ctx == ctx, "ntt_batch_dev: bad argument");
    CtxGuard g(ctx);
    size_t m = (size_t)1 << log_m;
    STARK_REQUIRE(log_m >= 1 && log_m <= 30 && v->n % m == 0 && v->n >= m, "ntt_batch_dev: length is not a multiple of 2^log_m");
    STARK_REQUIRE(log_m <= ctx->two_adicity, "ntt_batch_dev: 2^log_m does not divide p-1");
    size_t batch = v->n / m;
    DevBuf tmp(v->n * 4, ctx->stream);
    if (ntt_natural_supported(log_m, v->buf->p, tmp.p, v->buf->p) && v->n <= ((size_t)1 << 31)) {
        // strided passes into tmp, transposing last pass back into v: no permutation sweep, no copy
        ScaleTable st;
        if (inverse) build_scale_table(ctx, 1, h_inv(m % ctx->modulus, ctx->modulus), log_m, st);
        ntt_natural(ctx, v->buf->as<uint32_t>(), v->n, tmp.as<uint32_t>(), v->buf->as<uint32_t>(), log_m, inverse != 0, nullptr,
                    inverse ? &st.view : nullptr, batch);
        return ST_OK;
    }
    if (!inverse) {
        bitrev_permute(ctx, v->buf->as<uint32_t>(), tmp.as<uint32_t>(), log_m, nullptr, false, batch);
        ntt_dit(ctx, tmp.as<uint32_t>(), v->buf->as<uint32_t>(), log_m, 0, nullptr, false, batch);
    } else {
        ntt_dif(ctx, v->buf->as<uint32_t>(), log_m, true, batch);
        ScaleTable st;   // 1/m on the way out of the permutation
        build_scale_table(ctx, 1, h_inv(m % ctx->modulus, ctx->modulus), log_m, st);
        STARK_CUDA(cudaMemcpyAsync(tmp.p, v->buf->p, v->n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        bitrev_permute(ctx, tmp.as<uint32_t>(), v->buf->as<uint32_t>(), log_m, &st.view, false, batch);
    }
    API_END
}
extern "C" int stark_pow_mul_dev(stark_ctx* ctx, stark_vec* v, size_t inner_len, size_t outer0, int product, size_t inner_stride,
                                 uint64_t base, uint64_t c0, unsigned log_table) {
    API_BEGIN
    STARK_REQUIRE(ctx && v && v->ctx == ctx && inner_len >= 1, "pow_mul_dev: bad argument");
    STARK_REQUIRE(log_table <= 31, "pow_mul_dev: table too large");
    {   // every exponent the kernel forms must index inside the 2^log_table table
        const size_t n = v->n, outer_max = outer0 + (n ? (n - 1) / inner_len : 0), inner_max = std::min(inner_len, n ? n : 1) - 1;
        const unsigned __int128 e_max = product ? (unsigned __int128)inner_max * outer_max
                                                : (unsigned __int128)inner_max * inner_stride + outer_max;
        STARK_REQUIRE(e_max < ((unsigned __int128)1 << log_table), "pow_mul_dev: an exponent would fall outside the 2^log_table table");
    }
    CtxGuard g(ctx);
    ScaleTable st;
    build_scale_table(ctx, base % ctx->modulus, c0 % ctx->modulus, log_table, st);
    pow_mul(ctx, v->buf->as<uint32_t>(), v->n, inner_len, outer0, product != 0, inner_stride, st.view);
    API_END
}

// ======================================================================================= merkle
static void words_to_bytes(const uint32_t w[8], uint8_t out[32]) {
    for (int i = 0; i < 8; i++) { out[4 * i] = (uint8_t)(w[i] >> 24); out[4 * i + 1] = (uint8_t)(w[i] >> 16); out[4 * i + 2] = (uint8_t)(w[i] >> 8); out[4 * i + 3] = (uint8_t)w[i]; }
}
// Builds the tree over `src` on the stream (no sync); the root lands in ctx->h_result.
static std::unique_ptr<stark_tree> tree_launch(stark_ctx* ctx, DevBufPtr leaves, size_t n, const LeafSource& src, HostResult* result = nullptr,
                                               uint32_t top_seq = 0) {
    std::unique_ptr<stark_tree> t(new stark_tree());
    t->ctx = ctx; t->leaves = std::move(leaves);
    t->shape = TreeShape::make(n);
    t->nodes = DevBuf(t->shape.total * 32, ctx->stream);
    merkle_build(ctx, src, t->shape, t->nodes.as<uint32_t>(), result ? result : ctx->d_result, top_seq);
    return t;
}
static void tree_take_root(stark_tree* t) {   // after a stream sync
    memcpy(t->root_words, t->ctx->h_result->root, 32);
}
// Early hand-over of a tree's top (common.hpp: HostTop): spins on the sequence number the tail kernel publishes with its
// first level of <= 32 nodes (and on the coefficient job's, when the launch carried one), then finishes the tree on the host:
// pairwise SHA-256, a lone node promoted unchanged (the rs_merkle rule the kernels implement).  The kernel is still running
// when this returns -- whatever is launched next queues behind it.  Should the stream drain without the number showing up
// (it cannot, every exit of the kernel publishes), the root the kernel wrote is taken as before.
static void wait_tree_top(stark_ctx* ctx, uint32_t seq, bool has_job, uint32_t root_words[8]) {
    volatile HostTop* top = ctx->h_top;
    for (unsigned spins = 1;; spins++) {
        if (top->top_seq == seq && (!has_job || top->deg_seq == seq)) break;
        if ((spins & 4095) == 0 && cudaStreamQuery(ctx->stream) != cudaErrorNotReady) {
            STARK_CUDA(cudaStreamSynchronize(ctx->stream));
            if (top->top_seq == seq) break;
            memcpy(root_words, ctx->h_result->root, 32);
            return;
        }
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    size_t len = top->top_len;
    STARK_REQUIRE(len >= 1 && len <= (size_t)HOST_TOP_MAX, "merkle: malformed early hand-over of the tree top");
    uint8_t level[HOST_TOP_MAX][32], cat[64];
    for (size_t i = 0; i < len; i++) {
        uint32_t w[8];
        for (int k = 0; k < 8; k++) w[k] = top->node[i][k];
        words_to_bytes(w, level[i]);
    }
    while (len > 1) {
        const size_t out = (len + 1) / 2;
        for (size_t j = 0; j < out; j++) {
            if (2 * j + 1 < len) {
                memcpy(cat, level[2 * j], 32); memcpy(cat + 32, level[2 * j + 1], 32);
                HostSha256::digest(cat, 64, level[j]);
            } else if (j != 2 * j) memcpy(level[j], level[2 * j], 32);
        }
        len = out;
    }
    for (int i = 0; i < 8; i++)
        root_words[i] = ((uint32_t)level[0][4 * i] << 24) | ((uint32_t)level[0][4 * i + 1] << 16) | ((uint32_t)level[0][4 * i + 2] << 8) | level[0][4 * i + 3];
}
static std::unique_ptr<stark_tree> tree_commit(stark_ctx* ctx, DevBufPtr leaves, size_t n) {
    STARK_REQUIRE(n >= 1, "MerkleTree::new on an empty vector: root() would panic on unwrap (merkle/mod.rs:25)");
    STARK_REQUIRE(n <= ((size_t)1 << 32), "merkle: more than 2^32 leaves");
    LeafSource src; src.vals = leaves->as<uint32_t>();
    auto t = tree_launch(ctx, leaves, n, src);
    STARK_CUDA(cudaStreamSynchronize(ctx->stream));
    tree_take_root(t.get());
    return t;
}
extern "C" int stark_merkle_commit(stark_ctx* ctx, const uint64_t* leaves, size_t n, stark_tree** out) {
    API_BEGIN
    STARK_REQUIRE(ctx && out && (leaves || n == 0), "merkle_commit: null argument");
    CtxGuard g(ctx);
    STARK_REQUIRE(n >= 1, "MerkleTree::new on an empty vector: root() would panic on unwrap (merkle/mod.rs:25)");
    *out = tree_commit(ctx, upload_u64(ctx, leaves, n), n).release();
    API_END
}
extern "C" int stark_merkle_commit_dev(stark_ctx* ctx, const stark_vec* leaves, stark_tree** out) {
    API_BEGIN
    STARK_REQUIRE(ctx && out && leaves && leaves->ctx == ctx, "merkle_commit_dev: bad argument");
    CtxGuard g(ctx);
    *out = tree_commit(ctx, leaves->buf, leaves->n).release();
    API_END
}
extern "C" int stark_merkle_root(const stark_tree* t, uint8_t root[32]) {
    API_BEGIN
    STARK_REQUIRE(t && root, "merkle_root: null argument");
    words_to_bytes(t->root_words, root);
    API_END
}
extern "C" int stark_merkle_root_hex(const stark_tree* t, char out[65]) {
    API_BEGIN
    STARK_REQUIRE(t && out, "merkle_root_hex: null argument");
    uint8_t r[32]; words_to_bytes(t->root_words, r);
    std::string h = HostSha256::hex(r, 32);
    memcpy(out, h.c_str(), 65);
    API_END
}
extern "C" size_t stark_merkle_num_leaves(const stark_tree* t) { return t ? t->shape.n : 0; }
extern "C" size_t stark_merkle_depth(const stark_tree* t) { return t ? t->shape.depth : 0; }

// One launch for a batch of (tree, idx) records; records are BE8(value) || path.
static void open_records(stark_ctx* ctx, const std::vector<OpenDesc>& descs, size_t total_bytes, uint8_t* host_out) {
    if (descs.empty()) return;
    STARK_CUDA(cudaStreamSynchronize(ctx->stream));      // the pinned buffers may still be in use by an earlier launch
    ctx->pin_desc.ensure(descs.size() * sizeof(OpenDesc));
    ctx->pin_out.ensure(total_bytes);
    memcpy(ctx->pin_desc.h, descs.data(), descs.size() * sizeof(OpenDesc));
    merkle_open(ctx, static_cast<const OpenDesc*>(ctx->pin_desc.d), descs.size(), static_cast<uint8_t*>(ctx->pin_out.d));
    STARK_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(host_out, ctx->pin_out.h, total_bytes);
}
extern "C" int stark_merkle_open(const stark_tree* t, size_t idx, uint8_t* path, size_t cap, size_t* path_len) {
    API_BEGIN
    STARK_REQUIRE(t && path_len, "merkle_open: null argument");
    STARK_REQUIRE(idx < t->shape.n, "merkle_open: leaf index out of range");
    STARK_REQUIRE(!t->external, "merkle_open: this tree's levels are held by other ranks (leaf-range sharding); open on the owner");
    size_t pl = merkle_path_len(t->shape.n, idx);
    *path_len = pl;
    if (!path) return ST_OK;
    STARK_REQUIRE(cap >= pl, "merkle_open: buffer too small");
    CtxGuard g(t->ctx);
    std::vector<OpenDesc> d(1);
    d[0] = OpenDesc{t->leaves->as<uint32_t>(), t->nodes.as<uint32_t>(), t->shape.n, idx, 0};
    std::vector<uint8_t> rec(8 + pl);
    open_records(t->ctx, d, rec.size(), rec.data());
    memcpy(path, rec.data() + 8, pl);
    API_END
}
extern "C" int stark_merkle_node(const stark_tree* t, size_t level, size_t j, uint8_t out[32]) {
    API_BEGIN
    STARK_REQUIRE(t && out, "merkle_node: null argument");
    STARK_REQUIRE(level >= 1 && level <= t->shape.depth && j < t->shape.len[level], "merkle_node: no such node (level 0 digests are not stored)");
    STARK_REQUIRE(!t->external, "merkle_node: this tree's levels are held by other ranks");
    CtxGuard g(t->ctx);
    uint32_t w[8];
    STARK_CUDA(cudaMemcpyAsync(w, t->nodes.as<uint32_t>() + 8 * (t->shape.off[level] + j), 32, cudaMemcpyDeviceToHost, t->ctx->stream));
    STARK_CUDA(cudaStreamSynchronize(t->ctx->stream));
    words_to_bytes(w, out);
    API_END
}
extern "C" void stark_tree_destroy(stark_tree* t) {
    if (!t) return;
    cudaSetDevice(t->ctx->device);
    delete t;
}

// ======================================================================================= channel
extern "C" int stark_channel_new(uint64_t modulus, stark_channel** out) {
    API_BEGIN
    STARK_REQUIRE(out && modulus >= 2, "channel_new: bad argument");
    *out = new stark_channel(modulus);
    API_END
}
extern "C" void stark_channel_destroy(stark_channel* ch) { delete ch; }
extern "C" int stark_channel_send(stark_channel* ch, const uint8_t* msg, size_t len) {
    API_BEGIN
    STARK_REQUIRE(ch && (msg || len == 0), "channel_send: null argument");
    ch->ch.send(msg, len);
    API_END
}
extern "C" int stark_channel_receive_random_field_element(stark_channel* ch, uint64_t* out) {
    API_BEGIN
    STARK_REQUIRE(ch && out, "channel: null argument");
    STARK_REQUIRE(ch->ch.receive_random_field_element(out), "Channel state is not valid hex (receive before any send, channel.rs:64-65)");
    API_END
}
extern "C" int stark_channel_receive_random_int(stark_channel* ch, uint64_t min, uint64_t max, int show, uint64_t* out) {
    API_BEGIN
    STARK_REQUIRE(ch && out, "channel: null argument");
    STARK_REQUIRE(ch->ch.receive_random_int(min, max, show != 0, out), "Channel state is not valid hex / empty range (channel.rs:64-68)");
    API_END
}
extern "C" size_t stark_channel_proof_size(const stark_channel* ch) { return ch ? ch->ch.proof_size() : 0; }
extern "C" size_t stark_channel_compressed_proof_size(const stark_channel* ch) { return ch ? ch->ch.compressed_proof_size() : 0; }
extern "C" const char* stark_channel_state(const stark_channel* ch) { return ch ? ch->ch.state.c_str() : ""; }
extern "C" size_t stark_channel_proof_len(const stark_channel* ch) { return ch ? ch->ch.proof.size() : 0; }
extern "C" size_t stark_channel_proof_msg(const stark_channel* ch, size_t i, const uint8_t** data) {
    if (!ch || i >= ch->ch.proof.size()) return 0;
    if (data) *data = ch->ch.proof.data(i);              // valid until the next message is appended
    return ch->ch.proof.len(i);
}
extern "C" size_t stark_channel_compressed_len(const stark_channel* ch) { return ch ? ch->ch.compressed_idx.size() : 0; }
extern "C" size_t stark_channel_compressed_msg(const stark_channel* ch, size_t i, const uint8_t** data) {
    if (!ch || i >= ch->ch.compressed_idx.size()) return 0;
    const size_t k = ch->ch.compressed_idx[i];
    if (data) *data = ch->ch.proof.data(k);
    return ch->ch.proof.len(k);
}
extern "C" size_t stark_channel_proof_flat(const stark_channel* ch, uint8_t* out) {
    if (!ch) return 0;
    size_t w = 0;
    for (size_t i = 0; i < ch->ch.proof.size(); i++) {
        uint32_t n = (uint32_t)ch->ch.proof.len(i);
        if (out) { out[w] = n & 255; out[w + 1] = (n >> 8) & 255; out[w + 2] = (n >> 16) & 255; out[w + 3] = n >> 24; if (n) memcpy(out + w + 4, ch->ch.proof.data(i), n); }
        w += 4 + n;
    }
    return w;
}

// ======================================================================================= FRI
// ---- FRIProof.fri_layers by value: the layers stream to the host while the following ones are hashed ----
// The copy stream has the highest priority, so the widening kernel's CTAs are placed as soon as hashing CTAs retire; each
// layer is ~0.02 ms of HBM-bound widening and 8 B per element over PCIe, against 1384 integer instructions per leaf on the
// main stream -- the copies of a 2^24-domain proof (269 MB) end before its commit phase does.
static void ensure_copy_stream(stark_ctx* ctx) {
    if (ctx->copy_stream) return;
    int lo = 0, hi = 0;
    STARK_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    STARK_CUDA(cudaStreamCreateWithPriority(&ctx->copy_stream, cudaStreamNonBlocking, hi));
    STARK_CUDA(cudaEventCreateWithFlags(&ctx->copy_event, cudaEventDisableTiming));
}
static std::unique_ptr<stark_fri::LayerSink> make_sink(stark_ctx* ctx, unsigned log_n, uint64_t* host, size_t cap) {
    STARK_REQUIRE(host && cap >= ((size_t)1 << log_n), "fri: the host buffer must hold at least layer 0 (2^(log_n+1) elements hold every layer)");
    ensure_copy_stream(ctx);
    std::unique_ptr<stark_fri::LayerSink> s(new stark_fri::LayerSink());
    s->stream = ctx->copy_stream; s->host = host; s->cap = cap;
    s->stage = DevBuf(((size_t)8) << log_n, ctx->copy_stream);
    return s;
}
// `vals` is complete in main-stream order at the time of the call (after_main: the copy stream waits for that point;
// otherwise the caller has synchronised the main stream)
static void sink_push(stark_fri* f, const uint32_t* vals, size_t n, bool after_main) {
    stark_fri::LayerSink* s = f->sink.get();
    if (!s) return;
    stark_ctx* ctx = f->ctx;
    STARK_REQUIRE(n <= s->cap - s->off, "fri: the host buffer for the layers is full (2^(log_n+1) elements hold every layer)");
    if (after_main) {
        STARK_CUDA(cudaEventRecord(ctx->copy_event, ctx->stream));
        STARK_CUDA(cudaStreamWaitEvent(s->stream, ctx->copy_event, 0));
    }
    // experiment knob (tools/exp_by_value.py): 1 = no copy to the host, 2 = neither the widening kernel nor the copy
    static const int dbg = [] { const char* e = getenv("STARK_SINK_DEBUG"); return e ? atoi(e) : 0; }();
    if (dbg < 2) widen_u32_on(ctx, s->stream, vals, s->stage.as<uint64_t>(), n);
    if (dbg < 1) STARK_CUDA(cudaMemcpyAsync(s->host + s->off, s->stage.p, n * 8, cudaMemcpyDeviceToHost, s->stream));
    s->offs.push_back(s->off);
    s->off += n;
}
// every finished layer that has not been sent yet.  Called while the main stream is busy with the NEXT layer's launches
// (stark_fri_fold, between its launches and its synchronisation), so that the two enqueues per layer cost no time between a
// layer's root and the next launch; the last layer goes out with stark_fri_final / stark_fri_layers_wait.
static void sink_flush(stark_fri* f) {
    stark_fri::LayerSink* s = f->sink.get();
    if (!s) return;
    while (s->offs.size() < f->trees.size()) {
        const stark_tree* t = f->trees[s->offs.size()].get();
        sink_push(f, t->leaves->as<uint32_t>(), t->shape.n, false);
    }
}
// The layer the fold about to be launched will produce, computed a second time on the copy stream by the plain fold kernel
// (12 bytes of HBM traffic per point) and sent from there.  The fused fold-and-hash launch only has the layer complete when
// most of its tree is hashed, so a copy issued after it runs under the FOLLOWING, smaller layers -- and while the copy engine
// is bursting towards the host, every small write the latency-bound part of the commit sends the same way (a root into
// mapped memory, the completion of a stream synchronisation) queues behind it: +0.6 ms on a 2^24-domain commit
// (tools/exp_by_value.py).  Started at the beginning of its own tree, a layer's copy (8 B per point over PCIe) ends before that
// tree (one leaf hash + one node per point) does.
static void sink_push_fold(stark_fri* f, const LeafSource& src, bool after_main) {
    stark_fri::LayerSink* s = f->sink.get();
    if (!s) return;
    stark_ctx* ctx = f->ctx;
    sink_flush(f);
    STARK_REQUIRE(src.half <= s->cap - s->off, "fri: the host buffer for the layers is full (2^(log_n+1) elements hold every layer)");
    if (!s->fold_tmp.p) s->fold_tmp = DevBuf(((size_t)4 << f->log_n) / 2, s->stream);
    LeafSource s2 = src;
    s2.job = CoeffJob{};
    s2.fold_out = s->fold_tmp.as<uint32_t>();
    if (after_main) {                        // the twiddle table of this size has just been built on the main stream
        STARK_CUDA(cudaEventRecord(ctx->copy_event, ctx->stream));
        STARK_CUDA(cudaStreamWaitEvent(s->stream, ctx->copy_event, 0));
    }
    fri_fold_on(ctx, s->stream, s2);
    static const int dbg = [] { const char* e = getenv("STARK_SINK_DEBUG"); return e ? atoi(e) : 0; }();
    if (dbg < 2) widen_u32_on(ctx, s->stream, s2.fold_out, s->stage.as<uint64_t>(), src.half);
    if (dbg < 1) STARK_CUDA(cudaMemcpyAsync(s->host + s->off, s->stage.p, src.half * 8, cudaMemcpyDeviceToHost, s->stream));
    s->offs.push_back(s->off);
    s->off += src.half;
}

static void fri_begin_impl(stark_ctx* ctx, DevBufPtr coeffs_padded, size_t len, unsigned log_m, unsigned log_n,
                           uint64_t offset, stark_fri** out, uint8_t root[32], uint64_t* layers_out = nullptr, size_t layers_cap = 0) {
    STARK_REQUIRE(log_n <= ctx->two_adicity && log_n <= 30, "fri: 2^log_n does not divide p-1");
    STARK_REQUIRE(log_m <= log_n, "fri: polynomial has more coefficients than the domain has points");
    std::unique_ptr<stark_fri> f(new stark_fri());
    f->ctx = ctx; f->log_n = log_n; f->offset0 = offset % ctx->modulus;
    f->cur_log = log_n; f->cur_offset = f->offset0;
    f->coeffs = coeffs_padded; f->coeff_len = len;
    if (layers_out || layers_cap) f->sink = make_sink(ctx, log_n, layers_out, layers_cap);
    DevBufPtr ev = evaluate_on_coset(ctx, f->coeffs->as<uint32_t>(), log_m, log_n, offset);          // fri_commit.rs:78
    sink_push(f.get(), ev->as<uint32_t>(), (size_t)1 << log_n, true);     // ahead of the tree's launches: its CTAs go first
    LeafSource src; src.vals = ev->as<uint32_t>();
    auto t = tree_launch(ctx, ev, (size_t)1 << log_n, src);                                          // :79
    STARK_CUDA(cudaStreamSynchronize(ctx->stream));
    tree_take_root(t.get());
    if (root) words_to_bytes(t->root_words, root);
    f->trees.push_back(std::move(t));
    *out = f.release();
}
extern "C" int stark_fri_begin(stark_ctx* ctx, const uint64_t* coeffs, size_t n_coeffs, unsigned log_n, uint64_t offset,
                               stark_fri** out, uint8_t root[32]) {
    API_BEGIN
    STARK_REQUIRE(ctx && out && (coeffs || n_coeffs == 0), "fri_begin: null argument");
    CtxGuard g(ctx);
    check_offset(ctx, offset);
    size_t len = n_coeffs;                                               // Polynomial::new trims (ops.rs:19-37)
    while (len > 0 && coeffs[len - 1] % ctx->modulus == 0) len--;
    STARK_REQUIRE(log_n <= 30 && len <= ((size_t)1 << log_n), "fri: polynomial has more coefficients than the domain has points");
    unsigned log_m = ceil_log2(std::max<size_t>(len, 1));
    DevBufPtr c = upload_u64(ctx, coeffs, len, (size_t)1 << log_m);
    fri_begin_impl(ctx, c, len, log_m, log_n, offset, out, root);
    API_END
}
extern "C" int stark_fri_begin_to_host(stark_ctx* ctx, const uint64_t* coeffs, size_t n_coeffs, unsigned log_n, uint64_t offset,
                                       uint64_t* layers_out, size_t cap, stark_fri** out, uint8_t root[32]) {
    API_BEGIN
    STARK_REQUIRE(ctx && out && layers_out && (coeffs || n_coeffs == 0), "fri_begin_to_host: null argument");
    CtxGuard g(ctx);
    check_offset(ctx, offset);
    size_t len = n_coeffs;
    while (len > 0 && coeffs[len - 1] % ctx->modulus == 0) len--;
    STARK_REQUIRE(log_n <= 30 && len <= ((size_t)1 << log_n), "fri: polynomial has more coefficients than the domain has points");
    unsigned log_m = ceil_log2(std::max<size_t>(len, 1));
    DevBufPtr c = upload_u64(ctx, coeffs, len, (size_t)1 << log_m);
    fri_begin_impl(ctx, c, len, log_m, log_n, offset, out, root, layers_out, cap);
    API_END
}
extern "C" int stark_fri_layers_wait(const stark_fri* f) {
    API_BEGIN
    STARK_REQUIRE(f, "fri_layers_wait: null argument");
    if (f->sink) {
        CtxGuard g(f->ctx);
        sink_flush(const_cast<stark_fri*>(f));
        STARK_CUDA(cudaStreamSynchronize(f->sink->stream));
    }
    API_END
}
extern "C" size_t stark_fri_layer_host_offset(const stark_fri* f, size_t k) {
    return (f && f->sink && k < f->sink->offs.size()) ? f->sink->offs[k] : (size_t)-1;
}
extern "C" int stark_fri_begin_dev(stark_ctx* ctx, const stark_vec* coeffs, unsigned log_n, uint64_t offset, stark_fri** out,
                                   uint8_t root[32]) {
    API_BEGIN
    STARK_REQUIRE(ctx && out && coeffs && coeffs->ctx == ctx, "fri_begin_dev: bad argument");
    CtxGuard g(ctx);
    check_offset(ctx, offset);
    size_t len = 0;
    if (coeffs->n) {
        poly_degree(ctx, coeffs->buf->as<uint32_t>(), coeffs->n, ctx->d_result);
        STARK_CUDA(cudaStreamSynchronize(ctx->stream));
        len = (size_t)ctx->h_result->degree_plus1;
    }
    STARK_REQUIRE(log_n <= 30 && len <= ((size_t)1 << log_n), "fri: polynomial has more coefficients than the domain has points");
    unsigned log_m = ceil_log2(std::max<size_t>(len, 1));
    size_t m = (size_t)1 << log_m;
    DevBufPtr c = make_buf(m * 4, ctx->stream);       // private copy: folds overwrite it
    STARK_CUDA(cudaMemsetAsync(c->p, 0, m * 4, ctx->stream));
    if (len) STARK_CUDA(cudaMemcpyAsync(c->p, coeffs->buf->p, len * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    fri_begin_impl(ctx, c, len, log_m, log_n, offset, out, root);
    API_END
}
extern "C" int stark_fri_degree(const stark_fri* f, long long* degree) {
    API_BEGIN
    STARK_REQUIRE(f && degree, "fri_degree: null argument");
    *degree = (long long)f->coeff_len - 1;
    API_END
}
extern "C" int stark_fri_fold(stark_fri* f, uint64_t beta, uint8_t root[32]) {
    API_BEGIN
    STARK_REQUIRE(f, "fri_fold: null argument");
    stark_ctx* ctx = f->ctx;
    CtxGuard g(ctx);
    STARK_REQUIRE(f->cur_log >= 1, "fri_fold: the domain has one point left; the next layer would be empty and "
                                   "MerkleTree::root() would panic (fri_commit.rs:97-100, merkle/mod.rs:25)");
    const uint64_t p = ctx->modulus;
    const size_t n = (size_t)1 << f->cur_log, half = n >> 1;
    DevBufPtr old_coeffs;
    // evaluation space fused with the next tree (fri_commit.rs:53-65, :97)
    const stark_tree* prev = f->trees.back().get();
    DevBufPtr ev = make_buf(half * 4, ctx->stream);
    LeafSource src;
    src.prev = prev->leaves->as<uint32_t>(); src.fold_out = ev->as<uint32_t>(); src.half = half;
    // 1/2 and 1/(2 offset) without a modular inversion per layer (two Fermat powers were ~3 us of the ~6 us the host spends
    // between a layer's root and the next launch): (p + 1)/2 is 1/2, and 1/(2 o^2) = 2 * (1/(2 o))^2
    const uint64_t inv2 = (p + 1) / 2;
    if (f->half_over_offset == 0) f->half_over_offset = h_mul(inv2, h_inv(f->cur_offset, p), p);
    src.inv2_m = ctx->to_mont(inv2);
    src.sb_m = ctx->to_mont(h_mul(beta % p, f->half_over_offset, p));
    const bool had_tw = ctx->tw.count(f->cur_log) != 0;
    src.winv = ctx->twiddles(f->cur_log).inv();
    // coefficient space: exact degree of even + beta*odd (fri_commit.rs:32-50) -- a few CTAs of the tree's first launch
    if (f->coeff_len > 0) {
        size_t out_len = (f->coeff_len + 1) / 2;
        DevBufPtr nc = make_buf(out_len * 4, ctx->stream);
        coeff_fold_job(ctx, f->coeffs->as<uint32_t>(), f->coeff_len, ctx->to_mont(beta), nc->as<uint32_t>(), ctx->d_result,
                       merkle_first_launch_threads(half), src.job);
        old_coeffs = f->coeffs;              // read by the launch below: its stream-ordered release must come after it
        f->coeffs = nc;
    }
    // the host takes the tree's top over from the tail kernel instead of waiting for it to end (common.hpp: HostTop)
    static const bool early_top = [] { const char* e = getenv("STARK_EARLY_TOP"); return !e || atoi(e) != 0; }();
    uint32_t top_seq = early_top ? ++ctx->top_seq : 0;
    if (early_top && top_seq == 0) top_seq = ++ctx->top_seq;          // 0 means "off": skip it when the counter wraps
    const bool has_job = src.job.ctas != 0;
    if (top_seq && has_job) { src.job.top = ctx->d_top; src.job.seq = top_seq; }
    auto t = tree_launch(ctx, ev, half, src, nullptr, top_seq);
    {   // by-value layers: enqueued while the main stream is busy with the launches above (no API call between a root and the
        // next launch).  This layer from the start of its own tree (sink_push_fold); STARK_SINK_EARLY=0: the previous layer
        static const bool early = [] { const char* e = getenv("STARK_SINK_EARLY"); return !e || atoi(e) != 0; }();
        if (early) sink_push_fold(f, src, !had_tw);
        else sink_flush(f);
    }
    if (top_seq) wait_tree_top(ctx, top_seq, has_job, t->root_words);
    else {
        STARK_CUDA(cudaStreamSynchronize(ctx->stream));
        tree_take_root(t.get());
    }
    if (f->coeff_len > 0) f->coeff_len = (size_t)ctx->h_result->degree_plus1;
    if (root) words_to_bytes(t->root_words, root);
    f->trees.push_back(std::move(t));
    f->cur_log -= 1;
    f->cur_offset = h_mul(f->cur_offset, f->cur_offset, p);
    f->half_over_offset = h_mul(2 % p, h_mul(f->half_over_offset, f->half_over_offset, p), p);
    API_END
}
extern "C" int stark_fri_final(const stark_fri* f, uint64_t* value, size_t* final_poly_len) {
    API_BEGIN
    STARK_REQUIRE(f && value, "fri_final: null argument");
    CtxGuard g(f->ctx);
    sink_flush(const_cast<stark_fri*>(f));
    uint32_t c0 = 0;
    if (f->coeff_len > 0) {
        STARK_CUDA(cudaMemcpyAsync(&c0, f->coeffs->p, 4, cudaMemcpyDeviceToHost, f->ctx->stream));
        STARK_CUDA(cudaStreamSynchronize(f->ctx->stream));
    }
    *value = c0;                                            // fri_commit.rs:109-113
    if (final_poly_len) *final_poly_len = f->coeff_len;
    API_END
}
extern "C" size_t stark_fri_num_layers(const stark_fri* f) { return f ? f->trees.size() : 0; }
extern "C" size_t stark_fri_layer_len(const stark_fri* f, size_t k) { return (f && k < f->trees.size()) ? f->trees[k]->shape.n : 0; }
extern "C" int stark_fri_layer_read(const stark_fri* f, size_t k, size_t offset, size_t n, uint64_t* out) {
    API_BEGIN
    STARK_REQUIRE(f && (out || n == 0) && k < f->trees.size(), "fri_layer_read: bad argument");
    const stark_tree* t = f->trees[k].get();
    STARK_REQUIRE(offset <= t->shape.n && n <= t->shape.n - offset, "fri_layer_read: range out of bounds");
    CtxGuard g(f->ctx);
    download_u64(f->ctx, t->leaves->as<uint32_t>() + offset, n, out);
    API_END
}
extern "C" const stark_tree* stark_fri_layer_tree(const stark_fri* f, size_t k) { return (f && k < f->trees.size()) ? f->trees[k].get() : nullptr; }

// descriptors for one query index across all layers (fri_commit.rs:145-163)
static size_t fri_query_descs(const stark_fri* f, size_t index, size_t out_off, std::vector<OpenDesc>& d, size_t first_layer = 0) {
    for (size_t k = first_layer; k < f->trees.size(); k++) {
        const stark_tree* t = f->trees[k].get();
        STARK_REQUIRE(!t->external, "fri open: layer 0 was committed in leaf ranges on several GPUs; open it on the owners "
                                    "(stark_fri_open_layers with first_layer = 1 gives the rest)");
        size_t len = t->shape.n;
        size_t idx = index % len, sib = (idx + len / 2) % len;              // :152-153
        for (size_t which : {idx, sib}) {
            d.push_back(OpenDesc{t->leaves->as<uint32_t>(), t->nodes.as<uint32_t>(), len, which, out_off});
            out_off += 8 + merkle_path_len(len, which);
        }
    }
    return out_off;
}
extern "C" int stark_fri_open(const stark_fri* f, const uint64_t* indices, size_t n_idx, uint8_t* out, size_t cap, size_t* len) {
    API_BEGIN
    STARK_REQUIRE(f && len && (indices || n_idx == 0), "fri_open: null argument");
    std::vector<OpenDesc> d;
    d.reserve(n_idx * f->trees.size() * 2);
    size_t total = 0;
    for (size_t q = 0; q < n_idx; q++) total = fri_query_descs(f, (size_t)indices[q], total, d);
    *len = total;
    if (!out) return ST_OK;
    STARK_REQUIRE(cap >= total, "fri_open: buffer too small");
    CtxGuard g(f->ctx);
    open_records(f->ctx, d, total, out);
    API_END
}
extern "C" int stark_fri_open_layers(const stark_fri* f, size_t first_layer, const uint64_t* indices, size_t n_idx, uint8_t* out,
                                     size_t cap, size_t* len) {
    API_BEGIN
    STARK_REQUIRE(f && len && (indices || n_idx == 0) && first_layer <= f->trees.size(), "fri_open_layers: bad argument");
    std::vector<OpenDesc> d;
    size_t total = 0;
    for (size_t q = 0; q < n_idx; q++) total = fri_query_descs(f, (size_t)indices[q], total, d, first_layer);
    *len = total;
    if (!out) return ST_OK;
    STARK_REQUIRE(cap >= total, "fri_open_layers: buffer too small");
    CtxGuard g(f->ctx);
    open_records(f->ctx, d, total, out);
    API_END
}
// Layer 0 was evaluated and hashed elsewhere (four-step LDE + leaf-range subtrees on several GPUs): adopt the
// gathered evaluations and the combined root; folding and the later layers proceed as in stark_fri_begin.
extern "C" int stark_fri_begin_external(stark_ctx* ctx, const stark_vec* coeffs, unsigned log_n, uint64_t offset,
                                        const stark_vec* layer0, const uint8_t root0[32], stark_fri** out) {
    API_BEGIN
    STARK_REQUIRE(ctx && out && coeffs && layer0 && root0 && coeffs->ctx == ctx && layer0->ctx == ctx, "fri_begin_external: bad argument");
    CtxGuard g(ctx);
    check_offset(ctx, offset);
    STARK_REQUIRE(log_n <= 30 && log_n <= ctx->two_adicity && layer0->n == ((size_t)1 << log_n), "fri_begin_external: layer 0 must have 2^log_n evaluations");
    size_t len = 0;
    if (coeffs->n) {
        poly_degree(ctx, coeffs->buf->as<uint32_t>(), coeffs->n, ctx->d_result);
        STARK_CUDA(cudaStreamSynchronize(ctx->stream));
        len = (size_t)ctx->h_result->degree_plus1;
    }
    STARK_REQUIRE(len <= layer0->n, "fri: polynomial has more coefficients than the domain has points");
    std::unique_ptr<stark_fri> f(new stark_fri());
    f->ctx = ctx; f->log_n = log_n; f->offset0 = offset % ctx->modulus;
    f->cur_log = log_n; f->cur_offset = f->offset0;
    size_t m = (size_t)1 << ceil_log2(std::max<size_t>(len, 1));
    f->coeffs = make_buf(m * 4, ctx->stream);                 // private copy: folds replace it
    STARK_CUDA(cudaMemsetAsync(f->coeffs->p, 0, m * 4, ctx->stream));
    if (len) STARK_CUDA(cudaMemcpyAsync(f->coeffs->p, coeffs->buf->p, len * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    f->coeff_len = len;
    std::unique_ptr<stark_tree> t(new stark_tree());
    t->ctx = ctx; t->leaves = layer0->buf; t->shape = TreeShape::make(layer0->n); t->external = true;
    for (int i = 0; i < 8; i++)
        t->root_words[i] = ((uint32_t)root0[4 * i] << 24) | ((uint32_t)root0[4 * i + 1] << 16) | ((uint32_t)root0[4 * i + 2] << 8) | root0[4 * i + 3];
    f->trees.push_back(std::move(t));
    *out = f.release();
    API_END
}
extern "C" void stark_fri_destroy(stark_fri* f) {
    if (!f) return;
    cudaSetDevice(f->ctx->device);
    delete f;
}

// ---- whole-loop API with the library's Channel ----
static void send_root(Channel& ch, const stark_tree* t) {
    uint8_t r[32]; words_to_bytes(t->root_words, r);
    std::string h = HostSha256::hex(r, 32);
    ch.send(reinterpret_cast<const uint8_t*>(h.data()), 64);    // root().as_bytes(): 64 ASCII hex chars (fri_verify.rs:24-25)
}
static int fri_commit_loop(stark_fri* f, stark_channel* chan, bool send_first = true) {
    Channel& ch = chan->ch;
    if (send_first) send_root(ch, f->trees[0].get());                        // fri_commit.rs:86
    while ((long long)f->coeff_len - 1 >= 1) {                               // :89
        uint64_t beta;
        STARK_REQUIRE(ch.receive_random_field_element(&beta), "channel: receive before send");   // :91
        int rc = stark_fri_fold(f, beta, nullptr);                           // :94-97
        if (rc != ST_OK) return rc;
        send_root(ch, f->trees.back().get());                                // :100
    }
    uint64_t fv; size_t fl;
    int rc = stark_fri_final(f, &fv, &fl);
    if (rc != ST_OK) return rc;
    uint8_t b[8]; be8(fv, b);
    ch.send(b, 8);                                                           // :114
    return ST_OK;
}
extern "C" int stark_fri_commit(stark_ctx* ctx, const uint64_t* coeffs, size_t n_coeffs, unsigned log_n, uint64_t offset,
                                stark_channel* ch, stark_fri** out) {
    STARK_API_GUARD_NULL(ctx && ch && out);
    int rc = stark_fri_begin(ctx, coeffs, n_coeffs, log_n, offset, out, nullptr);
    if (rc != ST_OK) return rc;
    API_BEGIN
    rc = fri_commit_loop(*out, ch);
    if (rc != ST_OK) { stark_fri_destroy(*out); *out = nullptr; return rc; }
    API_END
}
static int fri_commit_to_host_impl(stark_ctx* ctx, const uint64_t* coeffs, size_t n_coeffs, unsigned log_n, uint64_t offset,
                                   stark_channel* ch, uint64_t* layers_out, size_t cap, stark_fri** out, bool wait) {
    STARK_API_GUARD_NULL(ctx && ch && out && layers_out);
    int rc = stark_fri_begin_to_host(ctx, coeffs, n_coeffs, log_n, offset, layers_out, cap, out, nullptr);
    if (rc != ST_OK) return rc;
    API_BEGIN
    rc = fri_commit_loop(*out, ch);
    if (rc == ST_OK && wait) rc = stark_fri_layers_wait(*out);               // by value: complete on return
    if (rc != ST_OK) { stark_fri_destroy(*out); *out = nullptr; return rc; }
    API_END
}
extern "C" int stark_fri_commit_to_host(stark_ctx* ctx, const uint64_t* coeffs, size_t n_coeffs, unsigned log_n, uint64_t offset,
                                        stark_channel* ch, uint64_t* layers_out, size_t cap, stark_fri** out) {
    return fri_commit_to_host_impl(ctx, coeffs, n_coeffs, log_n, offset, ch, layers_out, cap, out, true);
}
extern "C" int stark_fri_commit_to_host_async(stark_ctx* ctx, const uint64_t* coeffs, size_t n_coeffs, unsigned log_n, uint64_t offset,
                                              stark_channel* ch, uint64_t* layers_out, size_t cap, stark_fri** out) {
    return fri_commit_to_host_impl(ctx, coeffs, n_coeffs, log_n, offset, ch, layers_out, cap, out, false);
}
extern "C" int stark_fri_commit_dev(stark_ctx* ctx, const stark_vec* coeffs, unsigned log_n, uint64_t offset,
                                    stark_channel* ch, stark_fri** out) {
    STARK_API_GUARD_NULL(ctx && ch && out && coeffs);
    int rc = stark_fri_begin_dev(ctx, coeffs, log_n, offset, out, nullptr);
    if (rc != ST_OK) return rc;
    API_BEGIN
    rc = fri_commit_loop(*out, ch);
    if (rc != ST_OK) { stark_fri_destroy(*out); *out = nullptr; return rc; }
    API_END
}
// feeds one query's records to the channel in the reference's order (fri_commit.rs:145-163)
static void send_query_records(const stark_fri* f, const uint8_t* rec, Channel& ch, size_t index, size_t first_layer = 0) {
    size_t off = 0;
    for (size_t k = first_layer; k < f->trees.size(); k++) {
        auto& tp = f->trees[k];
        size_t len = tp->shape.n;
        size_t idx = index % len, sib = (idx + len / 2) % len;
        if (len == 1) ch.send(rec + off, 8);                                 // :147-149 (then falls through, as written)
        for (size_t which : {idx, sib}) {
            size_t pl = merkle_path_len(len, which);
            ch.send(rec + off, 8);                                           // :156 / :161
            ch.send(rec + off + 8, pl);                                      // :157-158 / :162-163
            off += 8 + pl;
        }
    }
}
// One index, all layers: the layer table travels as kernel arguments, the records land in mapped host memory.
static const uint8_t* open_one_index(const stark_fri* f, size_t index, size_t* total_out) {
    stark_ctx* ctx = f->ctx;
    STARK_REQUIRE(f->trees.size() <= (size_t)FRI_MAX_LAYERS, "fri: too many layers");
    FriOpenArgs a{};
    size_t total = 0;
    for (size_t k = 0; k < f->trees.size(); k++) {
        const stark_tree* t = f->trees[k].get();
        STARK_REQUIRE(!t->external, "fri open: layer 0 was committed in leaf ranges on several GPUs; open it on the owners");
        a.layers[k] = FriLayerDesc{t->leaves->as<uint32_t>(), t->nodes.as<uint32_t>(), t->shape.n};
        size_t len = t->shape.n, idx = index % len, sib = (idx + len / 2) % len;
        a.rec_off[2 * k] = (uint32_t)total;
        total += 8 + merkle_path_len(len, idx);
        a.rec_off[2 * k + 1] = (uint32_t)total;
        total += 8 + merkle_path_len(len, sib);
    }
    a.n_layers = (unsigned)f->trees.size(); a.first = 0; a.index = index;
    ctx->pin_out.ensure(total);
    a.out = static_cast<uint8_t*>(ctx->pin_out.d);
    fri_open_one(ctx, a);
    STARK_CUDA(cudaStreamSynchronize(ctx->stream));
    *total_out = total;
    return static_cast<const uint8_t*>(ctx->pin_out.h);
}
extern "C" int stark_decommit_fri_layers(const stark_fri* f, size_t index, stark_channel* ch) {
    API_BEGIN
    STARK_REQUIRE(f && ch, "decommit_fri_layers: null argument");
    CtxGuard g(f->ctx);
    size_t total = 0;
    const uint8_t* rec = open_one_index(f, index, &total);      // valid until the next opening on this context
    send_query_records(f, rec, ch->ch, index);
    API_END
}
// A run of queries on one proof through the resident opening server (merkle.cu): no launch and no stream
// synchronisation per query, one 64-bit word each way in mapped pinned memory.
#ifndef STARK_OPEN_SERVER
#define STARK_OPEN_SERVER 1
#endif
struct FriOpenServer {
    stark_ctx* ctx = nullptr;
    FriServerBox* h_box = nullptr;
    unsigned long long seq = 0;
    bool running = false;
    // The kernel leaves after this long without a request.  The first request is posted BEFORE the launch, so a kernel
    // that only starts once the host is blocked inside the launch call (CUDA_LAUNCH_BLOCKING=1, a profiler or debugger that
    // serialises launches) still answers it and then idles for this long at most -- 2 ms, ~40x the host's time between two
    // queries (transcript hashing, ~45 us) -- before the host falls back to one launch per query.  (Round 1 waited 250 ms
    // here: 99.8 % of a profiled smoke run)
    static constexpr unsigned long long kIdleNs = 2000000ull;
    static bool usable(const stark_fri* f) {
        // STARK_OPEN_SERVER=0 in the environment: one launch per query instead (bench.py --profile-mode sets it so that the
        // committed launch list shows every opening)
        static const bool enabled = [] { const char* e = getenv("STARK_OPEN_SERVER"); return STARK_OPEN_SERVER && !(e && e[0] == '0'); }();
        if (!enabled || f->trees.empty() || f->trees.size() > (size_t)FRI_MAX_LAYERS) return false;
        const size_t n0 = f->trees[0]->shape.n;
        if (n0 > ((size_t)1 << 32)) return false;
        for (auto& t : f->trees)
            if (t->external || t->shape.n == 0 || n0 % t->shape.n != 0) return false;
        return true;
    }
    // launches the server with request 1 = `index0` already posted
    void start(const stark_fri* f, size_t index0) {
        ctx = f->ctx;
        FriOpenArgs a{};
        size_t cap = 0;
        for (size_t k = 0; k < f->trees.size(); k++) {
            const stark_tree* t = f->trees[k].get();
            a.layers[k] = FriLayerDesc{t->leaves->as<uint32_t>(), t->nodes.as<uint32_t>(), t->shape.n};
            cap += 2 * (8 + 32 * (size_t)t->shape.depth);
        }
        a.n_layers = (unsigned)f->trees.size(); a.first = 0;
        STARK_CUDA(cudaStreamSynchronize(ctx->stream));            // the pinned buffers may still be in use by an earlier opening
        ctx->pin_out.ensure(cap);
        ctx->pin_desc.ensure(sizeof(FriServerBox));
        a.out = static_cast<uint8_t*>(ctx->pin_out.d);
        h_box = static_cast<FriServerBox*>(ctx->pin_desc.h);
        seq = 1;
        __atomic_store_n(&h_box->done, 0ull, __ATOMIC_RELEASE);
        __atomic_store_n(&h_box->req, (seq << 32) | (unsigned long long)index0, __ATOMIC_RELEASE);
        unsigned long long idle = kIdleNs;
        if (const char* e = getenv("STARK_OPEN_SERVER_IDLE_NS")) { unsigned long long v = strtoull(e, nullptr, 10); if (v) idle = v; }   // tests
        fri_open_server_launch(ctx, a, static_cast<FriServerBox*>(ctx->pin_desc.d), idle);
        running = true;
    }
    // waits for the answer to the request in flight; nullptr = the server has left (the caller launches per query)
    const uint8_t* wait(size_t* total) {
        const auto t0 = std::chrono::steady_clock::now();
        unsigned long long d;
        for (unsigned spin = 0;; spin++) {
            d = __atomic_load_n(&h_box->done, __ATOMIC_ACQUIRE);
            if ((d >> 32) == seq) break;
            if ((spin & 0x3ffu) == 0x3ffu) {
                if (cudaStreamQuery(ctx->stream) != cudaErrorNotReady) {
                    running = false;
                    STARK_CUDA(cudaStreamSynchronize(ctx->stream));
                    d = __atomic_load_n(&h_box->done, __ATOMIC_ACQUIRE);          // it may have answered on its way out
                    if ((d >> 32) == seq) break;
                    return nullptr;
                }
                if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(5)) { stop(); throw StarkError(ST_CUDA, "opening server timed out"); }
            }
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
        *total = (size_t)(d & 0xffffffffull);
        return static_cast<const uint8_t*>(ctx->pin_out.h);
    }
    const uint8_t* open(size_t index0, size_t* total) {
        seq++;
        __atomic_store_n(&h_box->req, (seq << 32) | (unsigned long long)index0, __ATOMIC_RELEASE);
        return wait(total);
    }
    void stop() {
        if (!running) return;
        running = false;
        __atomic_store_n(&h_box->req, ~0ull, __ATOMIC_RELEASE);
        cudaStreamSynchronize(ctx->stream);
    }
    ~FriOpenServer() { stop(); }
};
extern "C" int stark_decommit_fri(const stark_fri* f, size_t num_queries, size_t max_index, stark_channel* ch) {
    API_BEGIN
    STARK_REQUIRE(f && ch, "decommit_fri: null argument");
    if (num_queries >= 2 && FriOpenServer::usable(f)) {
        CtxGuard g(f->ctx);
        FriOpenServer srv;
        const size_t n0 = f->trees[0]->shape.n;
        for (size_t q = 0; q < num_queries; q++) {                           // :175-178
            uint64_t idx;
            STARK_REQUIRE(ch->ch.receive_random_int(0, max_index, true, &idx), "channel: receive before send");
            size_t total = 0;
            const uint8_t* rec = nullptr;
            if (q == 0) { srv.start(f, (size_t)idx % n0); rec = srv.wait(&total); }      // every layer length divides n0
            else if (srv.running) rec = srv.open((size_t)idx % n0, &total);
            if (!rec) rec = open_one_index(f, (size_t)idx, &total);
            send_query_records(f, rec, ch->ch, (size_t)idx);
        }
        srv.stop();
        STARK_CUDA(cudaGetLastError());
        return ST_OK;
    }
    for (size_t q = 0; q < num_queries; q++) {                               // :175-178
        uint64_t idx;
        STARK_REQUIRE(ch->ch.receive_random_int(0, max_index, true, &idx), "channel: receive before send");
        int rc = stark_decommit_fri_layers(f, (size_t)idx, ch);
        if (rc != ST_OK) return rc;
    }
    API_END
}

// ======================================================================================= shared with stark101.cu
namespace starkb200 {
DevBufPtr api_upload_u64(stark_ctx* ctx, const uint64_t* host, size_t n) { return upload_u64(ctx, host, n); }
DevBufPtr api_lde_on_coset(stark_ctx* ctx, const uint32_t* evals, unsigned log_n, uint64_t offset_in, unsigned log_blowup, uint64_t offset_out) {
    return lde_on_coset(ctx, evals, log_n, offset_in, log_blowup, offset_out);
}
DevBufPtr api_interpolate_on_coset(stark_ctx* ctx, const uint32_t* evals, unsigned log_n, uint64_t offset) {
    return interpolate_on_coset(ctx, evals, log_n, offset);
}
std::unique_ptr<stark_tree> api_tree_commit(stark_ctx* ctx, DevBufPtr leaves, size_t n) { return tree_commit(ctx, std::move(leaves), n); }
void api_send_root(Channel& ch, const stark_tree* t) { send_root(ch, t); }
void api_open_and_send(const stark_tree* t, size_t idx, Channel& ch) {
    STARK_REQUIRE(idx < t->shape.n, "open: leaf index out of range");
    size_t pl = merkle_path_len(t->shape.n, idx);
    std::vector<OpenDesc> d(1);
    d[0] = OpenDesc{t->leaves->as<uint32_t>(), t->nodes.as<uint32_t>(), t->shape.n, idx, 0};
    std::vector<uint8_t> rec(8 + pl);
    open_records(t->ctx, d, rec.size(), rec.data());
    ch.send(rec.data(), 8);
    ch.send(rec.data() + 8, pl);
}
void api_set_error(const std::string& s) { set_error(s); }
void api_send_root_bytes(Channel& ch, const uint8_t root[32]) {
    std::string h = HostSha256::hex(root, 32);
    ch.send(reinterpret_cast<const uint8_t*>(h.data()), 64);
}
// multi.cu: a tree over leaf values whose root lands in `result` (device alias of mapped host memory) -- no sync here
std::unique_ptr<stark_tree> api_tree_launch_values(stark_ctx* ctx, DevBufPtr leaves, size_t n, HostResult* result) {
    STARK_REQUIRE(n >= 1 && n <= ((size_t)1 << 32), "merkle: leaf count out of range");
    LeafSource src; src.vals = leaves->as<uint32_t>();
    return tree_launch(ctx, std::move(leaves), n, src, result);
}
int api_fri_commit_loop(stark_fri* f, stark_channel* chan) { return fri_commit_loop(f, chan); }
int api_fri_commit_loop_resume(stark_fri* f, stark_channel* chan) { return fri_commit_loop(f, chan, false); }
// One fold whose tree is built elsewhere (leaf ranges of the new layer hashed on several GPUs, multi.cu): the evaluation-space
// fold of the WHOLE layer and the coefficient-space fold with its exact degree are enqueued here; the caller hashes its
// range, synchronises the stream, and hands the layer back with the combined root through api_fri_adopt_layer.
DevBufPtr api_fri_fold_values(stark_fri* f, uint64_t beta) {
    stark_ctx* ctx = f->ctx;
    STARK_REQUIRE(f->cur_log >= 1, "fri_fold: the domain has one point left");
    const uint64_t p = ctx->modulus;
    const size_t half = ((size_t)1 << f->cur_log) >> 1;
    DevBufPtr ev = make_buf(half * 4, ctx->stream);
    LeafSource src;
    src.prev = f->trees.back()->leaves->as<uint32_t>(); src.fold_out = ev->as<uint32_t>(); src.half = half;
    const uint64_t inv2 = (p + 1) / 2;
    if (f->half_over_offset == 0) f->half_over_offset = h_mul(inv2, h_inv(f->cur_offset, p), p);
    src.inv2_m = ctx->to_mont(inv2);
    src.sb_m = ctx->to_mont(h_mul(beta % p, f->half_over_offset, p));
    src.winv = ctx->twiddles(f->cur_log).inv();
    fri_fold(ctx, src);
    if (f->coeff_len > 0) {
        const size_t out_len = (f->coeff_len + 1) / 2;
        DevBufPtr nc = make_buf(out_len * 4, ctx->stream);
        coeff_fold(ctx, f->coeffs->as<uint32_t>(), f->coeff_len, ctx->to_mont(beta), nc->as<uint32_t>(), ctx->d_result);
        f->coeffs = nc;            // (the old block goes back to the same stream: its release is ordered after the launch)
    }
    return ev;
}
// after a stream synchronisation: the degree the coefficient fold published, and the layer as a tree whose levels live elsewhere
void api_fri_adopt_layer(stark_fri* f, DevBufPtr layer, const uint8_t root[32]) {
    stark_ctx* ctx = f->ctx;
    const uint64_t p = ctx->modulus;
    if (f->coeff_len > 0) f->coeff_len = (size_t)ctx->h_result->degree_plus1;
    std::unique_ptr<stark_tree> t(new stark_tree());
    t->ctx = ctx; t->shape = TreeShape::make(((size_t)1 << f->cur_log) >> 1); t->leaves = std::move(layer); t->external = true;
    for (int i = 0; i < 8; i++)
        t->root_words[i] = ((uint32_t)root[4 * i] << 24) | ((uint32_t)root[4 * i + 1] << 16) | ((uint32_t)root[4 * i + 2] << 8) | root[4 * i + 3];
    f->trees.push_back(std::move(t));
    f->cur_log -= 1;
    f->cur_offset = h_mul(f->cur_offset, f->cur_offset, p);
    f->half_over_offset = h_mul(2 % p, h_mul(f->half_over_offset, f->half_over_offset, p), p);
}
// a batch of (element, path) records in one launch (BE8(value) || path each)
void api_open_records(stark_ctx* ctx, const std::vector<OpenDesc>& descs, size_t total_bytes, uint8_t* host_out) {
    open_records(ctx, descs, total_bytes, host_out);
}
void api_send_query_records(const stark_fri* f, const uint8_t* rec, Channel& ch, size_t index, size_t first_layer) {
    send_query_records(f, rec, ch, index, first_layer);
}
// fri_commit for a polynomial whose evaluations on the FRI domain already exist (the composition polynomial of
// stark101.cu is computed on the coset): layer 0 is `evals` as it stands -- no second transform, no copy of the
// coefficients, which are consumed (the folds replace them anyway).  Same messages as stark_fri_commit_dev.
int api_fri_commit_evaluated(stark_ctx* ctx, DevBufPtr coeffs, size_t n_coeffs, DevBufPtr evals, unsigned log_n, uint64_t offset,
                             stark_channel* chan, stark_fri** out) {
    STARK_REQUIRE(log_n <= ctx->two_adicity && log_n <= 30 && n_coeffs <= ((size_t)1 << log_n), "fri: bad sizes");
    check_offset(ctx, offset);
    size_t len = 0;
    if (n_coeffs) {
        poly_degree(ctx, coeffs->as<uint32_t>(), n_coeffs, ctx->d_result);
        STARK_CUDA(cudaStreamSynchronize(ctx->stream));
        len = (size_t)ctx->h_result->degree_plus1;
    }
    std::unique_ptr<stark_fri> f(new stark_fri());
    f->ctx = ctx; f->log_n = log_n; f->offset0 = offset % ctx->modulus;
    f->cur_log = log_n; f->cur_offset = f->offset0;
    f->coeffs = std::move(coeffs); f->coeff_len = len;
    LeafSource src; src.vals = evals->as<uint32_t>();
    auto t = tree_launch(ctx, evals, (size_t)1 << log_n, src);
    STARK_CUDA(cudaStreamSynchronize(ctx->stream));
    tree_take_root(t.get());
    f->trees.push_back(std::move(t));
    int rc = fri_commit_loop(f.get(), chan);
    if (rc != ST_OK) return rc;
    *out = f.release();
    return ST_OK;
}
}  // namespace starkb200
