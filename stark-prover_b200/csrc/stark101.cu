// stark101.cu — the FibonacciSq (STARK-101) prover on top of the hot path.
//
// BUILD-DEFINED: src/prover/*, src/trace/*, src/composition/* are empty files in the reference, so there
// is nothing to match there; the statement, constraints and transcript order are written down in
// DESIGN.md ("cfg1") and pinned by the oracle's literal (Horner / Lagrange / long-division) restatement.
// Every primitive it calls — interpolate, evaluate over the coset, MerkleTree::new, Channel,
// fri_commit, decommit — is the reference-backed one.
//
//   trace      a0 = 1, a1 = given, a_{n+2} = a_{n+1}^2 + a_n^2, T-1 rows (T = 2^log_trace)
//   f          interpolant of the trace over g^i, i < T-1 (degree <= T-2)
//   p0         (f(x) - 1) / (x - 1)
//   p1         (f(x) - a_{T-2}) / (x - g^(T-2))
//   p2         (f(g^2 x) - f(g x)^2 - f(x)^2) (x-g^(T-3))(x-g^(T-2))(x-g^(T-1)) / (x^T - 1)
//   CP         alpha0 p0 + alpha1 p1 + alpha2 p2, committed by fri_commit on the coset w*<h>
#include <string.h>

#include "../../include/stark_b200.h"
#include "handles.hpp"

using namespace starkb200;

namespace starkb200 {

// helpers implemented in api.cu
DevBufPtr api_upload_u64(stark_ctx* ctx, const uint64_t* host, size_t n);
DevBufPtr api_lde_on_coset(stark_ctx* ctx, const uint32_t* evals, unsigned log_n, uint64_t offset_in, unsigned log_blowup, uint64_t offset_out);
DevBufPtr api_interpolate_on_coset(stark_ctx* ctx, const uint32_t* evals, unsigned log_n, uint64_t offset);
std::unique_ptr<stark_tree> api_tree_commit(stark_ctx* ctx, DevBufPtr leaves, size_t n);
void api_send_root(Channel& ch, const stark_tree* t);
void api_open_and_send(const stark_tree* t, size_t idx, Channel& ch);
void api_set_error(const std::string& s);
int api_fri_commit_evaluated(stark_ctx* ctx, DevBufPtr coeffs, size_t n_coeffs, DevBufPtr evals, unsigned log_n, uint64_t offset,
                             stark_channel* chan, stark_fri** out);

struct FibSqDev {
    uint32_t offset;          // w (canonical)
    uint32_t one;             // 1
    uint32_t x_last;          // g^(T-2)
    uint32_t ex[3];           // g^(T-3), g^(T-2), g^(T-1)
    uint32_t last_value;      // a_{T-2}
    uint32_t alpha_m[3];      // Montgomery form
    uint32_t blow;            // 2^log_blowup
    const uint32_t* zinv_m;   // blow entries: (x^T - 1)^-1 for i mod blow, Montgomery form
    uint32_t start;           // domain index of the first point handled (a multiple of blow); 0 for the whole coset
    uint32_t f_mask;          // f is indexed (local + k*blow) & f_mask: N-1 for the whole coset (wraps), ~0 for a range with a halo
};

// d0[i] = x_i - 1, d1[i] = x_i - g^(T-2),  x_i = w h^i
__global__ void fibsq_denoms_kernel(FibSqDev q, PowTable tw, uint32_t* d0, uint32_t* d1, size_t n, FieldParams fp) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t x = mont_mul(pow_lookup(tw, q.start + (uint32_t)i, fp), q.offset, fp);
    d0[i] = fsub(x, q.one, fp);
    d1[i] = fsub(x, q.x_last, fp);
}
// i0 = 1/d0, i1 = 1/d1 (canonical) -> CP(x_i)
__global__ void fibsq_combine_kernel(FibSqDev q, PowTable tw, const uint32_t* f, const uint32_t* i0, const uint32_t* i1,
                                     uint32_t* cp, size_t n, FieldParams fp) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t x = mont_mul(pow_lookup(tw, q.start + (uint32_t)i, fp), q.offset, fp);
    uint32_t fx = f[i], fgx = f[(i + q.blow) & q.f_mask], fg2x = f[(i + 2 * (size_t)q.blow) & q.f_mask];   // g = h^blow
    uint32_t p0 = mont_mul(to_mont(fsub(fx, q.one, fp), fp), i0[i], fp);
    uint32_t p1 = mont_mul(to_mont(fsub(fx, q.last_value, fp), fp), i1[i], fp);
    uint32_t sq1 = mont_mul(to_mont(fgx, fp), fgx, fp), sq0 = mont_mul(to_mont(fx, fp), fx, fp);
    uint32_t num = fsub(fsub(fg2x, sq1, fp), sq0, fp);
    uint32_t e = mont_mul(to_mont(fsub(x, q.ex[0], fp), fp), fsub(x, q.ex[1], fp), fp);
    e = mont_mul(to_mont(e, fp), fsub(x, q.ex[2], fp), fp);
    uint32_t p2 = mont_mul(to_mont(num, fp), e, fp);
    p2 = mont_mul(p2, __ldg(q.zinv_m + ((q.start + i) & (q.blow - 1))), fp);
    uint32_t r = fadd(mont_mul(p0, q.alpha_m[0], fp), mont_mul(p1, q.alpha_m[1], fp), fp);
    cp[i] = fadd(r, mont_mul(p2, q.alpha_m[2], fp), fp);
}

// partial[b] = sum over block b of a[i] * w^i  (mod p): the device half of "choose the T-th trace value so that
// the top coefficient of the interpolant vanishes"
__global__ void dot_powers_kernel(const uint32_t* a, size_t n, PowTable tw, uint32_t* partial, FieldParams fp) {
    unsigned long long acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        acc += mont_mul(a[i], pow_lookup(tw, (uint32_t)i, fp), fp);          // < p each; at most 2^32 terms fit
    __shared__ unsigned long long sh[256];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = (uint32_t)(sh[0] % fp.p);
}
// sum_i a[i] * w_n^i over the first `count` entries of a device vector
uint64_t dot_powers(stark_ctx* ctx, const uint32_t* a, size_t count, unsigned log_n) {
    const unsigned blocks = 256;
    DevBuf part(blocks * 4, ctx->stream);
    dot_powers_kernel<<<blocks, 256, 0, ctx->stream>>>(a, count, ctx->twiddles(log_n).fwd(), part.as<uint32_t>(), ctx->fp);
    ctx->launches++;
    STARK_CUDA(cudaGetLastError());
    uint32_t h[blocks];
    STARK_CUDA(cudaMemcpyAsync(h, part.p, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    STARK_CUDA(cudaStreamSynchronize(ctx->stream));
    uint64_t s = 0;
    for (unsigned b = 0; b < blocks; b++) s = (s + h[b]) % ctx->modulus;
    return s;
}

// CP on the points start .. start+count-1 of the coset.  Whole coset: start = 0, count = N, f_eval holds N values.
// A range (several GPUs, each a contiguous block): f_eval holds count + 2*blow values, the block followed by the first
// 2*blow values of the next block (f(g x), f(g^2 x) of the last points), and start is a multiple of blow.
void fibsq_composition(stark_ctx* ctx, const uint32_t* f_eval, const FibSqParams& prm, uint32_t* cp_eval, size_t start, size_t count) {
    const uint64_t p = ctx->modulus;
    const unsigned log_N = prm.log_trace + prm.log_blowup;
    const size_t T = (size_t)1 << prm.log_trace, blow = (size_t)1 << prm.log_blowup;
    size_t N = (size_t)1 << log_N;
    if (count == 0) { start = 0; count = N; }
    STARK_REQUIRE(start % blow == 0 && start + count <= N, "fibsq_composition: bad range");
    const bool whole = count == N;
    N = count;
    const uint64_t g = ctx->root_of_unity(prm.log_trace), h = ctx->root_of_unity(log_N), w = prm.offset % p;
    FibSqDev q{};
    q.offset = (uint32_t)w; q.one = (uint32_t)(1 % p);
    q.x_last = (uint32_t)h_pow(g, T - 2, p);
    q.ex[0] = (uint32_t)h_pow(g, T - 3, p); q.ex[1] = q.x_last; q.ex[2] = (uint32_t)h_pow(g, T - 1, p);
    q.last_value = (uint32_t)(prm.last_value % p);
    for (int k = 0; k < 3; k++) q.alpha_m[k] = ctx->to_mont(prm.alpha[k]);
    q.blow = (uint32_t)blow;
    q.start = (uint32_t)start; q.f_mask = whole ? (uint32_t)(N - 1) : 0xffffffffu;
    // x^T - 1 takes `blow` distinct values on the coset: w^T (h^T)^i
    std::vector<uint32_t> zinv(blow);
    uint64_t wT = h_pow(w, T, p), hT = h_pow(h, T, p), cur = wT;
    for (size_t i = 0; i < blow; i++) { zinv[i] = ctx->to_mont(h_inv((cur + p - 1) % p, p)); cur = h_mul(cur, hT, p); }
    DevBuf d_z(blow * 4, ctx->stream), d0(N * 4, ctx->stream), d1(N * 4, ctx->stream);
    STARK_CUDA(cudaMemcpyAsync(d_z.p, zinv.data(), blow * 4, cudaMemcpyHostToDevice, ctx->stream));
    STARK_CUDA(cudaStreamSynchronize(ctx->stream));
    q.zinv_m = d_z.as<uint32_t>();
    PowTable tw = ctx->twiddles(log_N).fwd();
    unsigned blocks = (unsigned)((N + 255) / 256);
    fibsq_denoms_kernel<<<blocks, 256, 0, ctx->stream>>>(q, tw, d0.as<uint32_t>(), d1.as<uint32_t>(), N, ctx->fp);
    ctx->launches++;
    batch_inverse(ctx, d0.as<uint32_t>(), nullptr, d0.as<uint32_t>(), N);
    batch_inverse(ctx, d1.as<uint32_t>(), nullptr, d1.as<uint32_t>(), N);
    fibsq_combine_kernel<<<blocks, 256, 0, ctx->stream>>>(q, tw, f_eval, d0.as<uint32_t>(), d1.as<uint32_t>(), cp_eval, N, ctx->fp);
    ctx->launches++;
    STARK_CUDA(cudaGetLastError());
}

}  // namespace starkb200


// src/trace + the interpolation: the T values whose size-T interpolant over <g> is Polynomial::interpolate of the
// T-1 trace rows (device buffer, canonical u32), and a_{T-2}
static DevBufPtr fibsq_trace_column(stark_ctx* ctx, uint64_t a1, unsigned log_trace, uint64_t* last_value) {
    const uint64_t p = ctx->modulus;
    const size_t T = (size_t)1 << log_trace, rows = T - 1;
    const uint64_t g = ctx->root_of_unity(log_trace);
    // the recurrence is sequential, it stays on the host
    // (8M dependent steps at 2^23 rows.  Montgomery form with R = 2^64 and no correction on the dependency chain:
    // REDC(x^2) = hi(x^2) + p - hi(m p) lies in (0, p + 16) for x < 2^33, so a_{n+2} = REDC(a_{n+1}^2) + REDC(a_n^2)
    // stays below 2^33 unreduced and a step costs mul -> mul -> mulhi -> two adds (~12 cycles) instead of the 32-bit
    // form's compare/select after both the square and the sum; a_n^2 is carried over from the previous step.  The
    // canonical value leaves Montgomery form off the chain and is written as u32 straight into the pinned staging buffer
    // the context keeps.)
    ctx->pin_stage.ensure(T * 4);
    uint32_t* a = static_cast<uint32_t*>(ctx->pin_stage.h);
    {
        typedef unsigned __int128 u128;
        uint64_t pinv = p;                                         // p^-1 mod 2^64 (Newton)
        for (int i = 0; i < 6; i++) pinv *= 2 - p * pinv;
        auto redc_sq = [p, pinv](uint64_t x) {                     // x^2 / R, in (0, p + hi(x^2)]
            u128 t = (u128)x * x;
            uint64_t m = (uint64_t)t * pinv;
            return (uint64_t)(t >> 64) + p - (uint64_t)(((u128)m * p) >> 64);
        };
        auto out = [p, pinv](uint64_t x) {                         // x / R mod p, canonical (x < 2^64)
            uint64_t m = x * pinv, u = (uint64_t)(((u128)m * p) >> 64);
            return (uint32_t)(u ? p - u : 0);
        };
        auto to_m = [p](uint64_t v) { return (uint64_t)((((u128)(v % p)) << 64) % p); };
        uint64_t x1 = to_m(a1);
        a[0] = (uint32_t)(1 % p); a[1] = (uint32_t)(a1 % p);
        uint64_t sq0 = redc_sq(to_m(1));                           // a_n^2 (Montgomery form, lazily reduced)
        for (size_t i = 2; i < rows; i++) {
            uint64_t sq1 = redc_sq(x1);
            uint64_t x2 = sq1 + sq0;                               // < 2p + 32
            a[i] = out(x2);
            x1 = x2; sq0 = sq1;
        }
        a[rows] = 0;
    }
    // T-th value chosen so that the x^(T-1) coefficient of the size-T interpolant vanishes:
    // sum_{i<T} a_i g^i = 0  =>  the unique degree <= T-2 interpolant through the T-1 rows (Polynomial::interpolate).
    DevBufPtr tr = make_buf(T * 4, ctx->stream);
    STARK_CUDA(cudaMemcpyAsync(tr->p, a, T * 4, cudaMemcpyHostToDevice, ctx->stream));
    uint64_t s = dot_powers(ctx, tr->as<uint32_t>(), rows, log_trace);
    a[rows] = (uint32_t)h_mul((p - s) % p, h_inv(h_pow(g, rows, p), p), p);
    STARK_CUDA(cudaMemcpyAsync(tr->as<uint32_t>() + rows, &a[rows], 4, cudaMemcpyHostToDevice, ctx->stream));
    STARK_CUDA(cudaStreamSynchronize(ctx->stream));
    *last_value = a[rows - 1];
    return tr;
}

static void check_fibsq_sizes(stark_ctx* ctx, unsigned log_trace, unsigned log_blowup) {
    STARK_REQUIRE(log_trace >= 2 && log_trace + log_blowup <= ctx->two_adicity && log_trace + log_blowup <= 30,
                  "stark101: trace/blowup sizes not supported by this field");
}

// Building blocks of the prover for callers that spread the work over several GPUs (multi_gpu.py: stark101_prove_multi):
// the coefficients of the trace polynomial f (T of them, the top one zero) with a_{T-2}, and the composition
// polynomial on a contiguous range of the coset.
extern "C" int stark101_trace_poly(stark_ctx* ctx, uint64_t a1, unsigned log_trace, stark_vec** coeffs, uint64_t* last_value) {
    try {
        STARK_REQUIRE(ctx && coeffs && last_value, "stark101_trace_poly: null argument");
        check_fibsq_sizes(ctx, log_trace, 0);
        std::lock_guard<std::recursive_mutex> lk(ctx->mu);
        STARK_CUDA(cudaSetDevice(ctx->device));
        DevBufPtr tr = fibsq_trace_column(ctx, a1, log_trace, last_value);
        DevBufPtr c = api_interpolate_on_coset(ctx, tr->as<uint32_t>(), log_trace, 1);
        stark_vec* v = new stark_vec(); v->ctx = ctx; v->buf = c; v->n = (size_t)1 << log_trace;
        *coeffs = v;
        return ST_OK;
    } catch (const StarkError& e) { api_set_error(e.what()); return e.code; }
    catch (const std::exception& e) { api_set_error(e.what()); return ST_INTERNAL; }
}
extern "C" int stark101_composition_range(stark_ctx* ctx, const stark_vec* f_block, size_t start, size_t count, const uint64_t alpha[3],
                                          uint64_t last_value, unsigned log_trace, unsigned log_blowup, stark_vec** cp_block) {
    try {
        STARK_REQUIRE(ctx && f_block && alpha && cp_block && f_block->ctx == ctx, "stark101_composition_range: bad argument");
        check_fibsq_sizes(ctx, log_trace, log_blowup);
        const size_t N = (size_t)1 << (log_trace + log_blowup), blow = (size_t)1 << log_blowup;
        STARK_REQUIRE(count >= 1 && start + count <= N, "stark101_composition_range: range outside the domain");
        STARK_REQUIRE(f_block->n == (count == N ? N : count + 2 * blow), "stark101_composition_range: f_block must hold the range plus 2*blowup halo values");
        std::lock_guard<std::recursive_mutex> lk(ctx->mu);
        STARK_CUDA(cudaSetDevice(ctx->device));
        FibSqParams prm{};
        prm.log_trace = log_trace; prm.log_blowup = log_blowup; prm.offset = ctx->generator; prm.last_value = last_value;
        for (int k = 0; k < 3; k++) prm.alpha[k] = alpha[k];
        DevBufPtr out = make_buf(count * 4, ctx->stream);
        fibsq_composition(ctx, f_block->buf->as<uint32_t>(), prm, out->as<uint32_t>(), start, count);
        stark_vec* v = new stark_vec(); v->ctx = ctx; v->buf = out; v->n = count;
        *cp_block = v;
        return ST_OK;
    } catch (const StarkError& e) { api_set_error(e.what()); return e.code; }
    catch (const std::exception& e) { api_set_error(e.what()); return ST_INTERNAL; }
}

extern "C" int stark101_prove(stark_ctx* ctx, uint64_t a1, unsigned log_trace, unsigned log_blowup, size_t num_queries,
                              stark_channel* chan) {
    try {
        STARK_REQUIRE(ctx && chan, "stark101_prove: null argument");
        check_fibsq_sizes(ctx, log_trace, log_blowup);
        std::lock_guard<std::recursive_mutex> lk(ctx->mu);
        STARK_CUDA(cudaSetDevice(ctx->device));
        Channel& ch = chan->ch;
        const unsigned log_N = log_trace + log_blowup;
        const size_t N = (size_t)1 << log_N, blow = (size_t)1 << log_blowup;
        const uint64_t w = ctx->generator;
        uint64_t last_value = 0;
        DevBufPtr tr = fibsq_trace_column(ctx, a1, log_trace, &last_value);
        {   // the statement opens the transcript (host_channel.hpp: stark101_statement)
            uint8_t stmt[STARK101_STATEMENT_BYTES];
            stark101_statement(ctx->modulus, ctx->generator, log_trace, log_blowup, num_queries, last_value, stmt);
            ch.send(stmt, sizeof stmt);
        }
        // ---- LDE of the trace column and its commitment ----
        DevBufPtr f_eval = api_lde_on_coset(ctx, tr->as<uint32_t>(), log_trace, 1, log_blowup, w);
        auto f_tree = api_tree_commit(ctx, f_eval, N);
        api_send_root(ch, f_tree.get());
        FibSqParams prm{};
        prm.log_trace = log_trace; prm.log_blowup = log_blowup; prm.offset = w; prm.last_value = last_value;
        for (int k = 0; k < 3; k++) STARK_REQUIRE(ch.receive_random_field_element(&prm.alpha[k]), "channel: receive before send");
        // ---- src/composition: CP on the coset, then its coefficients ----
        DevBufPtr cp_eval = make_buf(N * 4, ctx->stream);
        fibsq_composition(ctx, f_eval->as<uint32_t>(), prm, cp_eval->as<uint32_t>(), 0, N);
        DevBufPtr cp_coef = api_interpolate_on_coset(ctx, cp_eval->as<uint32_t>(), log_N, w);
        // ---- src/fri: CP's evaluations on the coset ARE layer 0 of fri_commit(CP, domain, channel) ----
        stark_fri* fri = nullptr;
        int rc = api_fri_commit_evaluated(ctx, std::move(cp_coef), N, std::move(cp_eval), log_N, w, chan, &fri);
        if (rc != ST_OK) return rc;
        std::unique_ptr<stark_fri> fri_guard(fri);
        // ---- queries ----
        for (size_t qn = 0; qn < num_queries; qn++) {
            uint64_t idx;
            STARK_REQUIRE(ch.receive_random_int(0, N - 1 - 2 * blow, true, &idx), "channel: receive before send");
            api_open_and_send(f_tree.get(), (size_t)idx, ch);
            api_open_and_send(f_tree.get(), (size_t)idx + blow, ch);
            api_open_and_send(f_tree.get(), (size_t)idx + 2 * blow, ch);
            rc = stark_decommit_fri_layers(fri, (size_t)idx, chan);
            if (rc != ST_OK) return rc;
        }
        return ST_OK;
    } catch (const StarkError& e) {
        api_set_error(e.what());
        return e.code;
    } catch (const std::exception& e) {
        api_set_error(e.what());
        return ST_INTERNAL;
    }
}
