// fourstep.cu — the exchange steps of the multi-GPU four-step NTT, written straight into PEER memory
// (NVLink 5 / NVSwitch, P2P stores through IPC-mapped pointers) instead of pack -> NCCL all-to-all -> unpack.
//
//   N = N1*N2, n = n1*N2 + n2, k = k1 + N1*k2, G ranks, w = N2/G
//   phase A  (rank r)  A[n1][n2'] (n2 = r*w + n2'): column-batched DIF over n1  -> slot q holds k1 = bitrev(q)
//   exch. 1            row k1 of A times w_N^(n2*k1) -> owner s = k1/(N1/G), row k1 % (N1/G), columns r*w .. r*w+w-1
//                      (every row is one contiguous 4*w-byte store into the peer: fused twiddle + exchange)
//   phase C  (rank s)  B[k1'][n2]: row-batched DIF over n2 -> slot q2 holds k2 = bitrev(q2)
//   exch. 2            32x32 shared-memory transpose, then 128-byte stores into the peer that owns k2:
//                      natural-order block of rank t: position (k2 % (N2/G))*N1 + k1      (fused transpose + exchange)
// No bit-reversal pass and no separate pack/unpack pass exists anywhere in this pipeline.
#include <stdlib.h>

#include <algorithm>

#include "../../include/stark_b200.h"
#include "handles.hpp"

namespace starkb200 {

__device__ __forceinline__ uint32_t brev_bits(uint32_t x, unsigned bits) { return bits ? (__brev(x) >> (32 - bits)) : 0u; }
// output index held by slot q (kernels.hpp: SlotMap)
__device__ __forceinline__ uint32_t slot_index(const SlotMap& m, uint32_t q) {
    if (m.mode == 0) return brev_bits(q, m.total_bits);
    if (m.mode == 2) return q;
    uint32_t k = 0, off = 0, pos = m.total_bits;
#pragma unroll
    for (int d = 0; d < 4; d++) {            // digit d sits at address bits [pos - bits[d], pos) and is worth 2^off
        if (d < (int)m.nd) {
            pos -= m.bits[d];
            k |= ((q >> pos) & ((1u << m.bits[d]) - 1u)) << off;
            off += m.bits[d];
        }
    }
    return k;
}

__global__ void fourstep_stage_kernel(const uint32_t* __restrict__ c, size_t len, uint32_t* __restrict__ A, unsigned log_n1,
                                      unsigned log_n2, unsigned log_w, unsigned rank, int has_scale, PowTable scale, FieldParams fp) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >> (log_n1 + log_w)) return;
    size_t n1 = i >> log_w, n2 = ((size_t)rank << log_w) + (i & (((size_t)1 << log_w) - 1));
    size_t j = (n1 << log_n2) + n2;
    uint32_t v = j < len ? c[j] : 0u;
    if (has_scale && v) v = mont_mul(v, pow_lookup(scale, (uint32_t)j, fp), fp);
    A[i] = v;
}
void fourstep_stage_input(stark_ctx* ctx, const uint32_t* coeffs, size_t len, uint32_t* A, unsigned log_n1, unsigned log_n2,
                          unsigned world, unsigned rank, uint64_t offset) {
    unsigned log_g = 0; while ((1u << log_g) < world) log_g++;
    unsigned log_w = log_n2 - log_g;
    size_t total = (size_t)1 << (log_n1 + log_w);
    bool unit = offset % ctx->modulus == 1;
    ScaleTable st;
    if (!unit) build_scale_table(ctx, offset % ctx->modulus, 1, log_n1 + log_n2, st);
    KernelTimer kt(ctx, stark_ctx::CAT_OTHER, 8.0 * total);
    fourstep_stage_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(coeffs, len, A, log_n1, log_n2, log_w, rank, !unit,
                                                                                   st.view, ctx->fp);
    ctx->launches++;
    STARK_CUDA(cudaGetLastError());
}

// ---- device-side hand-over between the phases (no host barrier, no stream synchronisation) ---------------------
// Every rank owns a small peer-visible array of epoch words: flags[slot * MAX_PEERS + r] = the last transform for which
// rank r has finished storing into THIS rank's buffer (slot 0: the rows of exchange 1, slot 1: the block of exchange 2).
// A one-warp kernel behind each scatter kernel publishes the epoch to all peers (st.release.sys); the consumer runs a
// one-warp kernel in front of its next phase that spins with ld.acquire.sys until all `world` words have reached the
// epoch; stream order does the rest.
struct FlagPtrs { uint32_t* p[MAX_PEERS]; };
__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Hand-over: a ONE-WARP kernel enqueued right behind a scatter kernel publishes the epoch to every peer.  Stream order
// makes it start only when the scatter kernel's grid has completed, i.e. when all of its stores -- the peer stores over
// NVLink included -- have been performed; the fence + st.release.sys pair then orders the flag behind them for the
// peer's ld.acquire.sys.  (Measured alternatives, profiles/r02_fourstep.md: a system-scope fence in every thread of the
// scatter kernel: 91 / 113 us instead of 23 / 31 us for 2^23 elements; one fence per CTA + "last CTA publishes": still
// +23 us per kernel, with 8192 CTAs or with 1184 persistent ones -- a fence.sys stalls behind every outstanding peer
// store of its SM.  The extra launch costs ~3 us.)
__global__ void fourstep_publish_kernel(FlagPtrs flags, unsigned slot, unsigned rank, unsigned world, uint32_t epoch) {
    const unsigned s = threadIdx.x;
    if (s >= world) return;
    __threadfence_system();
    st_release_sys_u32(flags.p[s] + slot * MAX_PEERS + rank, epoch);
}
// Spins until every peer's epoch word of `slot` has reached `epoch` (wrap-safe compare).  After `timeout_ns` it gives up
// and raises result->flag so that the host reports a lost peer instead of hanging.
__global__ void fourstep_wait_kernel(const uint32_t* flags, unsigned slot, unsigned world, uint32_t epoch, HostResult* result,
                                     unsigned long long timeout_ns) {
    const unsigned r = threadIdx.x;
    if (r >= world) return;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        const uint32_t v = ld_acquire_sys_u32(flags + slot * MAX_PEERS + r);
        if ((int32_t)(v - epoch) >= 0) break;
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > timeout_ns) { if (result) result->flag = 0xdead0000u | (slot << 8) | r; break; }
        __nanosleep(200);
    }
}

// The peers' base pointers travel as a kernel-parameter array; indexing that array with a run-time owner makes the
// compiler copy all 16 pointers to every thread's local memory (a 128-byte stack frame: 268 MB of local stores for 2^21
// threads -- 75-82 us instead of ~20 for these kernels, profiles/r02_fourstep.md).  Staged once per CTA in shared memory
// with compile-time indices instead.
__device__ __forceinline__ void stage_peers(uint32_t** s_peer, const PeerPtrs& peers, unsigned tid) {
    if (tid != 0) return;
#pragma unroll
    for (int i = 0; i < MAX_PEERS; i++) s_peer[i] = peers.p[i];
}

// One CTA per (row, segment of up to 1024 columns); four consecutive columns per thread: one 16-byte load, one 16-byte
// store into the owner.  The twiddle of column c of row k1 is w_N^(k1 (n2_0 + c)): with c = 32 h + j it factors into
// Rhi[h] = w_N^(k1 (n2_0 + 32 h)) and Rlo[j] = w_N^(k1 j) -- 64 table look-ups per CTA, kept in shared memory, instead
// of two scattered global look-ups per ELEMENT (whose 32 different L1 lines per warp instruction bound the round-1 kernel).
// dst = base[owner] + row * row_pitch + col_off + c:
// peer memory: row_pitch = N2, col_off = rank * w (the owner's [N1/G][N2] matrix);  staging for NCCL: row_pitch = w, col_off = 0.
__global__ void fourstep_rows_kernel(const uint32_t* __restrict__ A, unsigned log_n1, unsigned log_w, unsigned log_g,
                                     unsigned rank, PeerPtrs peers, size_t row_pitch, size_t col_off, PowTable tw, FieldParams fp, SlotMap map) {
    __shared__ uint32_t* s_peer[MAX_PEERS];
    __shared__ uint32_t s_lo[32], s_hi[32];
    const unsigned tid = threadIdx.x, seg_len = blockDim.x * 4;          // columns of this CTA's segment
    const uint32_t q = blockIdx.y, c0 = blockIdx.x * seg_len;            // slot (row of A), first column of the segment
    const uint32_t k1 = slot_index(map, q);
    const uint32_t n2_0 = (rank << log_w) + c0;
    stage_peers(s_peer, peers, tid);
    for (unsigned i = tid; i < 64; i += blockDim.x) {                     // (a segment may have fewer than 64 threads)
        if (i < 32) s_lo[i] = pow_lookup(tw, k1 * i, fp);                                   // k1 * 31 < N
        else if ((i - 32) * 32 < seg_len) s_hi[i - 32] = pow_lookup(tw, k1 * (n2_0 + (i - 32) * 32), fp);   // k1 * n2 < N
    }
    const uint32_t c = c0 + tid * 4;
    const uint4 a = *reinterpret_cast<const uint4*>(A + ((size_t)q << log_w) + c);          // in flight across the barrier
    __syncthreads();
    const uint32_t hi = s_hi[tid >> 3], j = (tid & 7) * 4;
    uint4 v;
    v.x = mont_mul(a.x, mont_mul(s_lo[j], hi, fp), fp);
    v.y = mont_mul(a.y, mont_mul(s_lo[j + 1], hi, fp), fp);
    v.z = mont_mul(a.z, mont_mul(s_lo[j + 2], hi, fp), fp);
    v.w = mont_mul(a.w, mont_mul(s_lo[j + 3], hi, fp), fp);
    const unsigned rows_per = log_n1 - log_g;
    const uint32_t owner = k1 >> rows_per, row = k1 & ((1u << rows_per) - 1);
    *reinterpret_cast<uint4*>(s_peer[owner] + (size_t)row * row_pitch + col_off + c) = v;
}
static void fs_publish(stark_ctx* ctx, void* const* peer_flags, unsigned slot, unsigned rank, unsigned world, uint32_t epoch) {
    if (!peer_flags) return;
    FlagPtrs f{};
    for (unsigned s = 0; s < world; s++) f.p[s] = static_cast<uint32_t*>(peer_flags[s]);
    fourstep_publish_kernel<<<1, 32, 0, ctx->stream>>>(f, slot, rank, world, epoch);
    ctx->launches++;
}
void fourstep_twiddle_scatter_rows(stark_ctx* ctx, const uint32_t* A, unsigned log_n1, unsigned log_n2, unsigned world, unsigned rank,
                                   const PeerPtrs& peers, bool staged, void* const* peer_flags, uint32_t epoch, const SlotMap& map) {
    unsigned log_g = 0; while ((1u << log_g) < world) log_g++;
    unsigned log_w = log_n2 - log_g;
    STARK_REQUIRE(log_w >= 2, "fourstep: every rank needs >= 4 columns");
    size_t total = (size_t)1 << (log_n1 + log_w);
    const TwiddleSet& tws = ctx->twiddles(log_n1 + log_n2);
    const size_t w = (size_t)1 << log_w;
    STARK_REQUIRE(log_w >= 5, "fourstep: every rank needs >= 32 columns");
    KernelTimer kt(ctx, stark_ctx::CAT_OTHER, 16.0 * total);
    const unsigned seg = w < 1024 ? (unsigned)w : 1024u;                       // columns per CTA, 4 per thread
    STARK_REQUIRE(((size_t)1 << log_n1) <= 65535, "fourstep: too many rows for one grid dimension");
    fourstep_rows_kernel<<<dim3((unsigned)(w / seg), 1u << log_n1), seg / 4, 0, ctx->stream>>>(
        A, log_n1, log_w, log_g, rank, peers, staged ? w : ((size_t)1 << log_n2), staged ? 0 : (size_t)rank * w, tws.fwd(), ctx->fp, map);
    ctx->launches++;
    fs_publish(ctx, peer_flags, 0, rank, world, epoch);
    STARK_CUDA(cudaGetLastError());
}

// tile: 32 rows (k1') x 32 slots (q2); block (32, 8).  dst = base[owner] + col * col_pitch + k1_off + k1':
// peer memory: col_pitch = N1, k1_off = rank * N1/G (the owner's natural-order block);  staging: col_pitch = N1/G, k1_off = 0.
__global__ void fourstep_transpose_kernel(const uint32_t* __restrict__ X, unsigned log_n2, unsigned log_g, PeerPtrs peers,
                                          size_t col_pitch, size_t k1_off, SlotMap map) {
    __shared__ uint32_t tile[32][33];
    __shared__ uint32_t* s_peer[MAX_PEERS];
    stage_peers(s_peer, peers, threadIdx.y * 32 + threadIdx.x);
    const size_t n2 = (size_t)1 << log_n2;
    const uint32_t q0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
#pragma unroll
    for (int j = threadIdx.y; j < 32; j += 8) tile[j][threadIdx.x] = X[(size_t)(r0 + j) * n2 + q0 + threadIdx.x];
    __syncthreads();
    const unsigned cols_per = log_n2 - log_g;
#pragma unroll
    for (int j = threadIdx.y; j < 32; j += 8) {
        uint32_t k2 = slot_index(map, q0 + j);
        uint32_t owner = k2 >> cols_per, col = k2 & ((1u << cols_per) - 1);
        s_peer[owner][(size_t)col * col_pitch + k1_off + r0 + threadIdx.x] = tile[threadIdx.x][j];
    }
}
void fourstep_transpose_scatter(stark_ctx* ctx, const uint32_t* X, unsigned log_n1, unsigned log_n2, unsigned world, unsigned rank,
                                const PeerPtrs& peers, bool staged, void* const* peer_flags, uint32_t epoch, const SlotMap& map) {
    unsigned log_g = 0; while ((1u << log_g) < world) log_g++;
    unsigned log_r = log_n1 - log_g;
    STARK_REQUIRE(log_r >= 5 && log_n2 >= 5, "fourstep: tiles need >= 32 rows and columns per rank");
    dim3 grid(1u << (log_n2 - 5), 1u << (log_r - 5));
    KernelTimer kt(ctx, stark_ctx::CAT_OTHER, 16.0 * (double)((size_t)1 << (log_r + log_n2)));
    fourstep_transpose_kernel<<<grid, dim3(32, 8), 0, ctx->stream>>>(
        X, log_n2, log_g, peers, staged ? ((size_t)1 << log_r) : ((size_t)1 << log_n1), staged ? 0 : ((size_t)rank << log_r), map);
    ctx->launches++;
    fs_publish(ctx, peer_flags, 1, rank, world, epoch);
    STARK_CUDA(cudaGetLastError());
}
void fourstep_wait(stark_ctx* ctx, const void* own_flags, unsigned slot, unsigned world, uint32_t epoch) {
    KernelTimer kt(ctx, stark_ctx::CAT_OTHER, 0);
    fourstep_wait_kernel<<<1, 32, 0, ctx->stream>>>(static_cast<const uint32_t*>(own_flags), slot, world, epoch, ctx->d_result, 5000000000ull);
    ctx->launches++;
    STARK_CUDA(cudaGetLastError());
}

// the two compute phases, shared by the C entry points below and by the NCCL-staged transport of multi.cu
static SlotMap bitrev_map(unsigned bits) { SlotMap m{}; m.mode = 0; m.total_bits = bits; return m; }
void fourstep_phase_a_launch(stark_ctx* ctx, const uint32_t* coeffs, size_t n_coeffs, unsigned log_n, uint64_t offset, unsigned world,
                             unsigned rank, const PeerPtrs& dst, bool staged, void* const* peer_flags, uint32_t epoch) {
    unsigned a = log_n / 2, b = log_n - a, log_g = 0;
    while ((1u << log_g) < world) log_g++;
    const unsigned log_w = b - log_g;
    DevBuf A(((size_t)4) << (a + log_w), ctx->stream);
    SlotMap map = bitrev_map(a);
    if (a >= 10 && a + log_w <= 31 && (reinterpret_cast<uintptr_t>(coeffs) & 15) == 0) {
        // the strided passes of the natural-order transform (lazy 9-instruction butterflies, 16-byte accesses, one
        // row twiddle per tile row); the first pass gathers this rank's columns of the coefficient matrix and applies the
        // coset scale on the way in (no staging sweep); they leave digit-reversed slots, which the scatter below undoes for free
        const bool unit = offset % ctx->modulus == 1;
        ScaleTable st;
        if (!unit) build_scale_table(ctx, offset % ctx->modulus, 1, a + b, st);
        ColumnSource from{coeffs, n_coeffs, b, (unsigned)(rank << log_w), unit ? nullptr : &st.view};
        std::vector<unsigned> bits = ntt_columns_digitrev(ctx, A.as<uint32_t>(), a, log_w, false, &from);
        map.mode = 1; map.nd = (unsigned)bits.size();
        for (size_t i = 0; i < bits.size(); i++) map.bits[i] = bits[i];
    } else {
        fourstep_stage_input(ctx, coeffs, n_coeffs, A.as<uint32_t>(), a, b, world, rank, offset);
        ntt_dif_columns(ctx, A.as<uint32_t>(), a, log_w, false);
    }
    fourstep_twiddle_scatter_rows(ctx, A.as<uint32_t>(), a, b, world, rank, dst, staged, peer_flags, epoch, map);
}
void fourstep_phase_c_launch(stark_ctx* ctx, uint32_t* rows, unsigned log_n, unsigned world, unsigned rank, const PeerPtrs& dst,
                             bool staged, void* const* peer_flags, uint32_t epoch) {
    unsigned a = log_n / 2, b = log_n - a, log_g = 0;
    while ((1u << log_g) < world) log_g++;
    const size_t batch = (size_t)1 << (a - log_g), total = batch << b;
    SlotMap map = bitrev_map(b);
    if (ntt_natural_supported(b, rows, rows, rows) && total <= ((size_t)1 << 31)) {
        // natural -> natural row transforms (strided passes into a scratch array, transposing last pass back into `rows`)
        DevBuf work(total * 4, ctx->stream);
        ntt_natural(ctx, rows, total, work.as<uint32_t>(), rows, b, false, nullptr, nullptr, batch);
        map.mode = 2;
    } else {
        ntt_dif(ctx, rows, b, false, batch);
    }
    fourstep_transpose_scatter(ctx, rows, a, b, world, rank, dst, staged, peer_flags, epoch, map);
}

}  // namespace starkb200

using namespace starkb200;
namespace starkb200 { void api_set_error(const std::string& s); }

#define FS_BEGIN try {
#define FS_END                                                                              \
    }                                                                                       \
    catch (const StarkError& e) { api_set_error(e.what()); return e.code; }                 \
    catch (const std::exception& e) { api_set_error(e.what()); return ST_INTERNAL; }        \
    return ST_OK;

struct FsGuard {
    std::lock_guard<std::recursive_mutex> lk;
    explicit FsGuard(stark_ctx* c) : lk(c->mu) { STARK_CUDA(cudaSetDevice(c->device)); }
};

// ---- peer-visible buffers (cudaMalloc + CUDA IPC; the stream-ordered pool cannot be exported) ----
extern "C" int stark_peer_alloc(stark_ctx* ctx, size_t n, stark_vec** out, uint8_t handle[64]) {
    FS_BEGIN
    STARK_REQUIRE(ctx && out && handle && n >= 1, "peer_alloc: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    FsGuard g(ctx);
    auto buf = std::make_shared<DevBuf>();
    STARK_CUDA(cudaMalloc(&buf->p, n * 4));
    buf->bytes = n * 4; buf->stream = ctx->stream; buf->plain = true;
    STARK_CUDA(cudaMemsetAsync(buf->p, 0, n * 4, ctx->stream));
    cudaIpcMemHandle_t h;
    STARK_CUDA(cudaIpcGetMemHandle(&h, buf->p));
    memcpy(handle, &h, 64);
    stark_vec* v = new stark_vec(); v->ctx = ctx; v->buf = buf; v->n = n;
    *out = v;
    FS_END
}
extern "C" int stark_peer_open(stark_ctx* ctx, const uint8_t handle[64], void** dptr) {
    FS_BEGIN
    STARK_REQUIRE(ctx && handle && dptr, "peer_open: bad argument");
    FsGuard g(ctx);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    STARK_CUDA(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
    FS_END
}
extern "C" int stark_peer_close(stark_ctx* ctx, void* dptr) {
    FS_BEGIN
    STARK_REQUIRE(ctx && dptr, "peer_close: bad argument");
    FsGuard g(ctx);
    STARK_CUDA(cudaStreamSynchronize(ctx->stream));
    STARK_CUDA(cudaIpcCloseMemHandle(dptr));
    FS_END
}

static PeerPtrs make_peers(void* const* ptrs, unsigned world) {
    STARK_REQUIRE(world >= 1 && world <= (unsigned)MAX_PEERS && (world & (world - 1)) == 0, "fourstep: world must be a power of two <= 16");
    PeerPtrs p{};
    for (unsigned i = 0; i < world; i++) { STARK_REQUIRE(ptrs[i] != nullptr, "fourstep: null peer pointer"); p.p[i] = static_cast<uint32_t*>(ptrs[i]); }
    return p;
}
static void fs_dims(stark_ctx* ctx, unsigned log_n, unsigned world, unsigned& a, unsigned& b, unsigned& log_g) {
    a = log_n / 2; b = log_n - a;
    log_g = 0; while ((1u << log_g) < world) log_g++;
    STARK_REQUIRE(log_n <= ctx->two_adicity && log_n <= 30, "fourstep: 2^log_n does not divide p-1");
    STARK_REQUIRE(a >= log_g + 5 && b >= log_g + 5, "fourstep: every rank needs >= 32 rows and >= 32 columns (raise log_n or lower the world size)");
}

// Phase A + exchange 1: coefficients (device, every rank holds them) -> rows of the [N1/G][N2] matrix of each peer.
// peer_flags == NULL: returns after the stores are complete (the caller places a barrier before phase C).
// peer_flags != NULL: nothing is synchronised -- the scatter kernel publishes `epoch` into slot 0 of every peer's flags.
extern "C" int stark_fourstep_phase_a(stark_ctx* ctx, const stark_vec* coeffs, unsigned log_n, uint64_t offset, unsigned world,
                                      unsigned rank, void* const* peer_rows, void* const* peer_flags, uint32_t epoch) {
    FS_BEGIN
    STARK_REQUIRE(ctx && coeffs && coeffs->ctx == ctx && peer_rows && rank < world, "fourstep_phase_a: bad argument");
    FsGuard g(ctx);
    STARK_REQUIRE(offset % ctx->modulus != 0, "coset offset must be non-zero");
    unsigned a, b, log_g; fs_dims(ctx, log_n, world, a, b, log_g);
    STARK_REQUIRE(coeffs->n <= ((size_t)1 << log_n), "fourstep: more coefficients than domain points");
    PeerPtrs peers = make_peers(peer_rows, world);
    if (peer_flags) make_peers(peer_flags, world);
    fourstep_phase_a_launch(ctx, coeffs->buf->as<uint32_t>(), coeffs->n, log_n, offset, world, rank, peers, false, peer_flags, epoch);
    if (!peer_flags) STARK_CUDA(cudaStreamSynchronize(ctx->stream));      // the stores into the peers are complete when this returns
    FS_END
}
// Phase C + exchange 2: this rank's [N1/G][N2] rows (all peers have written them) -> natural-order blocks of each peer.
// With flags: first waits (on the device) until every peer's phase-A epoch has arrived, and publishes `epoch` into slot 1
// of every peer's flags when its own stores are done.
extern "C" int stark_fourstep_phase_c(stark_ctx* ctx, stark_vec* rows, unsigned log_n, unsigned world, unsigned rank,
                                      void* const* peer_blocks, void* const* peer_flags, uint32_t epoch) {
    FS_BEGIN
    STARK_REQUIRE(ctx && rows && rows->ctx == ctx && peer_blocks && rank < world, "fourstep_phase_c: bad argument");
    FsGuard g(ctx);
    unsigned a, b, log_g; fs_dims(ctx, log_n, world, a, b, log_g);
    STARK_REQUIRE(rows->n == ((size_t)1 << (log_n - log_g)), "fourstep_phase_c: rows vector has the wrong size");
    PeerPtrs peers = make_peers(peer_blocks, world);
    if (peer_flags) { make_peers(peer_flags, world); fourstep_wait(ctx, peer_flags[rank], 0, world, epoch); }
    fourstep_phase_c_launch(ctx, rows->buf->as<uint32_t>(), log_n, world, rank, peers, false, peer_flags, epoch);
    if (!peer_flags) STARK_CUDA(cudaStreamSynchronize(ctx->stream));
    FS_END
}
// Enqueues the device-side wait for slot (0: rows, 1: block) of this rank's own flag array.
extern "C" int stark_fourstep_wait(stark_ctx* ctx, const void* own_flags, unsigned slot, unsigned world, uint32_t epoch) {
    FS_BEGIN
    STARK_REQUIRE(ctx && own_flags && slot < 2 && world >= 1 && world <= (unsigned)MAX_PEERS, "fourstep_wait: bad argument");
    FsGuard g(ctx);
    fourstep_wait(ctx, own_flags, slot, world, epoch);
    FS_END
}
