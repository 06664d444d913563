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
#include "../../include/stark_b200.h"
#include "handles.hpp"

namespace starkb200 {

__device__ __forceinline__ uint32_t brev_bits(uint32_t x, unsigned bits) { return bits ? (__brev(x) >> (32 - bits)) : 0u; }

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

__global__ void fourstep_rows_kernel(const uint32_t* __restrict__ A, unsigned log_n1, unsigned log_n2, unsigned log_w, unsigned log_g,
                                     unsigned rank, PeerPtrs peers, PowTable tw, FieldParams fp) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >> (log_n1 + log_w)) return;
    uint32_t q = (uint32_t)(i >> log_w), c = (uint32_t)(i & (((size_t)1 << log_w) - 1));
    uint32_t k1 = brev_bits(q, log_n1);
    uint32_t n2 = (rank << log_w) + c;
    uint32_t v = mont_mul(A[i], pow_lookup(tw, k1 * n2, fp), fp);            // k1*n2 < N
    unsigned rows_per = log_n1 - log_g;
    uint32_t owner = k1 >> rows_per, row = k1 & ((1u << rows_per) - 1);
    peers.p[owner][((size_t)row << log_n2) + n2] = v;
}
void fourstep_twiddle_scatter_rows(stark_ctx* ctx, const uint32_t* A, unsigned log_n1, unsigned log_n2, unsigned world, unsigned rank,
                                   const PeerPtrs& peers) {
    unsigned log_g = 0; while ((1u << log_g) < world) log_g++;
    unsigned log_w = log_n2 - log_g;
    size_t total = (size_t)1 << (log_n1 + log_w);
    const TwiddleSet& tws = ctx->twiddles(log_n1 + log_n2);
    KernelTimer kt(ctx, stark_ctx::CAT_OTHER, 16.0 * total);
    fourstep_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(A, log_n1, log_n2, log_w, log_g, rank, peers, tws.fwd(), ctx->fp);
    ctx->launches++;
    STARK_CUDA(cudaGetLastError());
}

// tile: 32 rows (k1') x 32 slots (q2); block (32, 8)
__global__ void fourstep_transpose_kernel(const uint32_t* __restrict__ X, unsigned log_r, unsigned log_n1, unsigned log_n2, unsigned log_g,
                                          unsigned rank, PeerPtrs peers) {
    __shared__ uint32_t tile[32][33];
    const size_t n2 = (size_t)1 << log_n2;
    const uint32_t q0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
#pragma unroll
    for (int j = threadIdx.y; j < 32; j += 8) tile[j][threadIdx.x] = X[(size_t)(r0 + j) * n2 + q0 + threadIdx.x];
    __syncthreads();
    const unsigned cols_per = log_n2 - log_g;
#pragma unroll
    for (int j = threadIdx.y; j < 32; j += 8) {
        uint32_t k2 = brev_bits(q0 + j, log_n2);
        uint32_t owner = k2 >> cols_per, col = k2 & ((1u << cols_per) - 1);
        size_t k1 = ((size_t)rank << log_r) + r0 + threadIdx.x;
        peers.p[owner][((size_t)col << log_n1) + k1] = tile[threadIdx.x][j];
    }
}
void fourstep_transpose_scatter(stark_ctx* ctx, const uint32_t* X, unsigned log_n1, unsigned log_n2, unsigned world, unsigned rank,
                                const PeerPtrs& peers) {
    unsigned log_g = 0; while ((1u << log_g) < world) log_g++;
    unsigned log_r = log_n1 - log_g;
    STARK_REQUIRE(log_r >= 5 && log_n2 >= 5, "fourstep: tiles need >= 32 rows and columns per rank");
    dim3 grid(1u << (log_n2 - 5), 1u << (log_r - 5));
    KernelTimer kt(ctx, stark_ctx::CAT_OTHER, 16.0 * (double)((size_t)1 << (log_r + log_n2)));
    fourstep_transpose_kernel<<<grid, dim3(32, 8), 0, ctx->stream>>>(X, log_r, log_n1, log_n2, log_g, rank, peers);
    ctx->launches++;
    STARK_CUDA(cudaGetLastError());
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
extern "C" int stark_fourstep_phase_a(stark_ctx* ctx, const stark_vec* coeffs, unsigned log_n, uint64_t offset, unsigned world,
                                      unsigned rank, void* const* peer_rows) {
    FS_BEGIN
    STARK_REQUIRE(ctx && coeffs && coeffs->ctx == ctx && peer_rows && rank < world, "fourstep_phase_a: bad argument");
    FsGuard g(ctx);
    STARK_REQUIRE(offset % ctx->modulus != 0, "coset offset must be non-zero");
    unsigned a, b, log_g; fs_dims(ctx, log_n, world, a, b, log_g);
    STARK_REQUIRE(coeffs->n <= ((size_t)1 << log_n), "fourstep: more coefficients than domain points");
    PeerPtrs peers = make_peers(peer_rows, world);
    unsigned log_w = b - log_g;
    DevBuf A(((size_t)4) << (a + log_w), ctx->stream);
    fourstep_stage_input(ctx, coeffs->buf->as<uint32_t>(), coeffs->n, A.as<uint32_t>(), a, b, world, rank, offset);
    ntt_dif_columns(ctx, A.as<uint32_t>(), a, log_w, false);
    fourstep_twiddle_scatter_rows(ctx, A.as<uint32_t>(), a, b, world, rank, peers);
    STARK_CUDA(cudaStreamSynchronize(ctx->stream));      // the stores into the peers are complete when this returns
    FS_END
}
// Phase C + exchange 2: this rank's [N1/G][N2] rows (all peers have written them) -> natural-order blocks of each peer.
extern "C" int stark_fourstep_phase_c(stark_ctx* ctx, stark_vec* rows, unsigned log_n, unsigned world, unsigned rank,
                                      void* const* peer_blocks) {
    FS_BEGIN
    STARK_REQUIRE(ctx && rows && rows->ctx == ctx && peer_blocks && rank < world, "fourstep_phase_c: bad argument");
    FsGuard g(ctx);
    unsigned a, b, log_g; fs_dims(ctx, log_n, world, a, b, log_g);
    STARK_REQUIRE(rows->n == ((size_t)1 << (log_n - log_g)), "fourstep_phase_c: rows vector has the wrong size");
    PeerPtrs peers = make_peers(peer_blocks, world);
    ntt_dif(ctx, rows->buf->as<uint32_t>(), b, false, (size_t)1 << (a - log_g));
    fourstep_transpose_scatter(ctx, rows->buf->as<uint32_t>(), a, b, world, rank, peers);
    STARK_CUDA(cudaStreamSynchronize(ctx->stream));
    FS_END
}
