// multi.cu — the multi-GPU entry points of the C ABI: one process per GPU, one `stark_mg` per process.
//
// What shards, and nothing else does (north-star):
//   * independent trace columns: column c -> rank c mod G, LDE + Merkle tree per column on the owner, the 32-byte
//     roots all-gathered (stark_mg_commit_columns, BASELINE cfg4);
//   * one large column: the four-step NTT with an all-to-all transpose between its two phases -- over NCCL
//     (ncclSend/ncclRecv groups) or stored straight into peer memory over NVLink by the kernels that produce the data,
//     handed over with device-side epoch flags (fourstep.cu) -- then contiguous leaf ranges hashed per rank and the
//     subtree roots gathered (stark_mg_fourstep_lde, stark_mg_commit_leaf_ranges);
//   * fri_commit / decommit_fri with layer 0 distributed that way, and the leaf hashing of the later LARGE layers too:
//     the folds are replicated on every rank, each rank hashes its leaf range of the new layer, subtree roots are
//     gathered (sharded_fri_layers); the smaller trees and the channel stay on rank 0 (stark_mg_fri_commit,
//     stark_mg_decommit_fri, BASELINE cfg5).
//
// NCCL is resolved at run time (dlopen of libnccl.so.2: the copy a host framework has already loaded, or the system
// one), so the library has no link-time dependency on it and single-GPU users never touch it.
#include <dlfcn.h>
#include <string.h>

#include <algorithm>
#include <array>
#include <deque>

#include "../../include/stark_b200.h"
#include "handles.hpp"

using namespace starkb200;

namespace starkb200 {
void api_set_error(const std::string& s);
DevBufPtr api_upload_u64(stark_ctx* ctx, const uint64_t* host, size_t n);
DevBufPtr api_lde_on_coset(stark_ctx* ctx, const uint32_t* evals, unsigned log_n, uint64_t offset_in, unsigned log_blowup, uint64_t offset_out);
std::unique_ptr<stark_tree> api_tree_launch_values(stark_ctx* ctx, DevBufPtr leaves, size_t n, HostResult* result);
int api_fri_commit_loop(stark_fri* f, stark_channel* chan);
int api_fri_commit_loop_resume(stark_fri* f, stark_channel* chan);
DevBufPtr api_fri_fold_values(stark_fri* f, uint64_t beta);
void api_fri_adopt_layer(stark_fri* f, DevBufPtr layer, const uint8_t root[32]);
void api_open_records(stark_ctx* ctx, const std::vector<OpenDesc>& descs, size_t total_bytes, uint8_t* host_out);
void api_send_query_records(const stark_fri* f, const uint8_t* rec, Channel& ch, size_t index, size_t first_layer);
DevBufPtr api_interpolate_on_coset(stark_ctx* ctx, const uint32_t* evals, unsigned log_n, uint64_t offset);
void api_send_root_bytes(Channel& ch, const uint8_t root[32]);
}  // namespace starkb200

namespace {

// ---- the NCCL surface this file uses, bound at run time -------------------------------------------------------
typedef struct ncclComm* nccl_comm_t;
struct nccl_unique_id { char internal[128]; };
enum { NCCL_UINT8 = 1, NCCL_UINT32 = 3 };
struct Nccl {
    void* lib = nullptr;
    int (*GetUniqueId)(nccl_unique_id*) = nullptr;
    int (*CommInitRank)(nccl_comm_t*, int, nccl_unique_id, int) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
Nccl& nccl() {
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            n.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (n.lib) break;
        }
        if (!n.lib) return;
        auto sym = [&](const char* s) { return dlsym(n.lib, s); };
        n.GetUniqueId = (decltype(n.GetUniqueId))sym("ncclGetUniqueId");
        n.CommInitRank = (decltype(n.CommInitRank))sym("ncclCommInitRank");
        n.CommDestroy = (decltype(n.CommDestroy))sym("ncclCommDestroy");
        n.AllGather = (decltype(n.AllGather))sym("ncclAllGather");
        n.Broadcast = (decltype(n.Broadcast))sym("ncclBroadcast");
        n.Send = (decltype(n.Send))sym("ncclSend");
        n.Recv = (decltype(n.Recv))sym("ncclRecv");
        n.GroupStart = (decltype(n.GroupStart))sym("ncclGroupStart");
        n.GroupEnd = (decltype(n.GroupEnd))sym("ncclGroupEnd");
        n.GetErrorString = (decltype(n.GetErrorString))sym("ncclGetErrorString");
    });
    if (!n.lib || !n.GetUniqueId || !n.CommInitRank || !n.CommDestroy || !n.AllGather || !n.Broadcast || !n.Send || !n.Recv ||
        !n.GroupStart || !n.GroupEnd)
        throw StarkError(ST_UNSUPPORTED, "NCCL (libnccl.so.2) could not be loaded: the multi-GPU entry points need it");
    return n;
}
#define STARK_NCCL(expr)                                                                              \
    do {                                                                                              \
        int _r = (expr);                                                                              \
        if (_r != 0) {                                                                                \
            Nccl& _n = nccl();                                                                        \
            throw StarkError(ST_CUDA, std::string(#expr) + ": " + (_n.GetErrorString ? _n.GetErrorString(_r) : "NCCL error")); \
        }                                                                                             \
    } while (0)

unsigned ilog2(size_t n) { unsigned l = 0; while (((size_t)1 << l) < n) l++; return l; }

void words_to_bytes(const uint32_t w[8], uint8_t out[32]) {
    for (int i = 0; i < 8; i++) { out[4 * i] = (uint8_t)(w[i] >> 24); out[4 * i + 1] = (uint8_t)(w[i] >> 16); out[4 * i + 2] = (uint8_t)(w[i] >> 8); out[4 * i + 3] = (uint8_t)w[i]; }
}
void node_digest(const uint8_t* l, const uint8_t* r, uint8_t out[32]) {
    uint8_t cat[64];
    memcpy(cat, l, 32); memcpy(cat + 32, r, 32);
    HostSha256::digest(cat, 64, out);
}
// the levels above G subtree roots (rs_merkle rule; G is a power of two here, so no promotion happens)
void combine_roots(const uint8_t* roots, unsigned g, uint8_t out[32]) {
    std::vector<std::array<uint8_t, 32>> level(g);
    for (unsigned i = 0; i < g; i++) memcpy(level[i].data(), roots + 32 * i, 32);
    while (level.size() > 1) {
        std::vector<std::array<uint8_t, 32>> nxt((level.size() + 1) / 2);
        for (size_t i = 0; i + 1 < level.size(); i += 2) node_digest(level[i].data(), level[i + 1].data(), nxt[i / 2].data());
        if (level.size() & 1) nxt.back() = level.back();
        level.swap(nxt);
    }
    memcpy(out, level[0].data(), 32);
}
// sibling digests bottom -> top for subtree `owner` among the G subtree roots
size_t top_path(const uint8_t* roots, unsigned g, unsigned owner, uint8_t* out) {
    std::vector<std::array<uint8_t, 32>> level(g);
    for (unsigned i = 0; i < g; i++) memcpy(level[i].data(), roots + 32 * i, 32);
    size_t w = 0, j = owner;
    while (level.size() > 1) {
        size_t sib = j ^ 1;
        if (sib < level.size()) { memcpy(out + w, level[sib].data(), 32); w += 32; }
        std::vector<std::array<uint8_t, 32>> nxt((level.size() + 1) / 2);
        for (size_t i = 0; i + 1 < level.size(); i += 2) node_digest(level[i].data(), level[i + 1].data(), nxt[i / 2].data());
        if (level.size() & 1) nxt.back() = level.back();
        level.swap(nxt);
        j >>= 1;
    }
    return w;
}

struct P2PState {                       // peer-memory four-step buffers for one transform size
    unsigned log_n = 0;
    stark_vec *rows = nullptr, *block = nullptr, *flags = nullptr;
    std::vector<void*> peer_rows, peer_blocks, peer_flags, opened;
    uint32_t epoch = 0;
};
struct StagedState {                    // NCCL transport: send / receive staging and the two working arrays
    unsigned log_n = 0;
    DevBufPtr send, recv, rows, block;
};

}  // namespace

struct stark_mg {
    CtxRef ctx;
    unsigned rank = 0, world = 1;
    nccl_comm_t comm = nullptr;
    bool own_comm = false;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_ready = nullptr, ev_done = nullptr;   // main stream -> copy stream -> main stream (allgather_layer_begin / _end)
    DevBuf d_small;                     // device scratch of the small collectives
    PinnedBuf h_small;                  // their pinned host side; also the per-column root slots of commit_columns
    std::map<unsigned, std::unique_ptr<P2PState>> p2p;
    std::map<unsigned, std::unique_ptr<StagedState>> staged;
};
// One FRI layer >= 1 hashed in leaf ranges: this rank's range of the values (a private copy), its subtree, all subtree roots.
struct ShardedLayer {
    std::unique_ptr<stark_vec> block;
    std::unique_ptr<stark_tree> subtree;
    std::vector<uint8_t> subtree_roots;  // world * 32
    size_t n = 0;                        // leaves of the whole layer
};
struct stark_mg_fri {
    stark_mg* mg = nullptr;
    unsigned log_n = 0;
    stark_vec* block = nullptr;          // this rank's leaf range of layer 0 (borrowed from the transport when owned == false)
    stark_tree* subtree = nullptr;       // its tree
    std::vector<uint8_t> subtree_roots;  // world * 32
    std::vector<ShardedLayer> sharded;   // layers 1 .. S, hashed in leaf ranges as well (sharded_fri_layers)
    stark_fri* proof = nullptr;          // rank 0 (every rank when layers >= 1 are sharded: the folds are replicated):
                                         // layers 0 .. S adopted, the later ones built here on rank 0
};

namespace {

struct MgGuard {
    std::lock_guard<std::recursive_mutex> lk;
    explicit MgGuard(stark_mg* m) : lk(m->ctx->mu) { STARK_CUDA(cudaSetDevice(m->ctx->device)); }
};
#define MG_BEGIN try {
#define MG_END                                                                              \
    }                                                                                       \
    catch (const StarkError& e) { api_set_error(e.what()); return e.code; }                 \
    catch (const std::bad_alloc&) { api_set_error("host out of memory"); return ST_INTERNAL; } \
    catch (const std::exception& e) { api_set_error(e.what()); return ST_INTERNAL; }        \
    return ST_OK;

constexpr size_t SMALL_BYTES = (size_t)1 << 20;

// every rank contributes `bytes` (<= 512 KiB / (world + 1)): out = world * bytes, identical on every rank
void gather_bytes(stark_mg* mg, const void* in, size_t bytes, void* out) {
    if (mg->world == 1) { memcpy(out, in, bytes); return; }
    STARK_REQUIRE(bytes * (mg->world + 1) <= SMALL_BYTES / 2, "gather_bytes: payload too large");
    stark_ctx* ctx = mg->ctx;
    uint8_t* h = static_cast<uint8_t*>(mg->h_small.h);
    uint8_t* d = mg->d_small.as<uint8_t>();
    STARK_CUDA(cudaStreamSynchronize(ctx->stream));                 // the staging areas may still be in use by an earlier call
    memcpy(h, in, bytes);
    STARK_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
    STARK_NCCL(nccl().AllGather(d, d + bytes, bytes, NCCL_UINT8, mg->comm, ctx->stream));
    STARK_CUDA(cudaMemcpyAsync(h + bytes, d + bytes, bytes * mg->world, cudaMemcpyDeviceToHost, ctx->stream));
    STARK_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(out, h + bytes, bytes * mg->world);
}
uint64_t bcast_u64(stark_mg* mg, uint64_t v) {
    if (mg->world == 1) return v;
    stark_ctx* ctx = mg->ctx;
    uint64_t* h = static_cast<uint64_t*>(mg->h_small.h);
    uint64_t* d = mg->d_small.as<uint64_t>();
    STARK_CUDA(cudaStreamSynchronize(ctx->stream));
    *h = v;
    STARK_CUDA(cudaMemcpyAsync(d, h, 8, cudaMemcpyHostToDevice, ctx->stream));
    STARK_NCCL(nccl().Broadcast(d, d, 8, NCCL_UINT8, 0, mg->comm, ctx->stream));
    STARK_CUDA(cudaMemcpyAsync(h, d, 8, cudaMemcpyDeviceToHost, ctx->stream));
    STARK_CUDA(cudaStreamSynchronize(ctx->stream));
    return *h;
}
// all-to-all of equal chunks (u32 words) on the context's stream
void all_to_all(stark_mg* mg, const uint32_t* send, uint32_t* recv, size_t chunk) {
    stark_ctx* ctx = mg->ctx;
    if (mg->world == 1) {
        STARK_CUDA(cudaMemcpyAsync(recv, send, chunk * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        return;
    }
    Nccl& n = nccl();
    STARK_NCCL(n.GroupStart());
    for (unsigned p = 0; p < mg->world; p++) {
        STARK_NCCL(n.Send(send + p * chunk, chunk, NCCL_UINT32, (int)p, mg->comm, ctx->stream));
        STARK_NCCL(n.Recv(recv + p * chunk, chunk, NCCL_UINT32, (int)p, mg->comm, ctx->stream));
    }
    STARK_NCCL(n.GroupEnd());
}

void mg_common_init(stark_mg* mg) {
    stark_ctx* ctx = mg->ctx;
    STARK_REQUIRE(mg->world >= 1 && mg->world <= (unsigned)MAX_PEERS && mg->rank < mg->world, "stark_mg: rank / world out of range");
    STARK_CUDA(cudaStreamCreateWithFlags(&mg->copy_stream, cudaStreamNonBlocking));
    STARK_CUDA(cudaEventCreateWithFlags(&mg->ev_ready, cudaEventDisableTiming));
    STARK_CUDA(cudaEventCreateWithFlags(&mg->ev_done, cudaEventDisableTiming));
    mg->d_small = DevBuf(SMALL_BYTES, ctx->stream);
    mg->h_small.ensure(SMALL_BYTES);
}

P2PState* p2p_state(stark_mg* mg, unsigned log_n) {
    auto it = mg->p2p.find(log_n);
    if (it != mg->p2p.end()) return it->second.get();
    stark_ctx* ctx = mg->ctx;
    unsigned log_g = ilog2(mg->world);
    STARK_REQUIRE(((unsigned)1 << log_g) == mg->world, "four-step NTT: the world size must be a power of two");
    auto st = std::make_unique<P2PState>();
    st->log_n = log_n;
    const size_t n_loc = ((size_t)1 << log_n) >> log_g;
    uint8_t hs[3][64];
    int rc = stark_peer_alloc(ctx, n_loc, &st->rows, hs[0]);
    if (rc == ST_OK) rc = stark_peer_alloc(ctx, n_loc, &st->block, hs[1]);
    if (rc == ST_OK) rc = stark_peer_alloc(ctx, 2 * MAX_PEERS, &st->flags, hs[2]);
    if (rc != ST_OK) throw StarkError(rc, stark_last_error());
    STARK_CUDA(cudaStreamSynchronize(ctx->stream));                 // the zero fills are done before anyone can write
    std::vector<uint8_t> all((size_t)mg->world * 192);
    gather_bytes(mg, hs, 192, all.data());
    st->peer_rows.resize(mg->world); st->peer_blocks.resize(mg->world); st->peer_flags.resize(mg->world);
    for (unsigned r = 0; r < mg->world; r++) {
        if (r == mg->rank) {
            st->peer_rows[r] = st->rows->buf->p; st->peer_blocks[r] = st->block->buf->p; st->peer_flags[r] = st->flags->buf->p;
            continue;
        }
        void* ptr[3];
        for (int k = 0; k < 3; k++) {
            rc = stark_peer_open(ctx, all.data() + (size_t)r * 192 + 64 * k, &ptr[k]);
            if (rc != ST_OK) throw StarkError(rc, stark_last_error());
            st->opened.push_back(ptr[k]);
        }
        st->peer_rows[r] = ptr[0]; st->peer_blocks[r] = ptr[1]; st->peer_flags[r] = ptr[2];
    }
    // nobody may store into a peer before that peer's buffers exist and are zeroed: one gather as a barrier
    uint8_t token = 1;
    std::vector<uint8_t> tokens(mg->world);
    gather_bytes(mg, &token, 1, tokens.data());
    P2PState* raw = st.get();
    mg->p2p[log_n] = std::move(st);
    return raw;
}
void p2p_release(stark_mg* mg, P2PState* st) {
    for (void* p : st->opened) stark_peer_close(mg->ctx, p);
    stark_vec_destroy(st->rows); stark_vec_destroy(st->block); stark_vec_destroy(st->flags);
}

PeerPtrs chunk_ptrs(uint32_t* base, size_t chunk, unsigned world) {
    PeerPtrs p{};
    for (unsigned i = 0; i < world; i++) p.p[i] = base + (size_t)i * chunk;
    return p;
}

// this rank's natural-order block of the evaluations of `coeffs` on offset * <w_{2^log_n}>; returns the buffer that holds it
DevBufPtr fourstep_lde(stark_mg* mg, const stark_vec* coeffs, unsigned log_n, uint64_t offset, int transport) {
    stark_ctx* ctx = mg->ctx;
    const unsigned world = mg->world, rank = mg->rank, log_g = ilog2(world);
    STARK_REQUIRE(((unsigned)1 << log_g) == world, "four-step NTT: the world size must be a power of two");
    const unsigned a = log_n / 2, b = log_n - a;
    STARK_REQUIRE(log_n <= ctx->two_adicity && log_n <= 30, "fourstep: 2^log_n does not divide p-1");
    STARK_REQUIRE(a >= log_g + 5 && b >= log_g + 5, "fourstep: every rank needs >= 32 rows and >= 32 columns (raise log_n or lower the world size)");
    STARK_REQUIRE(coeffs->n <= ((size_t)1 << log_n), "fourstep: more coefficients than domain points");
    STARK_REQUIRE(offset % ctx->modulus != 0, "coset offset must be non-zero");
    if (transport == 1) {
        P2PState* st = p2p_state(mg, log_n);
        const uint32_t epoch = ++st->epoch;
        PeerPtrs rows{}, blocks{};
        for (unsigned r = 0; r < world; r++) { rows.p[r] = static_cast<uint32_t*>(st->peer_rows[r]); blocks.p[r] = static_cast<uint32_t*>(st->peer_blocks[r]); }
        fourstep_phase_a_launch(ctx, coeffs->buf->as<uint32_t>(), coeffs->n, log_n, offset, world, rank, rows, false, st->peer_flags.data(), epoch);
        fourstep_wait(ctx, st->peer_flags[rank], 0, world, epoch);
        fourstep_phase_c_launch(ctx, st->rows->buf->as<uint32_t>(), log_n, world, rank, blocks, false, st->peer_flags.data(), epoch);
        fourstep_wait(ctx, st->peer_flags[rank], 1, world, epoch);
        return st->block->buf;
    }
    STARK_REQUIRE(transport == 0, "fourstep: transport must be 0 (NCCL all-to-all) or 1 (peer-memory stores)");
    auto& slot = mg->staged[log_n];
    if (!slot) {
        slot = std::make_unique<StagedState>();
        slot->log_n = log_n;
        const size_t bytes = (((size_t)1 << log_n) >> log_g) * 4;
        slot->send = make_buf(bytes, ctx->stream); slot->recv = make_buf(bytes, ctx->stream);
        slot->rows = make_buf(bytes, ctx->stream); slot->block = make_buf(bytes, ctx->stream);
    }
    StagedState* st = slot.get();
    const size_t n1 = (size_t)1 << a, n2 = (size_t)1 << b, w = n2 >> log_g, rpr = n1 >> log_g;      // columns / rows per rank
    const size_t chunk = rpr * w;
    uint32_t *send = st->send->as<uint32_t>(), *recv = st->recv->as<uint32_t>(), *rows = st->rows->as<uint32_t>(), *block = st->block->as<uint32_t>();
    // phase A into per-destination chunks [N1/G][w]; exchange; chunk of source r -> columns r*w .. of the [N1/G][N2] matrix
    fourstep_phase_a_launch(ctx, coeffs->buf->as<uint32_t>(), coeffs->n, log_n, offset, world, rank, chunk_ptrs(send, chunk, world), true, nullptr, 0);
    all_to_all(mg, send, recv, chunk);
    for (unsigned r = 0; r < world; r++)
        STARK_CUDA(cudaMemcpy2DAsync(rows + r * w, n2 * 4, recv + r * chunk, w * 4, w * 4, rpr, cudaMemcpyDeviceToDevice, ctx->stream));
    // phase C into chunks [N2/G][N1/G]; exchange; chunk of source r -> positions r*N1/G .. of every k2 row of the block
    fourstep_phase_c_launch(ctx, rows, log_n, world, rank, chunk_ptrs(send, chunk, world), true, nullptr, 0);
    all_to_all(mg, send, recv, chunk);
    for (unsigned r = 0; r < world; r++)
        STARK_CUDA(cudaMemcpy2DAsync(block + r * rpr, n1 * 4, recv + r * chunk, rpr * 4, rpr * 4, w, cudaMemcpyDeviceToDevice, ctx->stream));
    return st->block;
}

// tree over this rank's block + gather of the subtree roots + the levels above them
std::unique_ptr<stark_tree> commit_leaf_range(stark_mg* mg, DevBufPtr block, size_t n_loc, uint8_t root[32], uint8_t* subtree_roots) {
    stark_ctx* ctx = mg->ctx;
    STARK_REQUIRE(n_loc >= 1 && (n_loc & (n_loc - 1)) == 0, "leaf ranges must be powers of two (exact subtrees)");
    auto t = api_tree_launch_values(ctx, std::move(block), n_loc, ctx->d_result);
    STARK_CUDA(cudaStreamSynchronize(ctx->stream));
    STARK_REQUIRE(ctx->h_result->flag == 0, "a peer never finished its part of the four-step exchange (device-side wait timed out)");
    memcpy(t->root_words, ctx->h_result->root, 32);
    uint8_t mine[32];
    words_to_bytes(t->root_words, mine);
    std::vector<uint8_t> all((size_t)mg->world * 32);
    gather_bytes(mg, mine, 32, all.data());
    combine_roots(all.data(), mg->world, root);
    if (subtree_roots) memcpy(subtree_roots, all.data(), all.size());
    return t;
}

// Openings of leaves of columns committed in leaf ranges, any number of columns / layers in ONE exchange: each owner opens
// its records with one launch, one all-gather tells everybody, rank 0 appends the levels above the subtree roots and
// feeds (element, path) pairs to the channel in the order of `items`.
struct LeafRangeSet {
    const stark_vec* block;            // this rank's range of the values
    const stark_tree* subtree;         // its tree
    const uint8_t* subtree_roots;      // world * 32
    size_t n_total;                    // leaves of the whole column
};
struct OpenItem { const LeafRangeSet* set; size_t which; };
void open_leaf_ranges_batch(stark_mg* mg, const std::vector<OpenItem>& items, Channel* ch) {
    const unsigned world = mg->world, rank = mg->rank;
    std::vector<size_t> off(items.size() + 1, 0);
    std::vector<OpenDesc> mine;
    for (size_t t = 0; t < items.size(); t++) {
        const LeafRangeSet& s = *items[t].set;
        const size_t blk = s.n_total / world;
        STARK_REQUIRE(items[t].which < s.n_total, "open_leaf_ranges: leaf index out of range");
        off[t + 1] = off[t] + 8 + 32 * ilog2(blk);
        if (items[t].which / blk == rank)
            mine.push_back(OpenDesc{s.block->buf->as<uint32_t>(), s.subtree->nodes.as<uint32_t>(), blk, items[t].which % blk, off[t]});
    }
    const size_t bytes = off.back();
    std::vector<uint8_t> payload(bytes, 0), all((size_t)world * bytes), top(32 * 8);
    if (!mine.empty()) api_open_records(mg->ctx, mine, bytes, payload.data());
    gather_bytes(mg, payload.data(), bytes, all.data());
    if (rank != 0) return;
    std::vector<uint8_t> path;
    for (size_t t = 0; t < items.size(); t++) {
        const LeafRangeSet& s = *items[t].set;
        const unsigned owner = (unsigned)(items[t].which / (s.n_total / world));
        const uint8_t* rec = all.data() + (size_t)owner * bytes + off[t];
        path.assign(rec + 8, rec + (off[t + 1] - off[t]));
        const size_t tl = top_path(s.subtree_roots, world, owner, top.data());
        path.insert(path.end(), top.begin(), top.begin() + tl);
        ch->send(rec, 8);
        ch->send(path.data(), path.size());
    }
}
void open_leaf_ranges(stark_mg* mg, const stark_vec* block, const stark_tree* subtree, const uint8_t* subtree_roots, const size_t* which,
                      size_t count, size_t n_total, Channel* ch) {
    LeafRangeSet set{block, subtree, subtree_roots, n_total};
    std::vector<OpenItem> items;
    for (size_t t = 0; t < count; t++) items.push_back(OpenItem{&set, which[t]});
    open_leaf_ranges_batch(mg, items, ch);
}
// idx and its FRI sibling in every layer hashed in leaf ranges (fri_commit.rs:152-153)
void sharded_layer_items(const std::vector<ShardedLayer>& layers, std::deque<LeafRangeSet>& sets, size_t index, std::vector<OpenItem>& items) {
    const size_t first = sets.size();
    for (const ShardedLayer& l : layers) sets.push_back(LeafRangeSet{l.block.get(), l.subtree.get(), l.subtree_roots.data(), l.n});
    for (size_t k = 0; k < layers.size(); k++) {
        const size_t len = layers[k].n, idx = index % len;
        items.push_back(OpenItem{&sets[first + k], idx});
        items.push_back(OpenItem{&sets[first + k], (idx + len / 2) % len});
    }
}

stark_vec* wrap_vec(stark_ctx* ctx, DevBufPtr b, size_t n) {
    stark_vec* v = new stark_vec(); v->ctx = ctx; v->buf = std::move(b); v->n = n; return v;
}

stark_vec* wrap_vec(stark_ctx* ctx, DevBufPtr b, size_t n);
std::unique_ptr<stark_tree> commit_leaf_range(stark_mg* mg, DevBufPtr block, size_t n_loc, uint8_t root[32], uint8_t* subtree_roots);
// ---- FRI layers >= 1 while they are large: leaf hashing in ranges, everything else replicated -------------------------
// The north-star partitions leaf hashing and nothing else.  A FRI layer is 12 bytes of HBM traffic per point to fold and
// 1384 integer instructions per leaf (plus the nodes) to commit, so the fold is replicated -- every rank holds layer 0 (one
// all-gather instead of the gather on rank 0), folds the WHOLE next layer and the coefficients itself, and no data-path
// exchange follows -- while the tree is hashed in leaf ranges: 32 bytes of subtree root per rank travel to everybody, the
// top log2(world) levels are finished on the host, rank 0 feeds the channel and beta (8 bytes) travels back.  Layers below
// 2^STARK_MG_FRI_SHARD_MIN_LOG leaves (default 2^20: a layer costs ~0.16 ns per leaf to hash against ~0.1 ms for the two
// small collectives and the extra launches; measured on 8 GPUs at a 2^26 domain, thresholds 2^25 / 23 / 22 / 21 / 20 / 19 / 18 /
// 17 -> 11.5 / 8.30 / 7.81 / 7.53 / 7.45 / 7.43 / 7.48 / 7.53 ms) are left to rank 0's ordinary loop.
unsigned shard_min_log() {
    const char* e = getenv("STARK_MG_FRI_SHARD_MIN_LOG");
    const int v = e ? atoi(e) : 0;
    return v >= 6 && v <= 30 ? (unsigned)v : 20u;
}
bool shard_forced() { const char* e = getenv("STARK_MG_FRI_SHARD_FORCE"); return e && atoi(e) != 0; }      // tests: one rank takes the sharded path
bool shard_wanted(const stark_mg* mg, unsigned log_n) {
    if (mg->world == 1 && !shard_forced()) return false;
    const char* off = getenv("STARK_MG_FRI_SHARD");
    if (off && atoi(off) == 0) return false;
    return log_n >= 1 && log_n - 1 >= shard_min_log() && (((size_t)1 << (log_n - 1)) / mg->world) >= 2;
}
// Every rank holds `proof` with identical layers and coefficients; `chan` on rank 0 only.  Collective.
void sharded_fri_layers(stark_mg* mg, stark_fri* proof, Channel* chan, std::vector<ShardedLayer>& out) {
    stark_ctx* ctx = mg->ctx;
    const unsigned world = mg->world, rank = mg->rank;
    const size_t min_len = (size_t)1 << shard_min_log();
    while ((long long)proof->coeff_len - 1 >= 1 && proof->cur_log >= 1) {                     // fri_commit.rs:89
        const size_t half = ((size_t)1 << proof->cur_log) >> 1, blk = half / world;
        if (half < min_len || blk < 2) break;
        uint64_t beta = 0;
        if (rank == 0) STARK_REQUIRE(chan->receive_random_field_element(&beta), "channel: receive before send");     // :91
        beta = bcast_u64(mg, beta);
        DevBufPtr layer = api_fri_fold_values(proof, beta);                                    // :94, whole layer, every rank
        DevBufPtr mine = make_buf(blk * 4, ctx->stream);
        STARK_CUDA(cudaMemcpyAsync(mine->p, layer->as<uint32_t>() + (size_t)rank * blk, blk * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        ShardedLayer sl;
        sl.n = half;
        sl.subtree_roots.resize((size_t)world * 32);
        uint8_t root[32];
        sl.subtree = commit_leaf_range(mg, mine, blk, root, sl.subtree_roots.data());           // :97 (synchronises: the degree is in too)
        api_fri_adopt_layer(proof, layer, root);
        if (rank == 0) api_send_root_bytes(*chan, root);                                       // :100
        sl.block.reset(wrap_vec(ctx, mine, blk));
        out.push_back(std::move(sl));
    }
}
// Layer 0 of a sharded proof on every rank: this rank's block + everybody else's.  The all-gather (4 N bytes per rank over
// NVLink) is issued on the copy stream BEFORE the leaf-range hashing of the same block is launched on the main stream, so
// the two overlap (the collective's few CTAs are placed first); allgather_layer_end makes the main stream wait for it.
DevBufPtr allgather_layer_begin(stark_mg* mg, const DevBufPtr& block, size_t blk) {
    stark_ctx* ctx = mg->ctx;
    if (mg->world == 1) return block;
    DevBufPtr all = make_buf(blk * mg->world * 4, ctx->stream);
    STARK_CUDA(cudaEventRecord(mg->ev_ready, ctx->stream));
    STARK_CUDA(cudaStreamWaitEvent(mg->copy_stream, mg->ev_ready, 0));
    STARK_NCCL(nccl().AllGather(block->p, all->p, blk, NCCL_UINT32, mg->comm, mg->copy_stream));
    STARK_CUDA(cudaEventRecord(mg->ev_done, mg->copy_stream));
    return all;
}
void allgather_layer_end(stark_mg* mg) {
    if (mg->world > 1) STARK_CUDA(cudaStreamWaitEvent(mg->ctx->stream, mg->ev_done, 0));
}

}  // namespace

// ======================================================================================= group
extern "C" int stark_mg_unique_id(uint8_t id[128]) {
    MG_BEGIN
    STARK_REQUIRE(id, "mg_unique_id: null argument");
    static_assert(sizeof(nccl_unique_id) == 128, "ncclUniqueId size");
    nccl_unique_id u;
    STARK_NCCL(nccl().GetUniqueId(&u));
    memcpy(id, &u, 128);
    MG_END
}
extern "C" int stark_mg_create(stark_ctx* ctx, const uint8_t id[128], unsigned rank, unsigned world, stark_mg** out) {
    MG_BEGIN
    STARK_REQUIRE(ctx && out && (id || world == 1), "mg_create: null argument");
    *out = nullptr;
    std::unique_ptr<stark_mg> mg(new stark_mg());
    mg->ctx = ctx; mg->rank = rank; mg->world = world;
    MgGuard g(mg.get());
    mg_common_init(mg.get());
    if (world > 1) {
        nccl_unique_id u;
        memcpy(&u, id, 128);
        STARK_NCCL(nccl().CommInitRank(&mg->comm, (int)world, u, (int)rank));
        mg->own_comm = true;
    }
    *out = mg.release();
    MG_END
}
extern "C" int stark_mg_adopt(stark_ctx* ctx, void* nccl_comm, unsigned rank, unsigned world, stark_mg** out) {
    MG_BEGIN
    STARK_REQUIRE(ctx && out && (nccl_comm || world == 1), "mg_adopt: null argument");
    *out = nullptr;
    std::unique_ptr<stark_mg> mg(new stark_mg());
    mg->ctx = ctx; mg->rank = rank; mg->world = world; mg->comm = static_cast<nccl_comm_t>(nccl_comm);
    MgGuard g(mg.get());
    if (world > 1) nccl();
    mg_common_init(mg.get());
    *out = mg.release();
    MG_END
}
extern "C" void stark_mg_destroy(stark_mg* mg) {
    if (!mg) return;
    cudaSetDevice(mg->ctx->device);
    cudaStreamSynchronize(mg->ctx->stream);
    if (mg->world > 1 && !mg->p2p.empty()) {          // nobody unmaps a buffer a peer may still be storing into
        try { uint8_t t = 1; std::vector<uint8_t> all(mg->world); gather_bytes(mg, &t, 1, all.data()); } catch (...) {}
    }
    for (auto& kv : mg->p2p) p2p_release(mg, kv.second.get());
    mg->p2p.clear(); mg->staged.clear();
    mg->d_small.release(); mg->h_small.release();
    if (mg->copy_stream) cudaStreamDestroy(mg->copy_stream);
    if (mg->ev_ready) cudaEventDestroy(mg->ev_ready);
    if (mg->ev_done) cudaEventDestroy(mg->ev_done);
    if (mg->own_comm && mg->comm) nccl().CommDestroy(mg->comm);
    delete mg;
}
extern "C" unsigned stark_mg_rank(const stark_mg* mg) { return mg ? mg->rank : 0; }
extern "C" unsigned stark_mg_world(const stark_mg* mg) { return mg ? mg->world : 0; }
extern "C" int stark_mg_barrier(stark_mg* mg) {
    MG_BEGIN
    STARK_REQUIRE(mg, "mg_barrier: null argument");
    MgGuard g(mg);
    uint8_t t = 1;
    std::vector<uint8_t> all(mg->world);
    gather_bytes(mg, &t, 1, all.data());
    MG_END
}

// ======================================================================================= cfg4: column-parallel commit
// Column c (2^log_rows evaluations on offset_in * <g>, host memory; pinned memory lets the upload of the next column run
// under the hashing of the current one) goes to rank c mod world: coset LDE to offset_out * <h> (blow-up 2^log_blowup)
// and a Merkle tree, all enqueued without a host round trip per column; the roots of ALL columns come back on every rank.
extern "C" int stark_mg_commit_columns(stark_mg* mg, size_t n_cols, const uint64_t* const* columns, unsigned log_rows, uint64_t offset_in,
                                       unsigned log_blowup, uint64_t offset_out, uint8_t* roots, stark_vec** ldes, stark_tree** trees) {
    MG_BEGIN
    STARK_REQUIRE(mg && columns && roots && n_cols >= 1, "mg_commit_columns: null argument");
    MgGuard g(mg);
    stark_ctx* ctx = mg->ctx;
    const unsigned world = mg->world, rank = mg->rank;
    STARK_REQUIRE(log_rows + log_blowup <= 30 && log_rows + log_blowup <= ctx->two_adicity, "mg_commit_columns: domain too large for this field");
    STARK_REQUIRE(offset_in % ctx->modulus != 0 && offset_out % ctx->modulus != 0, "coset offset must be non-zero");
    const size_t n = (size_t)1 << log_rows, N = n << log_blowup, per_rank = (n_cols + world - 1) / world;
    STARK_REQUIRE(per_rank * 32 * (world + 1) <= SMALL_BYTES / 2 && per_rank * sizeof(HostResult) <= SMALL_BYTES / 2, "mg_commit_columns: too many columns per call");
    std::vector<size_t> mine;
    for (size_t c = rank; c < n_cols; c += world) { STARK_REQUIRE(columns[c], "mg_commit_columns: an owned column is null"); mine.push_back(c); }
    if (ldes) for (size_t c = 0; c < n_cols; c++) ldes[c] = nullptr;
    if (trees) for (size_t c = 0; c < n_cols; c++) trees[c] = nullptr;
    // per-column root slots in mapped pinned memory (upper half of h_small; the collectives use the lower half)
    STARK_CUDA(cudaStreamSynchronize(ctx->stream));
    HostResult* h_slots = reinterpret_cast<HostResult*>(static_cast<uint8_t*>(mg->h_small.h) + SMALL_BYTES / 2);
    HostResult* d_slots = reinterpret_cast<HostResult*>(static_cast<uint8_t*>(mg->h_small.d) + SMALL_BYTES / 2);
    DevBuf stage[2] = {DevBuf(n * 8, ctx->stream), DevBuf(n * 8, ctx->stream)};
    cudaEvent_t up[2], used[2];
    for (int i = 0; i < 2; i++) { STARK_CUDA(cudaEventCreateWithFlags(&up[i], cudaEventDisableTiming)); STARK_CUDA(cudaEventCreateWithFlags(&used[i], cudaEventDisableTiming)); }
    std::vector<std::unique_ptr<stark_tree>> built;
    std::vector<DevBufPtr> lde_keep;
    try {
        STARK_CUDA(cudaEventRecord(used[0], ctx->stream)); STARK_CUDA(cudaEventRecord(used[1], ctx->stream));   // both stages exist from here on
        auto enqueue_upload = [&](size_t k) {
            const int s = (int)(k & 1);
            STARK_CUDA(cudaStreamWaitEvent(mg->copy_stream, used[s], 0));
            STARK_CUDA(cudaMemcpyAsync(stage[s].p, columns[mine[k]], n * 8, cudaMemcpyHostToDevice, mg->copy_stream));
            STARK_CUDA(cudaEventRecord(up[s], mg->copy_stream));
        };
        if (!mine.empty()) enqueue_upload(0);
        for (size_t k = 0; k < mine.size(); k++) {
            const int s = (int)(k & 1);
            if (k + 1 < mine.size()) enqueue_upload(k + 1);
            STARK_CUDA(cudaStreamWaitEvent(ctx->stream, up[s], 0));
            DevBufPtr col = make_buf(n * 4, ctx->stream);
            narrow_u64(ctx, stage[s].as<uint64_t>(), col->as<uint32_t>(), n);
            STARK_CUDA(cudaEventRecord(used[s], ctx->stream));
            DevBufPtr lde = api_lde_on_coset(ctx, col->as<uint32_t>(), log_rows, offset_in, log_blowup, offset_out);
            built.push_back(api_tree_launch_values(ctx, lde, N, d_slots + k));
            lde_keep.push_back(lde);
            if (!trees) { built.back().reset(); }                     // stream-ordered frees: the next column reuses the blocks
            if (!ldes) lde_keep.back().reset();
        }
        STARK_CUDA(cudaStreamSynchronize(ctx->stream));
        STARK_CUDA(cudaStreamSynchronize(mg->copy_stream));
    } catch (...) {
        for (int i = 0; i < 2; i++) { cudaEventDestroy(up[i]); cudaEventDestroy(used[i]); }
        throw;
    }
    for (int i = 0; i < 2; i++) { cudaEventDestroy(up[i]); cudaEventDestroy(used[i]); }
    std::vector<uint8_t> local(per_rank * 32, 0), all((size_t)world * per_rank * 32);
    for (size_t k = 0; k < mine.size(); k++) {
        words_to_bytes(h_slots[k].root, local.data() + 32 * k);
        if (trees) { memcpy(built[k]->root_words, h_slots[k].root, 32); trees[mine[k]] = built[k].release(); }
        if (ldes) ldes[mine[k]] = wrap_vec(ctx, lde_keep[k], N);
    }
    gather_bytes(mg, local.data(), local.size(), all.data());
    for (size_t c = 0; c < n_cols; c++) memcpy(roots + 32 * c, all.data() + ((c % world) * per_rank + c / world) * 32, 32);
    MG_END
}

// ======================================================================================= cfg5: one big column
// transport 0: exchanges over NCCL (ncclSend / ncclRecv groups between staged chunks);  1: the producing kernels store
// straight into peer memory over NVLink (CUDA IPC mappings, device-side epoch flags between the phases).
// *block: this rank's natural-order range [rank * N/G, (rank+1) * N/G) of the evaluations.  It aliases a buffer of the
// group that the next stark_mg_fourstep_lde of the same size and transport overwrites; work that reads it must be
// enqueued on this context before that call.
extern "C" int stark_mg_fourstep_lde(stark_mg* mg, const stark_vec* coeffs, unsigned log_n, uint64_t offset, int transport, stark_vec** block) {
    MG_BEGIN
    STARK_REQUIRE(mg && coeffs && block && coeffs->ctx == mg->ctx, "mg_fourstep_lde: bad argument");
    MgGuard g(mg);
    DevBufPtr b = fourstep_lde(mg, coeffs, log_n, offset, transport);
    *block = wrap_vec(mg->ctx, b, ((size_t)1 << log_n) / mg->world);
    MG_END
}
// Tree over this rank's contiguous leaf range (an exact subtree of the rs_merkle shape), subtree roots all-gathered, the
// top log2(world) levels finished identically on every rank.  subtree_roots (optional): world * 32 bytes.
extern "C" int stark_mg_commit_leaf_ranges(stark_mg* mg, const stark_vec* block, stark_tree** subtree, uint8_t root[32], uint8_t* subtree_roots) {
    MG_BEGIN
    STARK_REQUIRE(mg && block && subtree && root && block->ctx == mg->ctx, "mg_commit_leaf_ranges: bad argument");
    MgGuard g(mg);
    *subtree = commit_leaf_range(mg, block->buf, block->n, root, subtree_roots).release();
    MG_END
}

// fri_commit (src/fri/fri_commit.rs:72-122) with layer 0 spread over the group: four-step LDE -> every rank hashes its
// leaf range -> subtree roots gathered -> the evaluations are collected on rank 0, which runs the unpartitioned
// fold / commit loop against `ch` (ignored on the other ranks).  The transcript is the single-GPU one, byte for byte.
extern "C" int stark_mg_fri_commit(stark_mg* mg, const stark_vec* coeffs, unsigned log_n, uint64_t offset, int transport,
                                   stark_channel* ch, stark_mg_fri** out) {
    MG_BEGIN
    STARK_REQUIRE(mg && coeffs && out && coeffs->ctx == mg->ctx && (ch || mg->rank != 0), "mg_fri_commit: bad argument");
    *out = nullptr;
    MgGuard g(mg);
    stark_ctx* ctx = mg->ctx;
    const unsigned world = mg->world, rank = mg->rank;
    std::unique_ptr<stark_mg_fri> f(new stark_mg_fri());
    f->mg = mg; f->log_n = log_n;
    const size_t N = (size_t)1 << log_n, blk = N / world;
    DevBufPtr b = fourstep_lde(mg, coeffs, log_n, offset, transport);
    f->subtree_roots.resize((size_t)world * 32);
    uint8_t root0[32];
    // layer 0: on every rank when the following layers are hashed in leaf ranges too (the folds are replicated), else on rank 0
    const bool shard = shard_wanted(mg, log_n);
    DevBufPtr layer0;
    if (world > 1 && shard) layer0 = allgather_layer_begin(mg, b, blk);        // runs under the hashing below
    auto sub = commit_leaf_range(mg, b, blk, root0, f->subtree_roots.data());
    if (world == 1) layer0 = b;
    else if (shard) allgather_layer_end(mg);
    else {
        Nccl& n = nccl();
        if (rank == 0) {
            layer0 = make_buf(N * 4, ctx->stream);
            STARK_CUDA(cudaMemcpyAsync(layer0->p, b->p, blk * 4, cudaMemcpyDeviceToDevice, ctx->stream));
            STARK_NCCL(n.GroupStart());
            for (unsigned p = 1; p < world; p++) STARK_NCCL(n.Recv(layer0->as<uint32_t>() + p * blk, blk, NCCL_UINT32, (int)p, mg->comm, ctx->stream));
            STARK_NCCL(n.GroupEnd());
        } else {
            STARK_NCCL(n.Send(b->p, blk, NCCL_UINT32, 0, mg->comm, ctx->stream));
        }
    }
    // the leaf range must outlive the transport's next run: keep a private copy when it aliases the transport's buffer
    DevBufPtr keep = make_buf(blk * 4, ctx->stream);
    STARK_CUDA(cudaMemcpyAsync(keep->p, b->p, blk * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    sub->leaves = keep;
    f->block = wrap_vec(ctx, keep, blk);
    f->subtree = sub.release();
    if (rank == 0 || shard) {
        stark_vec l0; l0.ctx = ctx; l0.buf = layer0; l0.n = N;
        int rc = stark_fri_begin_external(ctx, coeffs, log_n, offset, &l0, root0, &f->proof);
        if (rc == ST_OK && !shard) rc = api_fri_commit_loop(f->proof, ch);
        if (rc == ST_OK && shard) {
            if (rank == 0) api_send_root_bytes(ch->ch, root0);                                 // fri_commit.rs:86
            sharded_fri_layers(mg, f->proof, rank == 0 ? &ch->ch : nullptr, f->sharded);
            if (rank == 0) rc = api_fri_commit_loop_resume(f->proof, ch);                      // the small layers and the final constant
        }
        if (rc != ST_OK) { stark_mg_fri_destroy(f.release()); return rc; }
    }
    STARK_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = f.release();
    MG_END
}
extern "C" void stark_mg_fri_destroy(stark_mg_fri* f) {
    if (!f) return;
    cudaSetDevice(f->mg->ctx->device);
    if (f->proof) stark_fri_destroy(f->proof);
    if (f->subtree) stark_tree_destroy(f->subtree);
    if (f->block) stark_vec_destroy(f->block);
    delete f;
}
extern "C" const stark_fri* stark_mg_fri_proof(const stark_mg_fri* f) { return f ? f->proof : nullptr; }
extern "C" const stark_tree* stark_mg_fri_subtree(const stark_mg_fri* f) { return f ? f->subtree : nullptr; }

// decommit_fri (fri_commit.rs:168-179) for a proof whose layer 0 lives in leaf ranges: the index is drawn on rank 0 and
// broadcast; the owners of idx and idx + N/2 open their subtree; rank 0 appends the levels above the subtree roots, feeds
// the channel in the reference's order and opens layers >= 1 itself.  Collective: every rank calls it.
extern "C" int stark_mg_decommit_fri(stark_mg_fri* f, size_t num_queries, size_t max_index, stark_channel* ch) {
    MG_BEGIN
    STARK_REQUIRE(f && (ch || f->mg->rank != 0), "mg_decommit_fri: bad argument");
    stark_mg* mg = f->mg;
    MgGuard g(mg);
    const unsigned world = mg->world, rank = mg->rank;
    const size_t N = (size_t)1 << f->log_n, blk = N / world;
    std::vector<uint8_t> blob;
    (void)blk;
    for (size_t q = 0; q < num_queries; q++) {
        uint64_t idx = 0;
        if (rank == 0) STARK_REQUIRE(ch->ch.receive_random_int(0, max_index, true, &idx), "channel: receive before send");
        idx = bcast_u64(mg, idx);
        const size_t i0 = (size_t)idx % N;
        std::deque<LeafRangeSet> sets;
        std::vector<OpenItem> items;
        sets.push_back(LeafRangeSet{f->block, f->subtree, f->subtree_roots.data(), N});
        items.push_back(OpenItem{&sets[0], i0});                                               // layer 0: :156-163
        items.push_back(OpenItem{&sets[0], (i0 + N / 2) % N});
        sharded_layer_items(f->sharded, sets, (size_t)idx, items);                             // layers 1 .. S, same exchange
        open_leaf_ranges_batch(mg, items, rank == 0 ? &ch->ch : nullptr);
        if (rank != 0) continue;
        const size_t first = 1 + f->sharded.size();
        size_t len = 0;
        const uint64_t i64 = idx;
        int rc = stark_fri_open_layers(f->proof, first, &i64, 1, nullptr, 0, &len);
        if (rc != ST_OK) return rc;
        blob.resize(len);
        if (len) {
            rc = stark_fri_open_layers(f->proof, first, &i64, 1, blob.data(), blob.size(), &len);
            if (rc != ST_OK) return rc;
        }
        api_send_query_records(f->proof, blob.data(), ch->ch, (size_t)idx, first);
    }
    MG_END
}

// ======================================================================================= cfg5: the whole FibonacciSq prover
// stark101_prove (stark101.cu) with everything the north-star partitions spread over the group, and the
// same transcript byte for byte: the trace LDE through the four-step NTT, the commitments of f and of the composition
// polynomial in leaf ranges, the composition polynomial point-wise on each rank's own range (it needs f at i, i + blow,
// i + 2 blow: a halo of 2 blow values from the next rank).  The sequential trace recurrence is replicated; the FRI
// layers >= 1, the channel and their openings stay on rank 0.  `ch` is used on rank 0 only.  Collective.
extern "C" int stark_mg_stark101_prove(stark_mg* mg, uint64_t a1, unsigned log_trace, unsigned log_blowup, size_t num_queries, int transport,
                                       stark_channel* ch) {
    MG_BEGIN
    STARK_REQUIRE(mg && (ch || mg->rank != 0), "mg_stark101_prove: bad argument");
    MgGuard g(mg);
    stark_ctx* ctx = mg->ctx;
    const unsigned world = mg->world, rank = mg->rank;
    const unsigned log_n = log_trace + log_blowup;
    STARK_REQUIRE(log_trace >= 2 && log_n <= ctx->two_adicity && log_n <= 30, "stark101: trace/blowup sizes not supported by this field");
    const size_t N = (size_t)1 << log_n, blow = (size_t)1 << log_blowup, blk = N / world;
    STARK_REQUIRE(blk >= 2 * blow && blk % blow == 0, "mg_stark101_prove: every rank's range must hold at least 2 * blowup points");
    const uint64_t w = ctx->generator;
    Channel* chan = rank == 0 ? &ch->ch : nullptr;
    // ---- src/trace: sequential recurrence + interpolation, replicated (every rank needs all coefficients)
    stark_vec* f_coef = nullptr;
    uint64_t last_value = 0;
    int rc = stark101_trace_poly(ctx, a1, log_trace, &f_coef, &last_value);
    if (rc != ST_OK) return rc;
    std::unique_ptr<stark_vec> f_coef_guard(f_coef);
    // ---- trace LDE (four-step) and its commitment in leaf ranges
    DevBufPtr f_alias = fourstep_lde(mg, f_coef, log_n, w, transport);
    DevBufPtr f_block = make_buf(blk * 4, ctx->stream);                       // private copy: the transport's buffer is reused
    STARK_CUDA(cudaMemcpyAsync(f_block->p, f_alias->p, blk * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    std::vector<uint8_t> f_subs((size_t)world * 32), cp_subs((size_t)world * 32);
    uint8_t f_root[32], cp_root[32];
    auto f_sub = commit_leaf_range(mg, f_block, blk, f_root, f_subs.data());
    uint64_t alpha[3] = {0, 0, 0};
    if (rank == 0) {
        uint8_t stmt[STARK101_STATEMENT_BYTES];
        stark101_statement(ctx->modulus, ctx->generator, log_trace, log_blowup, num_queries, last_value, stmt);
        chan->send(stmt, sizeof stmt);
        api_send_root_bytes(*chan, f_root);
        for (int k = 0; k < 3; k++) STARK_REQUIRE(chan->receive_random_field_element(&alpha[k]), "channel: receive before send");
    }
    for (int k = 0; k < 3; k++) alpha[k] = bcast_u64(mg, alpha[k]);
    // ---- src/composition on the local range: block + halo from the next rank (wraps around)
    stark_vec f_vec; f_vec.ctx = ctx; f_vec.buf = f_block; f_vec.n = blk;
    stark_vec* cp_vec = nullptr;
    if (world == 1) {
        rc = stark101_composition_range(ctx, &f_vec, 0, N, alpha, last_value, log_trace, log_blowup, &cp_vec);
    } else {
        DevBuf heads((size_t)world * 2 * blow * 4, ctx->stream);
        STARK_NCCL(nccl().AllGather(f_block->p, heads.p, 2 * blow, NCCL_UINT32, mg->comm, ctx->stream));
        stark_vec ext; ext.ctx = ctx; ext.n = blk + 2 * blow;
        ext.buf = make_buf(ext.n * 4, ctx->stream);
        STARK_CUDA(cudaMemcpyAsync(ext.buf->p, f_block->p, blk * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        STARK_CUDA(cudaMemcpyAsync(ext.buf->as<uint32_t>() + blk, heads.as<uint32_t>() + (size_t)((rank + 1) % world) * 2 * blow, 2 * blow * 4,
                                   cudaMemcpyDeviceToDevice, ctx->stream));
        rc = stark101_composition_range(ctx, &ext, rank * blk, blk, alpha, last_value, log_trace, log_blowup, &cp_vec);
    }
    if (rc != ST_OK) return rc;
    std::unique_ptr<stark_vec> cp_guard(cp_vec);
    const bool shard = shard_wanted(mg, log_n);
    DevBufPtr cp_all;
    if (world > 1 && shard) cp_all = allgather_layer_begin(mg, cp_vec->buf, blk);          // runs under the hashing below
    auto cp_sub = commit_leaf_range(mg, cp_vec->buf, blk, cp_root, cp_subs.data());
    // ---- src/fri: layer 0 = the CP evaluations, its coefficients by interpolation (degree tracking).  With the large layers
    // >= 1 hashed in leaf ranges every rank holds layer 0 and replicates the folds (sharded_fri_layers); otherwise rank 0 alone
    std::unique_ptr<stark_fri> proof;
    std::vector<ShardedLayer> sharded;
    if (world == 1 || rank == 0 || shard) {
        DevBufPtr layer0;
        if (world == 1) layer0 = cp_vec->buf;
        else if (shard) { layer0 = cp_all; allgather_layer_end(mg); }
        else {
            Nccl& n = nccl();
            layer0 = make_buf(N * 4, ctx->stream);
            STARK_CUDA(cudaMemcpyAsync(layer0->p, cp_vec->buf->p, blk * 4, cudaMemcpyDeviceToDevice, ctx->stream));
            STARK_NCCL(n.GroupStart());
            for (unsigned p = 1; p < world; p++) STARK_NCCL(n.Recv(layer0->as<uint32_t>() + p * blk, blk, NCCL_UINT32, (int)p, mg->comm, ctx->stream));
            STARK_NCCL(n.GroupEnd());
        }
        stark_vec l0; l0.ctx = ctx; l0.buf = layer0; l0.n = N;
        stark_vec cc; cc.ctx = ctx; cc.n = N;
        cc.buf = api_interpolate_on_coset(ctx, layer0->as<uint32_t>(), log_n, w);
        stark_fri* pf = nullptr;
        rc = stark_fri_begin_external(ctx, &cc, log_n, w, &l0, cp_root, &pf);
        if (rc != ST_OK) return rc;
        proof.reset(pf);
        if (!shard) rc = api_fri_commit_loop(pf, ch);
        else {
            if (rank == 0) api_send_root_bytes(*chan, cp_root);
            sharded_fri_layers(mg, pf, chan, sharded);
            if (rank == 0) rc = api_fri_commit_loop_resume(pf, ch);
        }
        if (rc != ST_OK) return rc;
    } else {
        STARK_NCCL(nccl().Send(cp_vec->buf->p, blk, NCCL_UINT32, 0, mg->comm, ctx->stream));
        STARK_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    // ---- queries: f(x), f(gx), f(g^2 x), layer 0 of the FRI and its layers hashed in leaf ranges from their owners in one
    // exchange per query, the other layers on rank 0
    std::vector<uint8_t> blob;
    for (size_t q = 0; q < num_queries; q++) {
        uint64_t idx = 0;
        if (rank == 0) STARK_REQUIRE(chan->receive_random_int(0, N - 1 - 2 * blow, true, &idx), "channel: receive before send");
        idx = bcast_u64(mg, idx);
        std::deque<LeafRangeSet> sets;
        std::vector<OpenItem> items;
        sets.push_back(LeafRangeSet{&f_vec, f_sub.get(), f_subs.data(), N});
        sets.push_back(LeafRangeSet{cp_vec, cp_sub.get(), cp_subs.data(), N});
        for (size_t k = 0; k < 3; k++) items.push_back(OpenItem{&sets[0], (size_t)idx + k * blow});
        items.push_back(OpenItem{&sets[1], (size_t)idx % N});
        items.push_back(OpenItem{&sets[1], ((size_t)idx + N / 2) % N});
        sharded_layer_items(sharded, sets, (size_t)idx, items);
        open_leaf_ranges_batch(mg, items, chan);
        if (rank != 0) continue;
        const size_t first = 1 + sharded.size();
        size_t len = 0;
        const uint64_t i64 = idx;
        rc = stark_fri_open_layers(proof.get(), first, &i64, 1, nullptr, 0, &len);
        if (rc != ST_OK) return rc;
        blob.resize(len);
        if (len) {
            rc = stark_fri_open_layers(proof.get(), first, &i64, 1, blob.data(), blob.size(), &len);
            if (rc != ST_OK) return rc;
        }
        api_send_query_records(proof.get(), blob.data(), *chan, (size_t)idx, first);
    }
    MG_END
}
