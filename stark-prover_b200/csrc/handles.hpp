// handles.hpp — the opaque objects behind the C ABI.
#pragma once
#include "host_channel.hpp"
#include "kernels.hpp"

// (the context reference is the FIRST member of every handle: members die in reverse order, so the buffers go back to the
// context's stream before the reference that may tear the context down is dropped)
struct stark_vec {
    starkb200::CtxRef ctx;
    starkb200::DevBufPtr buf;        // n canonical u32 values
    size_t n = 0;
};

// MerkleTree<M> (reference src/merkle/mod.rs:5-7): the leaf values it was built over (shared with the
// FRI layer or vector that owns them) plus levels 1..depth of digests.
struct stark_tree {
    starkb200::CtxRef ctx;
    starkb200::DevBufPtr leaves;
    starkb200::TreeShape shape;
    starkb200::DevBuf nodes;
    uint32_t root_words[8] = {0};
    bool external = false;           // levels live elsewhere (leaf ranges hashed on other GPUs): only leaves + root here
};

// FRIProof (reference src/fri/fri_commit.rs:9-13): trees[k]->leaves are fri_layers[k], trees[k] is
// fri_merkles[k]; `coeffs` tracks the folded polynomial so that its exact degree and the final
// constant are the reference's.
struct stark_fri {
    starkb200::CtxRef ctx;
    unsigned log_n = 0;
    uint64_t offset0 = 1;
    unsigned cur_log = 0;
    uint64_t cur_offset = 1;
    uint64_t half_over_offset = 0;   // 1 / (2 * cur_offset) mod p, squared-and-doubled along with cur_offset (0 = not formed yet)
    starkb200::DevBufPtr coeffs;
    size_t coeff_len = 0;            // degree + 1 (0 = zero polynomial)
    std::vector<std::unique_ptr<stark_tree>> trees;
    // FRIProof.fri_layers by value (fri_commit.rs:117-121): every layer is also widened to u64 and copied to host memory on
    // the context's copy stream while the main stream hashes the following layers (stark_fri_begin_to_host).  Declared
    // after `trees`: it dies first and waits for its copies, so no layer is released under a copy in flight.
    struct LayerSink {
        cudaStream_t stream = nullptr;
        uint64_t* host = nullptr;
        size_t cap = 0, off = 0;
        std::vector<size_t> offs;        // element offset of layer k in `host`
        starkb200::DevBuf stage;         // u64 staging for one layer (the largest); reused in copy-stream order
        starkb200::DevBuf fold_tmp;      // a folded layer computed on the copy stream ahead of the fused fold-and-hash launch
        ~LayerSink() { if (stream) cudaStreamSynchronize(stream); }
    };
    std::unique_ptr<LayerSink> sink;
};

struct stark_channel {
    starkb200::Channel ch;
    explicit stark_channel(uint64_t m) : ch(m) {}
};
