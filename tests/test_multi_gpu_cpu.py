"""CPU tests of the multi-GPU host logic: column sharding and root gathering over a real
world_size-2 gloo group (compute injected from the oracle), subtree-root combination, and the index algebra
of the four-step NTT against the oracle."""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = 3221225473


def _mg():
    return importlib.import_module("stark-prover_b200.multi_gpu")


def test_shard_columns():
    mg = _mg()
    assert mg.shard_columns(8, 1, 4) == [1, 5]
    got = sorted(c for r in range(3) for c in mg.shard_columns(10, r, 3))
    assert got == list(range(10))


def test_combine_subtree_roots_matches_full_tree(orc):
    mg = _mg()
    vals = orc.synthetic_column(3, 1 << 10)
    full = orc.Tree(vals)
    for g in (1, 2, 4, 8):
        blk = len(vals) // g
        subs = [orc.Tree(vals[r * blk:(r + 1) * blk]).root() for r in range(g)]
        assert mg.combine_subtree_roots(subs) == full.root()
        # a path = path inside the owner's subtree + the top levels
        idx = 777
        owner = idx // blk
        local = orc.Tree(vals[owner * blk:(owner + 1) * blk]).path(idx - owner * blk)
        assert local + mg.top_path(subs, owner) == full.path(idx)


def test_four_step_index_algebra(orc):
    mg = _mg()
    log_n, offset = 8, 5
    w = orc.root_of_unity(log_n)
    coeffs = orc.synthetic_column(21, 1 << 5)
    want = orc.coset_evaluate(coeffs, log_n, offset, w, P)
    for world in (1, 2, 4):
        blocks = mg.four_step_reference(coeffs, log_n, offset, lambda a, lm, ww: orc.ntt(a, lm, ww, P), P, w, world)
        assert np.array_equal(np.concatenate(blocks), want), world
        a, b = mg.four_step_plan(log_n, world)
        loc = mg.four_step_scatter_input(coeffs, log_n, world - 1, world)
        assert loc.shape == ((1 << b) // world, 1 << a)


def _worker(rank, world, port, n_cols, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import pyoracle as orc
    mg = importlib.import_module("stark-prover_b200.multi_gpu")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        log_rows, log_blowup = 6, 3
        calls = []

        def commit(c):     # the GPU committer's contract, computed by the oracle here
            calls.append(c)
            col = orc.synthetic_column(100 + c, 1 << log_rows)
            coef = orc.coset_interpolate(col, log_rows, 1, orc.root_of_unity(log_rows), P)
            lde = orc.coset_evaluate(coef, log_rows + log_blowup, 5, orc.root_of_unity(log_rows + log_blowup), P)
            return orc.Tree(lde).root()

        roots = mg.commit_columns(n_cols, commit, rank, world)
        big = orc.synthetic_column(9, 256)
        mine = big[rank * 128:(rank + 1) * 128]
        sub = orc.Tree(mine)
        top, subs = mg.commit_leaf_ranges(sub.root, rank, world)

        class Block:                                   # the two calls _open_leaf_range makes on a device Vec / MerkleTree
            def download(self, off, n): return mine[off:off + n]
        class Sub:
            def get_authentication_path(self, i): return sub.path(i)
        opened = [mg._open_leaf_range(Block(), Sub(), subs, which, 128, rank) for which in (3, 131, 255)]
        q.put((rank, calls, [r.hex() for r in roots], top.hex(), [(e.hex(), pth.hex()) for e, pth in opened]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_commit_columns_gloo_world2(orc):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    n_cols = 5
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_cols, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in procs)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    (r0, calls0, roots0, top0, open0), (r1, calls1, roots1, top1, open1) = res
    big = orc.synthetic_column(9, 256)
    full = orc.Tree(big)
    assert open0 == open1                                     # every rank learns the same opening, whoever owns the leaf
    for (elem, pth), which in zip(open0, (3, 131, 255)):
        assert elem == int(big[which]).to_bytes(8, "big").hex() and pth == full.path(which).hex()
    assert calls0 == [0, 2, 4] and calls1 == [1, 3]          # no column is computed twice
    assert roots0 == roots1 and top0 == top1                 # every rank ends with the same commitment
    for c in range(n_cols):                                   # and it is the single-process answer
        col = orc.synthetic_column(100 + c, 64)
        coef = orc.coset_interpolate(col, 6, 1, orc.root_of_unity(6), P)
        lde = orc.coset_evaluate(coef, 9, 5, orc.root_of_unity(9), P)
        assert roots0[c] == orc.Tree(lde).root_hex()
    assert top0 == orc.Tree(orc.synthetic_column(9, 256)).root_hex()


def test_feed_layer_records_order(orc):
    """The host-side assembly of a query's messages (layer 0 from the owners, later layers from rank 0) reproduces
    decommit_fri_layers (fri_commit.rs:137-165), including the length-1 layer that falls through."""
    mg = _mg()
    for log_n, log_deg in ((8, 5), (5, 5)):          # the second case folds down to a one-point layer
        w = orc.root_of_unity(log_n)
        c = orc.synthetic_poly_exact_degree(7, 1 << log_deg)
        ch_a, ch_b = orc.Channel(P), orc.Channel(P)
        pr_a = orc.fri_commit_fast(c, log_n, 5, w, ch_a, P)
        pr_b = orc.fri_commit_fast(c, log_n, 5, w, ch_b, P)
        index = 77 % (1 << log_n)
        orc.decommit_fri_layers(index, pr_a, ch_a)
        # layer 0 by hand (what the owners send), then the blob format of stark_fri_open_layers for layers >= 1
        n0 = 1 << log_n
        for which in (index % n0, (index % n0 + n0 // 2) % n0):
            ch_b.send(int(pr_b.layer(0)[which]).to_bytes(8, "big"))
            ch_b.send(pr_b.tree(0).path(which))
        blob, lens = b"", []
        for k in range(1, pr_b.num_layers):
            lay, n = pr_b.layer(k), len(pr_b.layer(k))
            lens.append(n)
            for which in (index % n, (index % n + n // 2) % n):
                blob += int(lay[which]).to_bytes(8, "big") + pr_b.tree(k).path(which)
        mg.feed_layer_records(ch_b, blob, lens, index)
        assert ch_a.state == ch_b.state and ch_a.proof == ch_b.proof


# ---- the large FRI layers hashed in leaf ranges (csrc/multi.cu: sharded_fri_layers / open_leaf_ranges_batch) -----------------
class _OracleEngine:
    """The compute of the protocol, from the oracle."""
    def __init__(self, orc):
        self.orc = orc
    def evaluate(self, c, log_n, offset): return self.orc.coset_evaluate(c, log_n, offset, self.orc.root_of_unity(log_n), P)
    def fold(self, e, beta, offset, log_len): return self.orc.fri_fold_evals(e, beta, offset, self.orc.root_of_unity(log_len), P)
    def fold_poly(self, c, beta): return self.orc.poly_trim(self.orc.next_fri_polynomial(c, beta, P))
    def tree(self, values): return self.orc.Tree(values)


def _fri_worker(rank, world, port, q, log_n, log_deg, queries, min_len):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import pyoracle as orc
    mg = importlib.import_module("stark-prover_b200.multi_gpu")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        c = orc.synthetic_poly_exact_degree(11, 1 << log_deg)
        ch = orc.Channel(P) if rank == 0 else None
        ranged, tail = mg.fri_commit_leaf_range_layers(_OracleEngine(orc), c, log_n, 5, P, ch, rank, world, min_len)
        mg.decommit_fri_leaf_range_layers(ranged, tail, queries, (1 << log_n) - 1, ch, rank, world)
        q.put((rank, len(ranged), len(tail), ch.state if ch else None, [m.hex() for m in ch.proof] if ch else None))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
@pytest.mark.parametrize("world,log_n,log_deg,min_len", [(2, 9, 6, 16), (2, 7, 7, 4), (4, 10, 5, 64)])
def test_fri_leaf_range_layers_gloo(orc, world, log_n, log_deg, min_len):
    """world_size-2 and -4 gloo groups run the protocol of the sharded FRI layers (replicated folds, leaf-range hashing, root
    gathers, beta broadcast, one exchange per query); rank 0's transcript is the single-process oracle's, byte for byte."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() * 7 + world * 13 + log_n) % 2000
    queries = 3
    procs = [ctx.Process(target=_fri_worker, args=(r, world, port, q, log_n, log_deg, queries, min_len)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=150) for _ in procs)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    och = orc.Channel(P)
    c = orc.synthetic_poly_exact_degree(11, 1 << log_deg)
    opr = orc.fri_commit_fast(c, log_n, 5, orc.root_of_unity(log_n), och, P)
    orc.decommit_fri(queries, (1 << log_n) - 1, opr, och)
    _, n_ranged, n_tail, state, proof = res[0]
    assert n_ranged >= 2 and n_ranged + n_tail == opr.num_layers          # at least one layer >= 1 went through the ranges
    assert all(r[1] == n_ranged for r in res)                             # every rank took the same number of sharded layers
    assert state == och.state and proof == [m.hex() for m in och.proof]
