"""GPU tests of the multi-GPU building blocks on ONE device: batched NTTs, the four-step flow with all ranks
emulated in one process (exchanges in memory), leaf-range subtree roots, the column committer."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
P = 3221225473


def _mg():
    return importlib.import_module("stark-prover_b200.multi_gpu")


# (5, 37), (3, 37), (9, 38): group counts that are not a multiple of the 32-column tile (found by tools/soak.py)
@pytest.mark.parametrize("log_m,batch", [(1, 8), (5, 3), (9, 64), (10, 7), (13, 16), (16, 2), (5, 37), (3, 37), (9, 38), (1, 99), (11, 5), (19, 3), (12, 33)])
def test_ntt_batch(sp, orc, ctx, log_m, batch):
    m = 1 << log_m
    a = orc.synthetic_column(log_m * 10 + batch, m * batch)
    w = orc.root_of_unity(log_m)
    v = ctx.upload(a)
    ctx.ntt_batch_dev(v, log_m)
    got = v.download().reshape(batch, m)
    for b in range(batch):
        assert np.array_equal(got[b], orc.ntt(a[b * m:(b + 1) * m], log_m, w, P)), b
    ctx.ntt_batch_dev(v, log_m, inverse=True)
    assert np.array_equal(v.download(), a)


@pytest.mark.parametrize("log_n,log_deg,world", [(10, 7, 1), (10, 7, 2), (14, 11, 4), (14, 14, 8), (18, 15, 8), (20, 17, 2)])
def test_four_step_emulated_ranks(sp, orc, ctx, log_n, log_deg, world):
    mg = _mg()
    coeffs = orc.synthetic_column(log_n + world, 1 << log_deg)
    want = orc.coset_evaluate(coeffs, log_n, 5, orc.root_of_unity(log_n), P)
    blocks = mg.four_step_lde_emulated(sp, ctx, coeffs, log_n, 5, world)
    got = np.concatenate([b.download() for b in blocks])
    assert np.array_equal(got, want)
    # leaf ranges are exact subtrees: per-rank roots combine to the root of the whole column
    subs = [sp.MerkleTree.new(ctx, b).root_bytes() for b in blocks]
    assert mg.combine_subtree_roots(subs) == orc.merkle_root_only(want)
    # and equals the single-GPU path
    assert sp.MerkleTree.new(ctx, ctx.coset_evaluate_dev(ctx.upload(coeffs), log_n, 5)).root_bytes() == mg.combine_subtree_roots(subs)


def test_commit_columns_single_rank(sp, orc, ctx):
    mg = _mg()
    log_rows, log_blowup, n_cols = 12, 3, 5
    cols = [orc.synthetic_column(100 + c, 1 << log_rows) for c in range(n_cols)]
    keep = {}
    roots = mg.commit_columns(n_cols, mg.gpu_column_committer(sp, ctx, lambda c: cols[c], log_blowup, 1, 5, keep), 0, 1)
    for c in range(n_cols):
        coef = orc.coset_interpolate(cols[c], log_rows, 1, orc.root_of_unity(log_rows), P)
        lde = orc.coset_evaluate(coef, log_rows + log_blowup, 5, orc.root_of_unity(log_rows + log_blowup), P)
        assert roots[c] == orc.merkle_root_only(lde)
        assert np.array_equal(keep[c][0].download(0, 64), lde[:64])       # sampled rows


def test_fri_commit_multi_degenerate_world1(sp, orc, ctx):
    """The sharded-layer-0 protocol with one rank: adopts the four-step layer 0, transcript == fri_commit's."""
    mg = _mg()
    log_n = 14
    c = orc.synthetic_poly_exact_degree(4, 1 << 11)
    ch, ch1, och = sp.Channel(P), sp.Channel(P), orc.Channel(P)
    mp = mg.fri_commit_multi(sp, ctx, c, log_n, 5, ch, 0, 1)
    mg.decommit_fri_multi(sp, mp, 3, (1 << log_n) - 1, ch, 0, 1)
    pr = sp.fri_commit(ctx, c, sp.CosetFri(ctx, 5, log_n), ch1)
    sp.decommit_fri(3, (1 << log_n) - 1, pr, ch1)
    opr = orc.fri_commit_fast(c, log_n, 5, orc.root_of_unity(log_n), och, P)
    orc.decommit_fri(3, (1 << log_n) - 1, opr, och)
    assert ch.state == ch1.state == och.state and ch.proof == ch1.proof == och.proof
    with pytest.raises(sp.StarkError):
        mp.proof.tree(0).get_authentication_path(0)          # layer 0's levels are not held by the proof object


@pytest.mark.parametrize("log_trace,log_blowup,a1,q", [(10, 3, 3141592, 3), (13, 3, 77, 2), (9, 4, 5, 2)])
def test_stark101_prove_multi_degenerate_world1(sp, orc, ctx, log_trace, log_blowup, a1, q):
    """The sharded prover with one rank: four-step LDE, leaf-range commitment, range composition, adopted layer 0 --
    the transcript is stark101_prove's (and the oracle's), and the verifier accepts it."""
    mg = _mg()
    ch, ch1, och = sp.Channel(P), sp.Channel(P), orc.Channel(P)
    mg.stark101_prove_multi(sp, ctx, ch, a1, log_trace, log_blowup, q, 0, 1)
    sp.stark101_prove(ctx, ch1, a1, log_trace, log_blowup, q)
    orc.stark101_prove(och, a1, log_trace, log_blowup, sp.G_DEFAULT, q, literal=False)
    assert ch.state == ch1.state == och.state and ch.proof == ch1.proof == och.proof
    claimed = int(orc.fibsq_trace(a1, (1 << log_trace) - 1)[(1 << log_trace) - 2])
    ok, why = sp.stark101_verify(ch.proof_flat(), claimed, log_trace, log_blowup, q)
    assert ok, why


def test_stark101_composition_range_matches_whole(sp, orc, ctx):
    """CP on ranges with a halo == CP on the whole coset (the kernel the ranks of stark101_prove_multi run)."""
    log_trace, log_blowup = 9, 3
    n, blow = 1 << (log_trace + log_blowup), 1 << log_blowup
    f_coef, last = sp.stark101_trace_poly(ctx, 3141592, log_trace)
    f = ctx.coset_evaluate(f_coef.download(), log_trace + log_blowup, sp.G_DEFAULT)
    alpha = [11, 22, 33]
    whole = sp.stark101_composition_range(ctx, ctx.upload(f), 0, n, alpha, last, log_trace, log_blowup).download()
    for world in (2, 8):
        blk = n // world
        for r in range(world):
            ext = np.concatenate([f[r * blk:(r + 1) * blk], np.roll(f, -((r + 1) * blk) % n)[: 2 * blow]])
            part = sp.stark101_composition_range(ctx, ctx.upload(ext), r * blk, blk, alpha, last, log_trace, log_blowup).download()
            assert np.array_equal(part, whole[r * blk:(r + 1) * blk]), (world, r)


@pytest.mark.parametrize("log_n,log_deg,world", [(10, 7, 1), (12, 9, 2), (14, 11, 4), (16, 13, 8), (16, 16, 2), (20, 17, 8), (21, 18, 4)])
def test_four_step_peer_memory_emulated(sp, orc, ctx, log_n, log_deg, world):
    """Fused twiddle+row-scatter and transpose+scatter kernels with every rank emulated on one GPU."""
    mg = _mg()
    coeffs = orc.synthetic_column(log_n * 3 + world, 1 << log_deg)
    want = orc.coset_evaluate(coeffs, log_n, 5, orc.root_of_unity(log_n), P)
    blocks = mg.four_step_p2p_emulated(sp, ctx, coeffs, log_n, 5, world)
    got = np.concatenate([b.download() for b in blocks])
    assert np.array_equal(got, want)


def test_four_step_peer_memory_rejects_tiny(sp, orc, ctx):
    mg = _mg()
    with pytest.raises(sp.StarkError):
        mg.four_step_p2p_emulated(sp, ctx, orc.synthetic_column(1, 64), 10, 5, 4)     # < 32 rows per rank


@pytest.mark.parametrize("log_n,log_deg,world", [(12, 9, 2), (16, 13, 8)])
def test_four_step_peer_memory_host_barrier_form(sp, orc, ctx, log_n, log_deg, world):
    """The same kernels without the epoch flags (every phase call returns after its stores: the two-barrier form)."""
    mg = _mg()
    coeffs = orc.synthetic_column(log_n * 5 + world, 1 << log_deg)
    want = orc.coset_evaluate(coeffs, log_n, 5, orc.root_of_unity(log_n), P)
    blocks = mg.four_step_p2p_emulated(sp, ctx, coeffs, log_n, 5, world, flags=False)
    assert np.array_equal(np.concatenate([b.download() for b in blocks]), want)


# ---- the C-level group (csrc/multi.cu) with one rank: every entry point, no NCCL involved --------------------------
def test_mg_commit_columns_world1(sp, orc, ctx):
    g = sp.MultiGpu(ctx, 0, 1)
    log_rows, log_blowup, n_cols = 12, 3, 5
    cols = [orc.synthetic_column(100 + c, 1 << log_rows) for c in range(n_cols)]
    roots, kept = g.commit_columns(cols, n_cols, log_rows, 1, log_blowup, 5, keep=True)
    for c in range(n_cols):
        coef = orc.coset_interpolate(cols[c], log_rows, 1, orc.root_of_unity(log_rows), P)
        lde = orc.coset_evaluate(coef, log_rows + log_blowup, 5, orc.root_of_unity(log_rows + log_blowup), P)
        assert roots[c] == orc.merkle_root_only(lde), c
        vec, tree = kept[c]
        assert np.array_equal(vec.download(), lde) and tree.root_bytes() == roots[c]
        assert orc.merkle_verify(roots[c], len(lde), 77, int(lde[77]), tree.get_authentication_path(77))
    assert g.commit_columns(cols, n_cols, log_rows, 1, log_blowup, 5) == roots      # without retained trees
    g.close()


@pytest.mark.parametrize("transport", [0, 1])
@pytest.mark.parametrize("log_n,log_deg", [(10, 7), (15, 15), (20, 17)])
def test_mg_fourstep_and_leaf_ranges_world1(sp, orc, ctx, log_n, log_deg, transport):
    g = sp.MultiGpu(ctx, 0, 1)
    coeffs = orc.synthetic_column(log_n + 7 * transport, 1 << log_deg)
    want = orc.coset_evaluate(coeffs, log_n, 5, orc.root_of_unity(log_n), P)
    cv = ctx.upload(coeffs)
    for _ in range(2):                                   # twice: epochs / staging buffers are reused
        blk = g.fourstep_lde(cv, log_n, 5, transport)
        assert np.array_equal(blk.download(), want)
    tree, root, subs = g.commit_leaf_ranges(blk)
    assert root == orc.merkle_root_only(want) and subs == [root] and tree.root_bytes() == root
    g.close()


@pytest.mark.parametrize("transport", [0, 1])
def test_mg_fri_commit_world1(sp, orc, ctx, transport):
    """stark_mg_fri_commit / stark_mg_decommit_fri with one rank: the transcript is fri_commit's and the oracle's."""
    g = sp.MultiGpu(ctx, 0, 1)
    log_n, q = 14, 3
    c = orc.synthetic_poly_exact_degree(4, 1 << 11)
    ch, ch1, och = sp.Channel(P), sp.Channel(P), orc.Channel(P)
    f = g.fri_commit(ctx.upload(c), log_n, 5, ch, transport)
    g.decommit_fri(f, q, (1 << log_n) - 1, ch)
    pr = sp.fri_commit(ctx, c, sp.CosetFri(ctx, 5, log_n), ch1)
    sp.decommit_fri(q, (1 << log_n) - 1, pr, ch1)
    opr = orc.fri_commit_fast(c, log_n, 5, orc.root_of_unity(log_n), och, P)
    orc.decommit_fri(q, (1 << log_n) - 1, opr, och)
    assert ch.state == ch1.state == och.state and ch.proof == ch1.proof == och.proof
    assert f.proof.num_layers == pr.num_layers
    ok, why = sp.verify_fri(ch.proof_flat(), log_n, 5, q, (1 << log_n) - 1, 11)
    assert ok, why
    f.free(); g.close()


@pytest.fixture
def forced_sharding(monkeypatch):
    """One rank takes the path that hashes the large FRI layers >= 1 in leaf ranges (on N > 1 GPUs it is the default)."""
    monkeypatch.setenv("STARK_MG_FRI_SHARD_FORCE", "1")
    monkeypatch.setenv("STARK_MG_FRI_SHARD_MIN_LOG", "8")


@pytest.mark.parametrize("log_n,log_deg,q", [(14, 11, 3), (16, 13, 5), (12, 12, 2), (10, 3, 2)])
def test_mg_fri_commit_sharded_layers_world1(sp, orc, ctx, forced_sharding, log_n, log_deg, q):
    """stark_mg_fri_commit with the FRI layers >= 1 hashed in leaf ranges too (replicated folds + stand-alone coefficient
    fold, trees adopted from their combined roots, every leaf-range opening of a query in one exchange): with one rank the
    transcript is fri_commit's and the oracle's, byte for byte, and the verifier accepts it."""
    g = sp.MultiGpu(ctx, 0, 1)
    c = orc.synthetic_poly_exact_degree(40 + log_n, 1 << log_deg)
    ch, ch1, och = sp.Channel(P), sp.Channel(P), orc.Channel(P)
    f = g.fri_commit(ctx.upload(c), log_n, 5, ch, 1)
    g.decommit_fri(f, q, (1 << log_n) - 1, ch)
    pr = sp.fri_commit(ctx, c, sp.CosetFri(ctx, 5, log_n), ch1)
    sp.decommit_fri(q, (1 << log_n) - 1, pr, ch1)
    opr = orc.fri_commit_fast(c, log_n, 5, orc.root_of_unity(log_n), och, P)
    orc.decommit_fri(q, (1 << log_n) - 1, opr, och)
    assert ch.state == ch1.state == och.state and ch.proof == ch1.proof == och.proof
    mp = f.proof
    assert mp.num_layers == pr.num_layers
    for k in range(mp.num_layers):
        assert mp.tree(k).root() == pr.tree(k).root() and np.array_equal(mp.layer(k), pr.layer(k)), f"layer {k}"
    if log_n >= 10 and log_deg >= 8:
        with pytest.raises(sp.StarkError):
            mp.tree(1).get_authentication_path(0)        # layer 1's levels live in the leaf-range subtree, not in the proof object
    ok, why = sp.verify_fri(ch.proof_flat(), log_n, 5, q, (1 << log_n) - 1, log_deg)
    assert ok, why
    f.free(); g.close()


def test_mg_fri_sharded_sparse_polynomial_world1(sp, orc, ctx, forced_sharding):
    """The loop condition of the sharded layers is the exact degree (fri_commit.rs:89), tracked by the stand-alone
    coefficient fold: a sparse polynomial (powers of x^4 only, trailing zeros trimmed by Polynomial::new) takes the
    reference's number of layers."""
    log_n = 12
    c = np.zeros(1 << 9, dtype=np.uint64)
    c[::4] = orc.synthetic_column(3, 1 << 7)
    c[-4] = 0                                         # the top power present is x^500
    g = sp.MultiGpu(ctx, 0, 1)
    ch, och = sp.Channel(P), orc.Channel(P)
    f = g.fri_commit(ctx.upload(c), log_n, 5, ch, 1)
    g.decommit_fri(f, 2, (1 << log_n) - 1, ch)
    opr = orc.fri_commit_fast(c, log_n, 5, orc.root_of_unity(log_n), och, P)
    orc.decommit_fri(2, (1 << log_n) - 1, opr, och)
    assert f.proof.num_layers == opr.num_layers
    assert ch.state == och.state and ch.proof == och.proof
    f.free(); g.close()


@pytest.mark.parametrize("log_trace,log_blowup,a1,q", [(8, 3, 3141592, 3), (13, 3, 99, 4)])
def test_mg_stark101_prove_sharded_layers_world1(sp, orc, ctx, forced_sharding, log_trace, log_blowup, a1, q):
    g = sp.MultiGpu(ctx, 0, 1)
    ch, och = sp.Channel(P), orc.Channel(P)
    g.stark101_prove(ch, a1, log_trace, log_blowup, q, 1)
    orc.stark101_prove(och, a1, log_trace, log_blowup, sp.G_DEFAULT, q, literal=False)
    assert ch.state == och.state and ch.proof == och.proof
    claimed = int(orc.fibsq_trace(a1, (1 << log_trace) - 1)[(1 << log_trace) - 2])
    ok, why = sp.stark101_verify(ch.proof_flat(), claimed, log_trace, log_blowup, q)
    assert ok, why
    g.close()


@pytest.mark.parametrize("log_trace,log_blowup,a1,q,transport", [(8, 3, 3141592, 3, 1), (12, 3, 99, 4, 0), (15, 3, 3141592, 3, 1)])
def test_mg_stark101_prove_world1(sp, orc, ctx, log_trace, log_blowup, a1, q, transport):
    """The C-level sharded prover with one rank: stark101_prove's transcript (and the oracle's); the verifier accepts it."""
    g = sp.MultiGpu(ctx, 0, 1)
    ch, ch1, och = sp.Channel(P), sp.Channel(P), orc.Channel(P)
    g.stark101_prove(ch, a1, log_trace, log_blowup, q, transport)
    sp.stark101_prove(ctx, ch1, a1, log_trace, log_blowup, q)
    orc.stark101_prove(och, a1, log_trace, log_blowup, sp.G_DEFAULT, q, literal=False)
    assert ch.state == ch1.state == och.state and ch.proof == ch1.proof == och.proof
    claimed = int(orc.fibsq_trace(a1, (1 << log_trace) - 1)[(1 << log_trace) - 2])
    ok, why = sp.stark101_verify(ch.proof_flat(), claimed, log_trace, log_blowup, q)
    assert ok, why
    g.close()


# ---- full-size parity for the sharded configs on one GPU (cfg4 / cfg5 were never oracle-checked at size) ----
def test_cfg4_one_column_full_size(sp, orc, ctx):
    """One cfg4 column at BASELINE size: 2^22 rows -> coset LDE 2^25 -> root == the oracle's, 64 sampled rows equal."""
    log_rows, c = 22, 17
    col = orc.synthetic_column(100 + c, 1 << log_rows)
    g = sp.MultiGpu(ctx, 0, 1)
    roots, kept = g.commit_columns({0: col}, 1, log_rows, 1, 3, 5, keep=True)
    coef = orc.coset_interpolate(col, log_rows, 1, orc.root_of_unity(log_rows), P)
    lde = orc.coset_evaluate(coef, log_rows + 3, 5, orc.root_of_unity(log_rows + 3), P)
    assert roots[0] == orc.merkle_root_only(lde)
    vec, tree = kept[0]
    rng = np.random.default_rng(4)
    for i in rng.integers(0, 1 << 25, 64):
        assert int(vec.download(int(i), 1)[0]) == int(lde[int(i)])
    assert np.array_equal(vec.download(0, 1 << 16), lde[: 1 << 16])
    # the committed golden of the full 64-column run holds the same root for this column
    import json, os
    gpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sharded.json")
    gold = json.load(open(gpath)).get("cfg4")
    if gold:
        assert gold["roots"][c] == roots[0].hex()
    vec.free(); tree.free(); g.close()


@pytest.mark.parametrize("kind", ["nccl_flow", "peer_memory"])
def test_four_step_eight_emulated_ranks_2e24(sp, orc, ctx, kind):
    """8 emulated ranks at a 2^24 domain, both exchange flavours, against the oracle's root and sampled blocks."""
    mg = _mg()
    log_n, world = 24, 8
    coeffs = orc.synthetic_poly_exact_degree(43, 1 << (log_n - 3))
    want = orc.coset_evaluate(coeffs, log_n, 5, orc.root_of_unity(log_n), P)
    cv = ctx.upload(coeffs)
    blocks = mg.four_step_lde_emulated(sp, ctx, cv, log_n, 5, world) if kind == "nccl_flow" else mg.four_step_p2p_emulated(sp, ctx, cv, log_n, 5, world)
    blk = (1 << log_n) // world
    subs = []
    for r, b in enumerate(blocks):
        assert np.array_equal(b.download(0, 4096), want[r * blk:r * blk + 4096]), r
        assert np.array_equal(b.download(blk - 4096, 4096), want[(r + 1) * blk - 4096:(r + 1) * blk]), r
        subs.append(sp.MerkleTree.new(ctx, b).root_bytes())
    assert mg.combine_subtree_roots(subs) == orc.merkle_root_only(want)
    for b in blocks:
        b.free()
