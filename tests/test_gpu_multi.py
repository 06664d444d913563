"""GPU tests of the multi-GPU building blocks on ONE device: batched NTTs, the four-step flow with all ranks
emulated in one process (exchanges in memory), leaf-range subtree roots, the column committer."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
P = 3221225473


def _mg():
    return importlib.import_module("stark-prover_b200.multi_gpu")


@pytest.mark.parametrize("log_m,batch", [(1, 8), (5, 3), (9, 64), (10, 7), (13, 16), (16, 2)])
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
