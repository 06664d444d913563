"""The host-side verifier: accepts proofs made by the CPU oracle (CPU test) and by the
GPU path (GPU test), rejects every single-byte corruption class."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

P = 3221225473


def _flat(msgs):
    return b"".join(len(m).to_bytes(4, "little") + m for m in msgs)


def _oracle_proof(orc, log_n, log_deg, q, seed=43):
    c = orc.synthetic_poly_exact_degree(seed, 1 << log_deg)
    ch = orc.Channel(P)
    pr = orc.fri_commit_fast(c, log_n, 5, orc.root_of_unity(log_n), ch, P)
    orc.decommit_fri(q, (1 << log_n) - 1, pr, ch)
    return ch.proof


@pytest.mark.parametrize("log_n,log_deg,q", [(10, 7, 3), (8, 8, 2), (6, 6, 2), (5, 0, 1)])
def test_verifier_accepts_oracle_proofs(sp, orc, log_n, log_deg, q):
    ok, why = sp.verify_fri(_flat(_oracle_proof(orc, log_n, log_deg, q)), log_n, 5, q, (1 << log_n) - 1)
    assert ok, why


def test_verifier_rejects_corruptions(sp, orc):
    log_n, q = 10, 3
    msgs = _oracle_proof(orc, log_n, 7, q)
    n_layers = 8
    first_query = 2 * n_layers                       # root0, (beta, root) x 7, final -> index at position 16
    cases = {"root": 0, "beta": 1, "final": first_query - 1, "index": first_query, "element": first_query + 1,
             "path": first_query + 2, "sibling element": first_query + 3, "deep path": len(msgs) - 1}
    for name, pos in cases.items():
        bad = [bytearray(m) for m in msgs]
        bad[pos][len(bad[pos]) // 2] ^= 0x01
        ok, why = sp.verify_fri(_flat([bytes(m) for m in bad]), log_n, 5, q, (1 << log_n) - 1)
        assert not ok and why, name
    ok, _ = sp.verify_fri(_flat(msgs[:-1]), log_n, 5, q, (1 << log_n) - 1)           # truncated
    assert not ok
    ok, _ = sp.verify_fri(_flat(msgs), log_n, 7, q, (1 << log_n) - 1)                # wrong coset offset: folds no longer match
    assert not ok
    ok, _ = sp.verify_fri(_flat(msgs), log_n, 5, q, (1 << log_n) - 2)                # wrong query range: indices no longer match
    assert not ok


def test_merkle_validate(sp, orc):
    vals = orc.synthetic_column(1, 100)
    t = orc.Tree(vals)
    for idx in (0, 37, 99):
        assert sp.merkle_validate(t.root(), 100, idx, int(vals[idx]), t.path(idx))
        assert not sp.merkle_validate(t.root(), 100, idx, (int(vals[idx]) + 1) % P, t.path(idx))
        assert not sp.merkle_validate(t.root(), 100, idx ^ 1, int(vals[idx]), t.path(idx))
    assert not sp.merkle_validate(t.root(), 100, 5, int(vals[5]), t.path(5)[:-32])


@pytest.mark.gpu
@pytest.mark.parametrize("log_n,log_deg,q", [(12, 9, 4), (20, 17, 8)])
def test_verifier_accepts_gpu_proofs(sp, orc, ctx, log_n, log_deg, q):
    c = orc.synthetic_poly_exact_degree(43, 1 << log_deg)
    ch = sp.Channel(P)
    pr = sp.fri_commit(ctx, c, sp.CosetFri(ctx, 5, log_n), ch)
    sp.decommit_fri(q, (1 << log_n) - 1, pr, ch)
    ok, why = sp.verify_fri(ch.proof_flat(), log_n, 5, q, (1 << log_n) - 1)
    assert ok, why
    t = pr.tree(0)
    assert sp.merkle_validate(t.root_bytes(), 1 << log_n, 5, int(pr.layer(0, 5, 1)[0]), t.get_authentication_path(5))


def test_stark101_verifier_on_oracle_transcript(sp, orc):
    """End-to-end soundness of the build-defined prover: the verifier links trace openings, composition polynomial
    and FRI.  CPU only (oracle transcript)."""
    ch = orc.Channel(P)
    orc.stark101_prove(ch, literal=False)
    claimed = int(orc.fibsq_trace(3141592, 1023)[1022])
    ok, why = sp.stark101_verify(ch.proof_flat(), claimed)
    assert ok, why
    ok, why = sp.stark101_verify(ch.proof_flat(), claimed + 1)          # a false claim about a_1022
    assert not ok and "composition" in why
    msgs = ch.proof
    for pos in (0, 1, 4, 20, 21, 22, 23, len(msgs) - 1):                 # trace root, alpha, CP root, final..., openings
        bad = [bytearray(m) for m in msgs]
        bad[pos][len(bad[pos]) // 2] ^= 0x04
        ok, why = sp.stark101_verify(b"".join(len(m).to_bytes(4, "little") + bytes(m) for m in bad), claimed)
        assert not ok and why, pos


@pytest.mark.gpu
@pytest.mark.parametrize("log_trace,a1", [(10, 3141592), (14, 7)])
def test_stark101_verifier_accepts_gpu_proof(sp, orc, ctx, log_trace, a1):
    ch = sp.Channel(P)
    sp.stark101_prove(ctx, ch, a1, log_trace, 3, 3)
    claimed = int(orc.fibsq_trace(a1, (1 << log_trace) - 1)[(1 << log_trace) - 2])
    ok, why = sp.stark101_verify(ch.proof_flat(), claimed, log_trace, 3, 3)
    assert ok, why
    ok, _ = sp.stark101_verify(ch.proof_flat(), (claimed + 5) % P, log_trace, 3, 3)
    assert not ok


def test_fri_verifier_checks_the_length_one_layer_element(sp, orc):
    """A FRI that folds down to a one-point layer sends that point once more before its openings (fri_commit.rs:147-149);
    the verifier must compare it with the opened value -- in the LAST query nothing after it depends on the channel
    state, so a flipped bit there was accepted (found by tools/soak_host.py)."""
    log_n, off, q = 3, 1365814300, 2
    c = orc.synthetic_poly_exact_degree(5, 7)
    ch = orc.Channel(P)
    pr = orc.fri_commit_fast(c, log_n, off, orc.root_of_unity(log_n), ch, P)
    orc.decommit_fri(q, (1 << log_n) - 1, pr, ch)
    assert len(pr.layer(pr.num_layers - 1)) == 1
    msgs = ch.proof
    flat = lambda ms: b"".join(len(m).to_bytes(4, "little") + bytes(m) for m in ms)
    assert sp.verify_fri(flat(msgs), log_n, off, q, (1 << log_n) - 1)[0]
    lone = len(msgs) - 5                      # ... lone element, elem, path, sibling elem, sibling path
    assert len(msgs[lone]) == 8 and msgs[lone] == msgs[lone + 1]
    bad = [bytearray(m) for m in msgs]
    bad[lone][7] ^= 1
    ok, why = sp.verify_fri(flat(bad), log_n, off, q, (1 << log_n) - 1)
    assert not ok and "length-1" in why


def test_host_soak_short():
    """A few seconds of the randomised host-side differential run (channel op sequences, Merkle validation, verifiers
    with bit flips); the long form is `python tools/soak_host.py 60`."""
    import subprocess, sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "soak_host.py"), "4", "99"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
