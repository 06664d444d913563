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
    flat = _flat(_oracle_proof(orc, log_n, log_deg, q))
    ok, why = sp.verify_fri(flat, log_n, 5, q, (1 << log_n) - 1, log_deg)
    assert ok, why
    if log_deg >= 1:                                  # the same proof against a tighter degree claim: one fold too many
        ok, why = sp.verify_fri(flat, log_n, 5, q, (1 << log_n) - 1, log_deg - 1)
        assert not ok and "degree bound" in why


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
        ok, why = sp.verify_fri(_flat([bytes(m) for m in bad]), log_n, 5, q, (1 << log_n) - 1, 7)
        assert not ok and why, name
    ok, _ = sp.verify_fri(_flat(msgs[:-1]), log_n, 5, q, (1 << log_n) - 1, 7)           # truncated
    assert not ok
    ok, _ = sp.verify_fri(_flat(msgs), log_n, 7, q, (1 << log_n) - 1, 7)                # wrong coset offset: folds no longer match
    assert not ok
    ok, _ = sp.verify_fri(_flat(msgs), log_n, 5, q, (1 << log_n) - 2, 7)                # wrong query range: indices no longer match
    assert not ok


def _forged_fri_transcript(orc, layer0, log_n, offset, q, folds):
    """What a cheating prover can send for an ARBITRARY function on the domain: honest evaluation-space folds (so
    every fold check holds), honest trees, and as many layers as it likes (here `folds`), ending with the first value of
    the last layer as the 'final constant'.  With folds == log_n the last layer has one point and equals it trivially."""
    ch = orc.Channel(P)
    layers, trees = [np.array(layer0, dtype=np.uint64)], []
    off, w = offset, orc.root_of_unity(log_n)
    trees.append(orc.Tree(layers[0]))
    ch.send(trees[0].root_hex().encode())
    for _ in range(folds):
        beta = ch.receive_random_field_element()
        layers.append(orc.fri_fold_evals(layers[-1], beta, off, w, P))
        off, w = off * off % P, w * w % P
        trees.append(orc.Tree(layers[-1]))
        ch.send(trees[-1].root_hex().encode())
    ch.send(int(layers[-1][0]).to_bytes(8, "big"))
    for _ in range(q):
        idx = ch.receive_random_int(0, (1 << log_n) - 1, True)
        for lay, t in zip(layers, trees):
            n = len(lay)
            if n == 1:
                ch.send(int(lay[0]).to_bytes(8, "big"))
            for which in (idx % n, (idx % n + n // 2) % n):
                ch.send(int(lay[which]).to_bytes(8, "big"))
                ch.send(t.path(which))
    return ch.proof


def test_verifier_enforces_the_degree_bound(sp, orc):
    """the layer count is the prover's choice, so without a verifier-side degree bound a random
    function -- degree ~ n-1 -- passed every check by folding down to a one-point layer.  It is accepted ONLY under the
    vacuous bound (log_degree_bound == log_n, i.e. 'any function') and rejected under every real one."""
    log_n, q = 6, 8
    rnd = orc.synthetic_column(777, 1 << log_n)                      # not low degree
    msgs = _forged_fri_transcript(orc, rnd, log_n, 5, q, folds=log_n)
    ok, _ = sp.verify_fri(_flat(msgs), log_n, 5, q, (1 << log_n) - 1, log_n)
    assert ok                                                         # consistent folds of *some* function of degree < n
    for bound in range(log_n):
        ok, why = sp.verify_fri(_flat(msgs), log_n, 5, q, (1 << log_n) - 1, bound)
        assert not ok and "degree bound" in why, bound
    # the same cheat with exactly as many folds as the bound allows: the last layer has 2^(log_n - bound) points and is not
    # constant, so the queries catch it (16 queries, blow-up 8: escape probability ~ 2^-40 for this seed it does not)
    q2, bound = 16, log_n - 3
    msgs = _forged_fri_transcript(orc, rnd, log_n, 5, q2, folds=bound)
    ok, why = sp.verify_fri(_flat(msgs), log_n, 5, q2, (1 << log_n) - 1, bound)
    assert not ok and "final constant" in why
    ok, why = sp.verify_fri(_flat(msgs), log_n, 5, q2, (1 << log_n) - 1, log_n + 1)     # a bound larger than the domain is an error
    assert not ok


def _forged_stark_transcript(sp, orc, log_t, log_b, q, claimed, fe, cp_of, folds):
    """A cheating FibonacciSq prover: arbitrary 'trace LDE' `fe`, layer 0 of the FRI = cp_of(alphas) (an array on the
    coset), honest evaluation-space folds, `folds` layers after layer 0."""
    log_n = log_t + log_b
    N, blow = 1 << log_n, 1 << log_b
    ft = orc.Tree(fe)
    ch = orc.Channel(P)
    ch.send(sp.stark101_statement(P, 5, log_t, log_b, q, claimed))
    ch.send(ft.root_hex().encode())
    al = [ch.receive_random_field_element() for _ in range(3)]
    cp = cp_of(al)
    layers, trees, off, om = [cp], [orc.Tree(cp)], 5, orc.root_of_unity(log_n)
    ch.send(trees[0].root_hex().encode())
    for _ in range(folds):
        beta = ch.receive_random_field_element()
        layers.append(orc.fri_fold_evals(layers[-1], beta, off, om, P))
        off, om = off * off % P, om * om % P
        trees.append(orc.Tree(layers[-1]))
        ch.send(trees[-1].root_hex().encode())
    ch.send(int(layers[-1][0]).to_bytes(8, "big"))
    for _ in range(q):
        idx = ch.receive_random_int(0, N - 1 - 2 * blow, True)
        for k in range(3):
            ch.send(int(fe[idx + k * blow]).to_bytes(8, "big"))
            ch.send(ft.path(idx + k * blow))
        for lay, t in zip(layers, trees):
            n = len(lay)
            if n == 1:
                ch.send(int(lay[0]).to_bytes(8, "big"))
            for which in (idx % n, (idx % n + n // 2) % n):
                ch.send(int(lay[which]).to_bytes(8, "big"))
                ch.send(t.path(which))
    return ch.proof


def test_stark101_verifier_rejects_a_garbage_trace(sp, orc):
    """ADVICE r1 (high), the STARK side: a random 'trace LDE', the composition polynomial computed point-wise from it
    (so the trace <-> CP link holds) and an arbitrary claimed a_{T-2}; FRI folded down to one point.  Rejected: the
    statement caps the FRI at log_trace folds.  With the cap respected the non-constant last layer is caught; with a
    layer 0 that is not the composition polynomial the trace <-> CP link is."""
    log_t, log_b, q = 4, 3, 16
    log_n = log_t + log_b
    T, N, blow = 1 << log_t, 1 << log_n, 1 << log_b
    g, h, w = orc.root_of_unity(log_t), orc.root_of_unity(log_n), 5
    claimed = 123456789
    fe = orc.synthetic_column(4242, N)                               # garbage "f on the coset"
    inv = lambda v: pow(int(v), P - 2, P)
    x_last, ex = pow(g, T - 2, P), [pow(g, T - 3, P), pow(g, T - 2, P), pow(g, T - 1, P)]

    def cp_pointwise(al):
        cp = np.zeros(N, dtype=np.uint64)
        for i in range(N):
            x = w * pow(h, i, P) % P
            fx, fgx, fg2x = int(fe[i]), int(fe[(i + blow) % N]), int(fe[(i + 2 * blow) % N])
            p0 = (fx - 1) * inv(x - 1) % P
            p1 = (fx - claimed) * inv(x - x_last) % P
            e3 = (x - ex[0]) * (x - ex[1]) * (x - ex[2]) % P
            p2 = (fg2x - fgx * fgx - fx * fx) * e3 % P * inv(pow(x, T, P) - 1) % P
            cp[i] = (al[0] * p0 + al[1] * p1 + al[2] * p2) % P
        return cp

    msgs = _forged_stark_transcript(sp, orc, log_t, log_b, q, claimed, fe, cp_pointwise, folds=log_n)
    ok, why = sp.stark101_verify(_flat(msgs), claimed, log_t, log_b, q)
    assert not ok and "degree bound" in why, why
    msgs = _forged_stark_transcript(sp, orc, log_t, log_b, q, claimed, fe, cp_pointwise, folds=log_t)
    ok, why = sp.stark101_verify(_flat(msgs), claimed, log_t, log_b, q)
    assert not ok and "final constant" in why, why
    msgs = _forged_stark_transcript(sp, orc, log_t, log_b, q, claimed, fe, lambda al: orc.synthetic_column(99, N), folds=log_t)
    ok, why = sp.stark101_verify(_flat(msgs), claimed, log_t, log_b, q)
    assert not ok and "composition" in why, why


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
    ok, why = sp.verify_fri(ch.proof_flat(), log_n, 5, q, (1 << log_n) - 1, log_deg)
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
    ok, why = sp.stark101_verify(ch.proof_flat(), claimed + 1)          # a false claim about a_1022: not this transcript's statement
    assert not ok and "statement" in why
    msgs = ch.proof
    for pos in (0, 1, 2, 5, 21, 22, 23, 24, len(msgs) - 1):              # statement, trace root, alpha, CP root, final..., openings
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
    assert sp.verify_fri(flat(msgs), log_n, off, q, (1 << log_n) - 1, log_n)[0]
    lone = len(msgs) - 5                      # ... lone element, elem, path, sibling elem, sibling path
    assert len(msgs[lone]) == 8 and msgs[lone] == msgs[lone + 1]
    bad = [bytearray(m) for m in msgs]
    bad[lone][7] ^= 1
    ok, why = sp.verify_fri(flat(bad), log_n, off, q, (1 << log_n) - 1, log_n)
    assert not ok and "length-1" in why


def test_host_soak_short():
    """A few seconds of the randomised host-side differential run (channel op sequences, Merkle validation, verifiers
    with bit flips); the long form is `python tools/soak_host.py 60`."""
    import subprocess, sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "soak_host.py"), "4", "99"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
