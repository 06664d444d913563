"""CPU tests: the oracle against the reference's own known answers, the spec anchors, an independent
hashlib restatement, and itself (literal tier == fast tier)."""
import hashlib

import numpy as np
import pytest

P = 3221225473


# ---------------------------------------------------------------- reference KATs (GF(7))
def test_reference_field_kats(orc, golden):
    k = golden["ref_kat"]
    M = k["modulus"]
    for c in k["field"]:
        if c["op"] == "inverse":
            assert orc.fe_inverse(c["a"], M) == c["out"], c["ref"]
        elif c["op"] == "pow":
            assert orc.fe_pow(c["a"], c["e"], M) == c["out"], c["ref"]
        elif c["op"] == "div":
            assert orc.fe_div(c["a"], c["b"], M) == c["out"], c["ref"]


def test_reference_poly_kats(orc, golden):
    k = golden["ref_kat"]
    M = k["modulus"]
    for c in k["poly_mul"]:
        assert orc.poly_mul(c["a"], c["b"], M).tolist() == c["out"], c["ref"]
    for c in k["poly_div_rem"]:
        q, r = orc.poly_div_rem(c["a"], c["b"], M)
        assert q.tolist() == c["q"] and r.tolist() == c["r"], c["ref"]
    for c in k["from_roots"]:
        assert orc.poly_from_roots(c["roots"], M).tolist() == c["out"], c["ref"]
    for c in k["interpolate"]:
        assert orc.poly_interpolate(c["xs"], c["ys"], M).tolist() == c["out"], c["ref"]


def test_lagrange_basis_delta(orc):
    # src/polynomial/interpolation.rs:187-221: L_i(x_j) == delta_ij
    xs, M = [1, 2, 3, 4, 5], 7
    L = orc.lagrange_basis(xs, M)
    for i in range(5):
        for j in range(5):
            assert orc.poly_evaluate(L[i], xs[j], M) == (1 if i == j else 0)


def test_field_semantics(orc):
    M = 7
    assert orc.fe_inverse(0, M) == 0                      # element.rs:54-57: 0^(M-2) == 0, no panic
    assert orc.fe_neg(0, M) == 0                          # :130-136
    assert orc.fe_sub(2, 5, M) == 4                       # :86-92
    assert orc.fe_from_int(-3, M) == 4                    # :138-147
    assert orc.fe_new(10, M) == 3
    assert orc.poly_trim([1, 2, 0, 0]).tolist() == [1, 2]  # ops.rs:19-37
    with pytest.raises(ZeroDivisionError):
        orc.poly_div_rem([1, 2], [0, 0], M)               # ops.rs:142-144


def test_div_rem_invariant(orc):
    # ops.rs:1044-1067: a == q*b + r, deg r < deg b
    rng = np.random.default_rng(1)
    for _ in range(20):
        a = rng.integers(0, P, rng.integers(1, 40)).astype(np.uint64)
        b = rng.integers(0, P, rng.integers(1, 12)).astype(np.uint64)
        if len(orc.poly_trim(b)) == 0:
            continue
        q, r = orc.poly_div_rem(a, b, P)
        back = orc.poly_add(orc.poly_mul(q, b, P), r, P)
        assert back.tolist() == orc.poly_trim(a % np.uint64(P)).tolist()
        assert len(r) < len(orc.poly_trim(b))


# ---------------------------------------------------------------- SHA-256 / merkle / channel
def test_sha256_vectors(orc):
    # FIPS 180-4 examples + lengths around the padding boundaries, both code paths
    kat = {b"abc": "ba7816bf8f01cfea414140de5dae2223b00361a396177a9cb410ff61f20015ad",
           b"": "e3b0c44298fc1c149afbf4c8996fb92427ae41e4649b934ca495991b7852b855",
           b"abcdbcdecdefdefgefghfghighijhijkijkljklmklmnlmnomnopnopq":
               "248d6a61d20638b8e5c026930c3e6039a33ce45964ff2167f6ecedd419db06c1"}
    for accel in (0, 1):
        orc.lib().or_sha256_set_accel(accel)
        for m, h in kat.items():
            assert orc.sha256(m).hex() == h
        for n in (1, 8, 55, 56, 57, 63, 64, 65, 119, 120, 128, 1000):
            m = bytes((i * 7 + 3) & 255 for i in range(n))
            assert orc.sha256(m) == hashlib.sha256(m).digest()
    orc.lib().or_sha256_set_accel(1)


def test_spec_anchors(orc, golden):
    a = golden["spec_anchors"]
    for v, h in a["leaf"].items():
        assert orc.Tree([int(v)]).root_hex() == h
    assert orc.Tree(range(8)).root_hex() == a["root_0_to_7"]
    assert orc.Tree([0, 1, 2]).root_hex() == a["root_0_1_2"]          # promotion rule
    c = a["channel"]
    ch = orc.Channel(a["modulus"])
    ch.send(a["root_0_to_7"].encode())
    assert ch.state == c["after_send_root"]
    assert ch.receive_random_field_element() == c["field_element"] and ch.state == c["after_field_element"]
    assert ch.receive_random_int(0, 8191, True) == c["random_int_0_8191"] and ch.state == c["after_random_int"]
    assert ch.proof_size() == c["proof_size"] and ch.compressed_proof_size() == c["compressed_proof_size"]
    s = a["stark101"]
    assert orc.root_of_unity(10) == s["g_1024"] and orc.root_of_unity(13) == s["h_8192"]
    assert orc.root_of_unity(20) == s["w_2e20"] and orc.root_of_unity(24) == s["w_2e24"]
    assert orc.root_of_unity(25) == s["w_2e25"] and orc.root_of_unity(26) == s["w_2e26"]
    assert int(orc.fibsq_trace(3141592, 1023)[1022]) == s["a_1022"]


def test_rs_merkle_published_vector(orc, golden):
    """The tree rule against the root rs_merkle publishes for the leaves "a".."f" (6 leaves: exercises promotion)."""
    v = golden["spec_anchors"]["rs_merkle_published"]
    digests = [hashlib.sha256(x.encode()).digest() for x in v["leaves_utf8"]]
    assert orc.merkle_root_from_digests(digests).hex() == v["root_hex"]
    lv = list(digests)                                   # and the hashlib twin's rule
    while len(lv) > 1:
        nxt = [hashlib.sha256(lv[i] + lv[i + 1]).digest() for i in range(0, len(lv) - 1, 2)]
        if len(lv) & 1:
            nxt.append(lv[-1])
        lv = nxt
    assert lv[0].hex() == v["root_hex"]


@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 13, 64, 100, 257])
def test_merkle_vs_hashlib(orc, n):
    vals = orc.synthetic_column(n, n).tolist()
    t = orc.Tree(vals)
    lv = orc.py_merkle_levels(vals)
    assert t.root_hex() == lv[-1][0].hex()
    assert t.depth == len(lv) - 1
    for l, level in enumerate(lv):
        for j, d in enumerate(level):
            assert t.node(l, j) == d
    for idx in {0, n // 2, n - 1}:
        path = t.path(idx)
        assert path == orc.py_merkle_path(vals, idx)
        assert orc.merkle_verify(t.root(), n, idx, vals[idx], path)
        assert not orc.merkle_verify(t.root(), n, idx, (vals[idx] + 1) % P, path)
    assert orc.merkle_root_only(vals) == t.root()


def test_merkle_empty_panics(orc):
    with pytest.raises(ValueError):
        orc.Tree([])                                       # merkle/mod.rs:25: unwrap on None


def test_channel_vs_hashlib(orc):
    a, b = orc.Channel(P), orc.PyChannel(P)
    for step in range(40):
        if step % 3 == 0:
            m = bytes((step * 11 + i) & 255 for i in range(step % 9 + 1))
            a.send(m); b.send(m)
        elif step % 3 == 1:
            assert a.receive_random_field_element() == b.receive_random_field_element()
        else:
            lo, hi = step, step * 1000 + 7
            assert a.receive_random_int(lo, hi, step % 2 == 0) == b.receive_random_int(lo, hi, step % 2 == 0)
        assert a.state == b.state
    assert a.proof == b.proof and a.compressed_proof == b.compressed_proof
    assert a.proof_size() == b.proof_size()


# ---------------------------------------------------------------- literal tier == fast tier
@pytest.mark.parametrize("log_n,log_deg,offset", [(4, 2, 5), (8, 5, 5), (10, 7, 3), (9, 9, 7)])
def test_ntt_matches_horner(orc, log_n, log_deg, offset):
    w = orc.root_of_unity(log_n)
    c = orc.synthetic_column(log_n * 31 + log_deg, 1 << log_deg)
    d = orc.coset_domain(offset, w, 1 << log_n, P)
    lit = np.array([orc.poly_evaluate(c, int(x), P) for x in d], dtype=np.uint64)
    assert np.array_equal(orc.coset_evaluate(c, log_n, offset, w, P), lit)
    if log_deg == log_n:
        assert np.array_equal(orc.coset_interpolate(lit, log_n, offset, w, P), c)
    if offset == 5:
        # offset-1 transforms
        d1 = orc.coset_domain(1, w, 1 << log_n, P)
        cc = np.zeros(1 << log_n, dtype=np.uint64); cc[: len(c)] = c
        e1 = np.array([orc.poly_evaluate(c, int(x), P) for x in d1], dtype=np.uint64)
        assert np.array_equal(orc.ntt(cc, log_n, w, P), e1)
        assert np.array_equal(orc.intt(e1, log_n, w, P), cc)


def test_interpolate_matches_lagrange(orc):
    log_n, offset = 5, 5
    w = orc.root_of_unity(log_n)
    xs = orc.coset_domain(offset, w, 1 << log_n, P)
    ys = orc.synthetic_column(9, 1 << log_n)
    lit = orc.poly_interpolate(xs, ys, P)
    fast = orc.poly_trim(orc.coset_interpolate(ys, log_n, offset, w, P))
    assert np.array_equal(lit, fast)


def test_batch_inverse(orc):
    a = orc.synthetic_column(3, 1000)
    a[[0, 17, 999]] = 0
    inv = orc.batch_inverse(a, P)
    for i in (0, 1, 17, 500, 999):
        assert int(inv[i]) == orc.fe_inverse(int(a[i]), P)
    assert inv[0] == 0 and inv[17] == 0


def test_fold_matches_coefficient_fold(orc):
    log_n, offset, beta = 8, 5, 123456789
    w = orc.root_of_unity(log_n)
    c = orc.synthetic_column(5, 1 << 5)
    e = orc.coset_evaluate(c, log_n, offset, w, P)
    c2 = orc.next_fri_polynomial(c, beta, P)
    d2 = orc.next_fri_domain(orc.coset_domain(offset, w, 1 << log_n, P), P)
    lit = np.array([orc.poly_evaluate(c2, int(x), P) for x in d2], dtype=np.uint64)
    assert np.array_equal(orc.fri_fold_evals(e, beta, offset, w, P), lit)


@pytest.mark.parametrize("log_n,log_deg,offset,q", [(10, 7, 5, 3), (8, 8, 3, 2), (6, 0, 5, 1), (7, 3, 9, 2)])
def test_fri_literal_equals_fast(orc, log_n, log_deg, offset, q):
    w = orc.root_of_unity(log_n)
    c = orc.synthetic_poly_exact_degree(100 + log_n, 1 << log_deg)
    d = orc.coset_domain(offset, w, 1 << log_n, P)
    c1, c2 = orc.Channel(P), orc.Channel(P)
    p1 = orc.fri_commit_literal(c, d, c1, P)
    p2 = orc.fri_commit_fast(c, log_n, offset, w, c2, P)
    orc.decommit_fri(q, (1 << log_n) - 1, p1, c1)
    orc.decommit_fri(q, (1 << log_n) - 1, p2, c2)
    assert p1.num_layers == p2.num_layers == log_deg + 1
    for k in range(p1.num_layers):
        assert np.array_equal(p1.layer(k), p2.layer(k))
    assert c1.state == c2.state and c1.proof == c2.proof
    assert np.array_equal(p1.final_poly(), p2.final_poly())


def test_fri_zero_and_constant_poly(orc):
    w = orc.root_of_unity(4)
    for coeffs, fin in (([0, 0, 0], []), ([7], [7])):
        ch = orc.Channel(P)
        pr = orc.fri_commit_fast(coeffs, 4, 5, w, ch, P)
        assert pr.num_layers == 1 and pr.final_poly().tolist() == fin
        assert len(ch.proof) == 2 and ch.proof[1] == (fin[0] if fin else 0).to_bytes(8, "big")


def test_golden_transcripts(orc, golden):
    """The committed fixtures still reproduce (fast tier against literal-tier goldens)."""
    g = golden["transcripts"]
    for case in g["fri"]:
        log_n = case["log_n"]
        w = orc.root_of_unity(log_n)
        c = orc.synthetic_poly_exact_degree(case["seed"], 1 << case["log_deg"])
        ch = orc.Channel(P)
        pr = orc.fri_commit_fast(c, log_n, case["offset"], w, ch, P)
        assert ch.state == case["state_after_commit"]
        orc.decommit_fri(case["queries"], (1 << log_n) - 1, pr, ch)
        assert pr.num_layers == case["num_layers"]
        assert [pr.tree(k).root_hex() for k in range(pr.num_layers)] == case["roots"]
        assert [hashlib.sha256(pr.layer(k).astype("<u8").tobytes()).hexdigest() for k in range(pr.num_layers)] == case["layer_sha256"]
        assert ch.state == case["final_state"] and ch.proof_size() == case["proof_size"]
        assert hashlib.sha256(ch.proof_flat()).hexdigest() == case["proof_sha256"]
    s = g["stark101"]
    for literal in (False, True):
        ch = orc.Channel(P)
        orc.stark101_prove(ch, literal=literal)
        assert ch.state == s["final_state"] and ch.proof_size() == s["proof_size"]
        assert hashlib.sha256(ch.proof_flat()).hexdigest() == s["proof_sha256"]
        assert ch.proof[0].hex() == s["statement"] and ch.proof[1].decode() == s["first_root"]
        assert ch.proof[0] == b"".join(int(v).to_bytes(8, "big") for v in (P, 5, 10, 3, 3, int(orc.fibsq_trace(3141592, 1023)[1022])))
