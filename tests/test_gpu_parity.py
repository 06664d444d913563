"""GPU parity tests: every result of the CUDA path, called through the C ABI, must be bit-identical to
the CPU oracle on the same seeded inputs (integer work: exact equality, no tolerance)."""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
P = 3221225473


# ---------------------------------------------------------------- merkle
@pytest.mark.parametrize("n", [1, 2, 3, 5, 7, 8, 9, 13, 64, 100, 257, 511, 512, 513, 1000, 4096, 4097, 12345, 1 << 16])
def test_merkle_root_and_levels(sp, orc, ctx, n):
    vals = orc.synthetic_column(n, n)
    t = sp.MerkleTree.new(ctx, vals)
    ot = orc.Tree(vals)
    assert t.root() == ot.root_hex()
    assert t.depth == ot.depth and t.num_leaves == n
    rng = np.random.default_rng(n)
    for l in range(1, t.depth + 1):
        m = -(-n // (1 << l))
        for j in {0, m - 1, int(rng.integers(0, m))}:
            assert t.node(l, j) == ot.node(l, j), (l, j)
    for idx in {0, n - 1, n // 2, int(rng.integers(0, n))}:
        assert t.get_authentication_path(idx) == ot.path(idx), idx


@pytest.mark.parametrize("n", [255, 256, 258, 769, 32767, 32768, 32769, 65537, 100003])
def test_merkle_every_node(sp, orc, ctx, n):
    """Sizes around the chunk (256 items per CTA) and launch (2^15 items) boundaries of the tail kernel; ragged at many
    levels; every stored node compared, not a sample."""
    vals = orc.synthetic_column(1000 + n, n)
    t = sp.MerkleTree.new(ctx, vals)
    ot = orc.Tree(vals)
    assert t.root() == ot.root_hex()
    for l in range(1, t.depth + 1):
        m = -(-n // (1 << l))
        step = 1 if m <= 4096 else 7                      # all nodes of the upper levels, every 7th (+ the last) below
        for j in list(range(0, m, step)) + [m - 1]:
            assert t.node(l, j) == ot.node(l, j), (l, j)


def test_merkle_anchors(sp, ctx, golden):
    a = golden["spec_anchors"]
    for v, h in a["leaf"].items():
        assert sp.MerkleTree.new(ctx, [int(v)]).root() == h
    assert sp.MerkleTree.new(ctx, list(range(8))).root() == a["root_0_to_7"]
    assert sp.MerkleTree.new(ctx, [0, 1, 2]).root() == a["root_0_1_2"]


def test_merkle_empty_is_error(sp, ctx):
    with pytest.raises(sp.StarkError) as e:
        sp.MerkleTree.new(ctx, [])
    assert e.value.code == 1


def test_merkle_reduces_like_field_new(sp, orc, ctx):
    # FieldElement::new reduces mod p before hashing (element.rs:13-17)
    vals = np.array([P, P + 1, 2 * P + 5, 7], dtype=np.uint64)
    assert sp.MerkleTree.new(ctx, vals).root() == orc.Tree(vals % np.uint64(P)).root_hex()


def test_merkle_large_2e20(sp, orc, ctx):
    n = 1 << 20
    vals = orc.synthetic_column(42, n)
    t = sp.MerkleTree.new(ctx, vals)
    assert t.root_bytes() == orc.merkle_root_only(vals)
    ot = orc.Tree(vals)
    for idx in (0, 1, n - 1, 777777):
        p = t.get_authentication_path(idx)
        assert p == ot.path(idx)
        assert orc.merkle_verify(t.root_bytes(), n, idx, int(vals[idx]), p)


# ---------------------------------------------------------------- NTT / LDE
@pytest.mark.parametrize("log_n", list(range(0, 21)))
def test_ntt_roundtrip_and_oracle(sp, orc, ctx, log_n):
    n = 1 << log_n
    a = orc.synthetic_column(1000 + log_n, n)
    w = orc.root_of_unity(log_n)
    fwd = ctx.ntt(a, log_n)
    assert np.array_equal(fwd, orc.ntt(a, log_n, w, P))
    assert np.array_equal(ctx.intt(fwd, log_n), a)
    assert np.array_equal(ctx.intt(a, log_n), orc.intt(a, log_n, w, P))


@pytest.mark.parametrize("log_n,log_deg,offset", [(3, 0, 5), (4, 2, 5), (10, 7, 5), (13, 10, 5), (13, 13, 3), (17, 14, 5),
                                                   (20, 17, 5), (19, 10, 7), (12, 0, 5), (22, 19, 5)])
def test_coset_evaluate(sp, orc, ctx, log_n, log_deg, offset):
    c = orc.synthetic_column(log_n * 100 + log_deg, 1 << log_deg)
    w = orc.root_of_unity(log_n)
    got = ctx.coset_evaluate(c, log_n, offset)
    assert np.array_equal(got, orc.coset_evaluate(c, log_n, offset, w, P))
    if log_n <= 10:   # against the reference's literal Horner
        d = orc.coset_domain(offset, w, 1 << log_n, P)
        assert got.tolist() == [orc.poly_evaluate(c, int(x), P) for x in d]


def test_coset_evaluate_ragged_lengths(sp, orc, ctx):
    w = orc.root_of_unity(10)
    for ln in (0, 1, 3, 5, 100, 129, 1023):
        c = orc.synthetic_column(ln + 1, ln)
        assert np.array_equal(ctx.coset_evaluate(c, 10, 5), orc.coset_evaluate(c, 10, 5, w, P)), ln


@pytest.mark.parametrize("log_n,offset", [(0, 5), (1, 5), (5, 5), (10, 3), (14, 5), (18, 5)])
def test_coset_interpolate(sp, orc, ctx, log_n, offset):
    e = orc.synthetic_column(77 + log_n, 1 << log_n)
    w = orc.root_of_unity(log_n)
    got = ctx.coset_interpolate(e, log_n, offset)
    assert np.array_equal(got, orc.coset_interpolate(e, log_n, offset, w, P))
    if log_n == 5:   # the reference's Lagrange interpolation
        xs = orc.coset_domain(offset, w, 1 << log_n, P)
        assert np.array_equal(orc.poly_trim(got), orc.poly_interpolate(xs, e, P))


@pytest.mark.parametrize("log_n,log_blowup,off_in,off_out", [(0, 3, 1, 5), (3, 1, 1, 5), (10, 3, 1, 5), (12, 3, 5, 5),
                                                             (17, 3, 1, 5), (16, 4, 7, 3), (20, 2, 1, 5), (9, 0, 1, 5)])
def test_coset_lde(sp, orc, ctx, log_n, log_blowup, off_in, off_out):
    e = orc.synthetic_column(5 + log_n, 1 << log_n)
    got = ctx.coset_lde(e, log_n, off_in, log_blowup, off_out)
    coeffs = orc.coset_interpolate(e, log_n, off_in, orc.root_of_unity(log_n), P)
    want = orc.coset_evaluate(coeffs, log_n + log_blowup, off_out, orc.root_of_unity(log_n + log_blowup), P)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("log_n", list(range(10, 21)))
def test_natural_transform_every_size(sp, orc, ctx, log_n):
    """The permutation-free natural-order transform (ntt.cu, ntt_natural): 2, 3 digits, coset scaling on the way in
    (evaluate) and on the way out (interpolate), blow-ups 1, 2 and 4, coefficient counts that are not a multiple of 4."""
    w = orc.root_of_unity(log_n)
    n = 1 << log_n
    for nco in (n, n - 1, n // 2, n // 2 + 3, n // 4, n // 4 - 5, 1):
        c = orc.synthetic_column(31 * log_n + nco % 97, nco)
        assert np.array_equal(ctx.coset_evaluate(c, log_n, 7), orc.coset_evaluate(c, log_n, 7, w, P)), nco
    e = orc.synthetic_column(log_n, n)
    e[:4] = [P - 1, 0, P - 1, 1]
    got = ctx.coset_interpolate(e, log_n, 11)
    assert np.array_equal(got, orc.coset_interpolate(e, log_n, 11, w, P))
    assert np.array_equal(ctx.coset_evaluate(got, log_n, 11), e)


@pytest.mark.parametrize("modulus,gen", [(2013265921, 31), (998244353, 3), (4293918721, 19), (3489660929, 3), (2281701377, 3)])
def test_natural_transform_other_moduli(sp, orc, modulus, gen):
    """strict butterflies for p < 2^31, weak ones above; values next to p."""
    c = sp.Context(modulus, gen, 0)
    try:
        for log_n in (10, 13, 17, 19):
            if log_n > c.two_adicity:
                continue
            w = orc.root_of_unity(log_n, modulus, gen)
            a = orc.synthetic_column(log_n, 1 << log_n, modulus)
            a[:6] = [modulus - 1, modulus - 1, 0, modulus - 2, 1, modulus - 1]
            assert np.array_equal(c.coset_evaluate(a, log_n, gen), orc.coset_evaluate(a, log_n, gen, w, modulus)), log_n
            assert np.array_equal(c.coset_interpolate(a, log_n, gen), orc.coset_interpolate(a, log_n, gen, w, modulus)), log_n
            assert np.array_equal(c.ntt(a, log_n), orc.ntt(a, log_n, w, modulus)), log_n
    finally:
        c.close()


def test_natural_transform_four_digits_2e28(sp, orc, ctx):
    """2^28 points = four digits of 7 bits: a sparse polynomial evaluated on the coset, spot-checked with pow()."""
    log_n, off = 28, 5
    n = 1 << log_n
    terms = {0: 17, 1: P - 1, 130: 99, (1 << 14) + 5: 123456789, (1 << 21) + 77: 3, n - 1: P - 2, n // 2: 42}
    c = np.zeros(n, dtype=np.uint64)
    for j, v in terms.items():
        c[j] = v
    cv = ctx.upload(c)
    del c
    ev = ctx.coset_evaluate_dev(cv, log_n, off)
    w = orc.root_of_unity(log_n)
    rng = np.random.default_rng(5)
    idx = [0, 1, 127, 128, n - 1, n // 2, (1 << 21) + 1] + [int(i) for i in rng.integers(0, n, 40)]
    for i in idx:
        x = off * pow(w, i, P) % P
        want = sum(v * pow(x, j, P) for j, v in terms.items()) % P
        assert int(ev.download(i, 1)[0]) == want, i
    back = ctx.coset_interpolate_dev(ev, off)
    ev.free()
    for j, v in terms.items():
        assert int(back.download(j, 1)[0]) == v
    blk = back.download(1 << 14, 1 << 12)
    assert int(blk[5]) == 123456789 and np.count_nonzero(blk) == 1
    back.free(); cv.free()


def test_natural_transform_maximum_size_2e30(sp, orc, ctx):
    """The largest domain this field has (2-adicity 30; digits 8+8+7+7): a sparse polynomial built on the device,
    evaluated on the coset, spot-checked with pow(), and interpolated back."""
    import torch
    log_n, off = 30, 5
    n = 1 << log_n
    terms = {0: 5, 3: P - 1, (1 << 15) + 9: 77, (1 << 22) + 1: 1234567, (1 << 29) + 12345: 42, n - 1: P - 3}
    t = torch.zeros(n, dtype=torch.int32, device="cuda")
    for j, v in terms.items():
        t[j] = v - (1 << 32) if v >= (1 << 31) else v          # the u32 bit pattern
    torch.cuda.synchronize()
    cv = ctx.from_device(t.data_ptr(), n)
    del t
    torch.cuda.empty_cache()
    ev = ctx.coset_evaluate_dev(cv, log_n, off)
    cv.free()
    w = orc.root_of_unity(log_n)
    rng = np.random.default_rng(30)
    for i in [0, 1, 255, 256, n - 1, n // 2, (1 << 23) + 5] + [int(x) for x in rng.integers(0, n, 24)]:
        x = off * pow(w, i, P) % P
        assert int(ev.download(i, 1)[0]) == sum(v * pow(x, j, P) for j, v in terms.items()) % P, i
    back = ctx.coset_interpolate_dev(ev, off)
    ev.free()
    for j, v in terms.items():
        assert int(back.download(j, 1)[0]) == v, j
    blk = back.download((1 << 29) + 12288, 1 << 12)
    assert int(blk[12345 - 12288]) == 42 and np.count_nonzero(blk) == 1
    back.free()


def test_coset_domain(sp, orc, ctx):
    for log_n in (0, 1, 7, 13):
        assert np.array_equal(ctx.coset_domain(log_n, 5), orc.coset_domain(5, orc.root_of_unity(log_n), 1 << log_n, P))


def test_ntt_linearity_large(sp, orc, ctx):
    # size-independent property at 2^22: NTT(a + c*b) == NTT(a) + c*NTT(b)
    log_n, c = 22, 987654321
    a, b = orc.synthetic_column(1, 1 << log_n), orc.synthetic_column(2, 1 << log_n)
    comb = ((a.astype(object) + c * b.astype(object)) % P).astype(np.uint64)
    fa, fb, fc = ctx.ntt(a, log_n), ctx.ntt(b, log_n), ctx.ntt(comb, log_n)
    want = ((fa.astype(object) + c * fb.astype(object)) % P).astype(np.uint64)
    assert np.array_equal(fc, want)


def test_unsupported_sizes(sp, ctx):
    with pytest.raises(sp.StarkError):
        ctx.coset_evaluate(np.zeros(4, dtype=np.uint64), 31, 5)
    with pytest.raises(sp.StarkError):
        ctx.coset_evaluate(np.zeros(32, dtype=np.uint64), 4, 5)      # more coefficients than points
    with pytest.raises(sp.StarkError):
        ctx.coset_evaluate(np.zeros(4, dtype=np.uint64), 4, 0)       # zero offset


# ---------------------------------------------------------------- inverse / quotient
@pytest.mark.parametrize("n", [0, 1, 7, 2048, 2049, 100003])
def test_batch_inverse(sp, orc, ctx, n):
    a = orc.synthetic_column(n + 3, n)
    if n > 5:
        a[[0, 3, n - 1]] = 0                                          # inverse(0) == 0 (element.rs:54-57)
    got = ctx.batch_inverse(a)
    assert np.array_equal(got, orc.batch_inverse(a, P))
    if n:
        i = n // 2
        assert int(got[i]) == orc.fe_inverse(int(a[i]), P)


@pytest.mark.parametrize("n", [4095, 4096, 4097, 8191, 12289])
def test_batch_inverse_cta_boundaries_and_zero_patterns(sp, orc, ctx, n):
    """The kernel inverts one product per CTA of 4096 elements (16 strided elements per thread, 256 threads): sizes
    around that boundary, threads / warps / whole CTAs whose elements are all zero, and the values 1 and p-1."""
    a = orc.synthetic_column(n, n)
    a[5::256] = 0                      # every element of thread 5 of every CTA (element i belongs to thread i mod 256)
    a[32:64] = 0                       # a whole warp's first element
    a[1024:1024 + 7] = [1, P - 1, 2, P - 2, 0, 1, P - 1]
    got = ctx.batch_inverse(a)
    assert np.array_equal(got, orc.batch_inverse(a, P))
    z = np.zeros(n, dtype=np.uint64)
    assert not ctx.batch_inverse(z).any()                            # all zeros: every product is the empty product
    if n > 4096:
        a[:4096] = 0                                                 # first CTA entirely zero, the rest not
        assert np.array_equal(ctx.batch_inverse(a), orc.batch_inverse(a, P))


def test_quotient_pointwise(sp, orc, ctx):
    n = 5000
    num, den = orc.synthetic_column(1, n), orc.synthetic_column(2, n)
    den[[5, 77]] = 0
    got = ctx.quotient_pointwise(num, den)
    inv = orc.batch_inverse(den, P)
    want = np.array([orc.fe_mul(int(x), int(y), P) for x, y in zip(num, inv)], dtype=np.uint64)
    assert np.array_equal(got, want)
    assert got[5] == 0                                                # a / 0 == a * 0 (element.rs:116-122)


def test_vanishing_quotient_is_polynomial(sp, orc, ctx):
    # (f(x) - f(1)) / (x - 1) evaluated point-wise on the coset interpolates to a polynomial of degree deg f - 1
    log_t, log_n = 6, 9
    f = orc.synthetic_column(11, 1 << log_t)
    ev = ctx.coset_evaluate(f, log_n, 5)
    f1 = orc.poly_evaluate(f, 1, P)
    dom = ctx.coset_domain(log_n, 5)
    num = (ev + np.uint64(P) - np.uint64(f1)) % np.uint64(P)
    den = (dom + np.uint64(P) - np.uint64(1)) % np.uint64(P)
    q = orc.poly_trim(ctx.coset_interpolate(ctx.quotient_pointwise(num, den), log_n, 5))
    ql, rl = orc.poly_div_rem(orc.poly_sub(f, [f1], P), [P - 1, 1], P)            # ops.rs:141-191
    assert len(rl) == 0 and np.array_equal(q, ql)


# ---------------------------------------------------------------- FRI
def _check_fri(sp, orc, ctx, coeffs, log_n, offset, queries):
    w = orc.root_of_unity(log_n)
    ch, och = sp.Channel(P), orc.Channel(P)
    pr = sp.fri_commit(ctx, coeffs, sp.CosetFri(ctx, offset, log_n), ch)
    opr = orc.fri_commit_fast(coeffs, log_n, offset, w, och, P)
    assert pr.num_layers == opr.num_layers
    for k in range(pr.num_layers):
        assert pr.layer_len(k) == len(opr.layer(k))
        assert np.array_equal(pr.layer(k), opr.layer(k)), f"layer {k}"
        assert pr.tree(k).root() == opr.tree(k).root_hex(), f"root {k}"
    assert np.array_equal(pr.final_poly(), opr.final_poly())
    assert ch.state == och.state
    sp.decommit_fri(queries, (1 << log_n) - 1, pr, ch)
    orc.decommit_fri(queries, (1 << log_n) - 1, opr, och)
    assert ch.state == och.state and ch.proof == och.proof
    assert ch.proof_size() == och.proof_size() and ch.compressed_proof_size() == och.compressed_proof_size()
    return pr, opr, ch


@pytest.mark.parametrize("log_n,log_deg,offset,q", [(10, 7, 5, 3), (13, 10, 5, 3), (8, 8, 3, 2), (6, 6, 3, 2), (4, 0, 5, 1),
                                                    (16, 13, 5, 4), (12, 3, 9, 2), (3, 1, 5, 1), (20, 17, 5, 8)])
def test_fri_commit_and_decommit(sp, orc, ctx, log_n, log_deg, offset, q):
    c = orc.synthetic_poly_exact_degree(43 + log_n, 1 << log_deg)
    _check_fri(sp, orc, ctx, c, log_n, offset, q)


def test_fri_non_power_of_two_degree_and_trailing_zeros(sp, orc, ctx):
    c = orc.synthetic_column(9, 300)
    c[-40:] = 0                                                       # Polynomial::new trims (ops.rs:19-37)
    _check_fri(sp, orc, ctx, c, 12, 5, 2)


def test_fri_zero_and_constant(sp, orc, ctx):
    for coeffs in ([0, 0, 0], [7], [0]):
        pr, _, ch = _check_fri(sp, orc, ctx, np.array(coeffs, dtype=np.uint64), 4, 5, 1)
        assert pr.num_layers == 1


def test_fri_leading_coefficient_cancels(sp, orc, ctx):
    """A fold whose beta kills the leading coefficient drops the degree by more than half; the layer
    count must follow the exact degree (fri_commit.rs:89), so drive the step API with a chosen beta."""
    log_n = 8
    c = orc.synthetic_poly_exact_degree(5, 32)
    beta = orc.fe_mul(orc.fe_neg(int(c[30]), P), orc.fe_inverse(int(c[31]), P), P)   # c30 + beta*c31 == 0
    pr, _ = sp.fri_begin(ctx, c, log_n, 5)
    assert pr.degree == 31
    pr.fold(beta)
    want = orc.next_fri_polynomial(c, beta, P)
    assert pr.degree == len(want) - 1 < 15
    e1 = orc.fri_fold_evals(orc.coset_evaluate(c, log_n, 5, orc.root_of_unity(log_n), P), beta, 5, orc.root_of_unity(log_n), P)
    assert np.array_equal(pr.layer(1), e1)


def test_fri_golden_transcripts(sp, orc, ctx, golden):
    for case in golden["transcripts"]["fri"]:
        log_n = case["log_n"]
        c = orc.synthetic_poly_exact_degree(case["seed"], 1 << case["log_deg"])
        ch = sp.Channel(P)
        pr = sp.fri_commit(ctx, c, sp.CosetFri(ctx, case["offset"], log_n), ch)
        assert ch.state == case["state_after_commit"]
        sp.decommit_fri(case["queries"], (1 << log_n) - 1, pr, ch)
        assert [pr.tree(k).root() for k in range(pr.num_layers)] == case["roots"]
        assert [hashlib.sha256(pr.layer(k).astype("<u8").tobytes()).hexdigest() for k in range(pr.num_layers)] == case["layer_sha256"]
        assert pr.final_poly().tolist() == case["final_poly"]
        assert ch.state == case["final_state"] and ch.proof_size() == case["proof_size"]
        assert hashlib.sha256(ch.proof_flat()).hexdigest() == case["proof_sha256"]


def test_fri_batched_open_matches_sequential(sp, orc, ctx):
    log_n = 12
    c = orc.synthetic_poly_exact_degree(3, 1 << 9)
    ch = sp.Channel(P)
    pr = sp.fri_commit(ctx, c, sp.CosetFri(ctx, 5, log_n), ch)
    idxs = [0, 1, 4095, 2048, 1234]
    blob = pr.open(idxs)
    ch2 = sp.Channel(P); ch2.send(b"x")
    before = len(ch2.proof)
    for i in idxs:
        sp.decommit_fri_layers(i, pr, ch2)
    assert b"".join(ch2.proof[before:]) == blob
    # every opened path verifies against its layer root
    off = 0
    for i in idxs:
        for k in range(pr.num_layers):
            n = pr.layer_len(k)
            for which in (i % n, (i % n + n // 2) % n):
                val = int.from_bytes(blob[off:off + 8], "big"); off += 8
                plen = 32 * (n.bit_length() - 1)
                assert orc.merkle_verify(pr.tree(k).root_bytes(), n, which, val, blob[off:off + plen]); off += plen
    assert off == len(blob)


def test_opening_server_falls_back_when_it_times_out(sp, orc, ctx, monkeypatch):
    """decommit_fri runs its queries through a resident kernel that gives up when no request arrives in time; the host
    must then finish with one launch per query and produce the same transcript.  A 1 us idle limit forces the hand-over."""
    log_n, q = 14, 6
    coeffs = orc.synthetic_poly_exact_degree(77, 1 << (log_n - 3))
    och = orc.Channel(P)
    opr = orc.fri_commit_fast(coeffs, log_n, 5, orc.root_of_unity(log_n), och, P)
    orc.decommit_fri(q, (1 << log_n) - 1, opr, och)
    for idle in ("1000", None):
        if idle is None:
            monkeypatch.delenv("STARK_OPEN_SERVER_IDLE_NS", raising=False)
        else:
            monkeypatch.setenv("STARK_OPEN_SERVER_IDLE_NS", idle)
        ch = sp.Channel(P)
        pr = sp.fri_commit(ctx, coeffs, sp.CosetFri(ctx, 5, log_n), ch)
        sp.decommit_fri(q, (1 << log_n) - 1, pr, ch)
        assert ch.state == och.state and ch.proof == och.proof, idle
        pr.free()


def test_opening_server_when_launches_are_serialised(orc):
    """CUDA_LAUNCH_BLOCKING=1 (like a profiler or a debugger) keeps the host inside the launch call until the resident
    opening kernel has left: the first request is already posted, so it is answered; the kernel then idles for 2 ms at most
    (250 ms in round 1) and the remaining queries are one launch each.  Same transcript, and no quarter second lost."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r"""
import importlib, sys, time
sys.path.insert(0, %r)
sp = importlib.import_module("stark-prover_b200")
from oracle import pyoracle as orc
P = sp.P_DEFAULT
log_n, q = 14, 16
coeffs = orc.synthetic_poly_exact_degree(78, 1 << (log_n - 3))
och = orc.Channel(P)
opr = orc.fri_commit_fast(coeffs, log_n, 5, orc.root_of_unity(log_n), och, P)
orc.decommit_fri(q, (1 << log_n) - 1, opr, och)
ctx = sp.Context()
best = 1e9
for rep in range(3):
    ch = sp.Channel(P)
    pr = sp.fri_commit(ctx, coeffs, sp.CosetFri(ctx, 5, log_n), ch)
    t0 = time.perf_counter()
    sp.decommit_fri(q, (1 << log_n) - 1, pr, ch)
    best = min(best, time.perf_counter() - t0)
    assert ch.state == och.state and ch.proof == och.proof
    pr.free()
print("decommit_seconds", best)
""" % root
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, CUDA_LAUNCH_BLOCKING="1"), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    secs = float(r.stdout.split("decommit_seconds")[1].split()[0])
    assert secs < 0.1, f"decommit_fri of 16 queries took {secs:.3f} s with serialised launches"


def test_fri_device_resident_input(sp, orc, ctx):
    log_n = 14
    c = orc.synthetic_poly_exact_degree(8, 1 << 11)
    v = ctx.upload(c)
    ch1, ch2 = sp.Channel(P), sp.Channel(P)
    p1 = sp.fri_commit(ctx, v, sp.CosetFri(ctx, 5, log_n), ch1)
    p2 = sp.fri_commit(ctx, c, sp.CosetFri(ctx, 5, log_n), ch2)
    assert ch1.state == ch2.state and p1.num_layers == p2.num_layers
    assert np.array_equal(v.download(), c)


def test_fri_layers_by_value_streamed_to_host(sp, orc, ctx):
    """FRIProof.fri_layers by value (fri_commit.rs:117-121): the layers a *_to_host commit streams to host memory on the
    copy stream equal the oracle's layers and what stark_fri_layer_read returns; same transcript as the resident commit;
    pageable and pinned destinations, the step API, a destination that is too small."""
    import torch
    for log_n, log_deg, pinned in ((12, 9, False), (16, 13, True), (20, 17, True), (5, 5, False), (4, 0, False)):
        c = orc.synthetic_poly_exact_degree(77 + log_n, 1 << log_deg)
        cap = 2 << log_n
        keep = torch.empty(cap, dtype=torch.int64).pin_memory() if pinned else None
        buf = keep.numpy().view(np.uint64) if pinned else np.empty(cap, dtype=np.uint64)
        buf[:] = np.uint64(0xDEADBEEFDEADBEEF)
        ch, ch0, och = sp.Channel(P), sp.Channel(P), orc.Channel(P)
        pr = sp.fri_commit(ctx, c, sp.CosetFri(ctx, 5, log_n), ch, layers_out=buf)
        p0 = sp.fri_commit(ctx, c, sp.CosetFri(ctx, 5, log_n), ch0)
        opr = orc.fri_commit_fast(c, log_n, 5, orc.root_of_unity(log_n), och, P)
        assert ch.state == ch0.state == och.state and pr.num_layers == opr.num_layers
        off = 0
        for k in range(pr.num_layers):
            ln = pr.layer_len(k)
            assert pr.layer_host_offset(k) == off
            assert np.array_equal(buf[off:off + ln], opr.layer(k)), f"log_n {log_n} layer {k}"
            assert np.array_equal(buf[off:off + ln], pr.layer(k))
            off += ln
        assert pr.layer_host_offset(pr.num_layers) == -1 and p0.layer_host_offset(0) == -1
        assert np.all(buf[off:] == np.uint64(0xDEADBEEFDEADBEEF))
        sp.decommit_fri(3, (1 << log_n) - 1, pr, ch)
        orc.decommit_fri(3, (1 << log_n) - 1, opr, och)
        assert ch.state == och.state
        pr.free(); p0.free()
    # the asynchronous form: the transcript is complete on return, the copies after layers_wait (here: under the openings)
    log_n = 18
    c = orc.synthetic_poly_exact_degree(91, 1 << 15)
    keep = torch.empty(2 << log_n, dtype=torch.int64).pin_memory()
    buf = keep.numpy().view(np.uint64)
    ch, ch0 = sp.Channel(P), sp.Channel(P)
    pr = sp.fri_commit(ctx, c, sp.CosetFri(ctx, 5, log_n), ch, layers_out=buf, wait=False)
    sp.decommit_fri(4, (1 << log_n) - 1, pr, ch)
    pr.layers_wait()
    p0 = sp.fri_commit(ctx, c, sp.CosetFri(ctx, 5, log_n), ch0)
    sp.decommit_fri(4, (1 << log_n) - 1, p0, ch0)
    assert ch.state == ch0.state
    for k in range(p0.num_layers):
        o = pr.layer_host_offset(k)
        assert np.array_equal(buf[o:o + p0.layer_len(k)], p0.layer(k)), f"async layer {k}"
    pr.free(); p0.free()
    # step API: the folds keep streaming; layers_wait before reading
    log_n = 10
    c = orc.synthetic_poly_exact_degree(5, 1 << 7)
    buf = np.zeros(2 << log_n, dtype=np.uint64)
    pr, root = sp.fri_begin(ctx, c, log_n, 5, layers_out=buf)
    betas = [3, 1234567, 99]
    for b in betas:
        pr.fold(b)
    pr.layers_wait()
    ev = orc.coset_evaluate(c, log_n, 5, orc.root_of_unity(log_n), P)
    assert np.array_equal(buf[:1 << log_n], ev)
    off = 0
    for k in range(pr.num_layers):
        assert np.array_equal(buf[off:off + pr.layer_len(k)], pr.layer(k))
        off += pr.layer_len(k)
    pr.free()
    # a destination that cannot hold layer 0 is refused up front; one that fills up later fails in the fold that overflows it
    with pytest.raises(sp.StarkError):
        sp.fri_begin(ctx, c, log_n, 5, layers_out=np.zeros((1 << log_n) - 1, dtype=np.uint64))
    small = np.zeros((1 << log_n) + 100, dtype=np.uint64)
    with pytest.raises(sp.StarkError):
        sp.fri_commit(ctx, c, sp.CosetFri(ctx, 5, log_n), sp.Channel(P), layers_out=small)
    # the context is still usable afterwards
    _check_fri(sp, orc, ctx, c, log_n, 5, 2)


def test_fri_too_many_folds_is_error(sp, orc, ctx):
    # degree >= domain size: the reference would reach an empty layer and panic in MerkleTree::root()
    pr, _ = sp.fri_begin(ctx, orc.synthetic_poly_exact_degree(1, 4), 2, 5)
    pr.fold(3); pr.fold(5)
    with pytest.raises(sp.StarkError):
        pr.fold(7)


# ---------------------------------------------------------------- cfg1: the STARK-101 transcript
def test_stark101_transcript(sp, orc, ctx, golden):
    s = golden["transcripts"]["stark101"]
    ch = sp.Channel(P)
    sp.stark101_prove(ctx, ch, s["a1"], s["log_trace"], s["log_blowup"], s["queries"])
    och = orc.Channel(P)
    orc.stark101_prove(och, literal=True)
    assert ch.proof == och.proof and ch.state == och.state
    assert ch.state == s["final_state"] and ch.proof_size() == s["proof_size"]
    assert ch.compressed_proof_size() == s["compressed_proof_size"]
    assert hashlib.sha256(ch.proof_flat()).hexdigest() == s["proof_sha256"]


@pytest.mark.parametrize("log_trace,log_blowup,a1,q", [(5, 2, 7, 2), (8, 3, 3141592, 3), (12, 3, 99, 4), (14, 4, 5, 2)])
def test_stark101_other_sizes(sp, orc, ctx, log_trace, log_blowup, a1, q):
    ch, och = sp.Channel(P), orc.Channel(P)
    sp.stark101_prove(ctx, ch, a1, log_trace, log_blowup, q)
    orc.stark101_prove(och, a1=a1, log_trace=log_trace, log_blowup=log_blowup, num_queries=q, literal=False)
    assert ch.proof == och.proof and ch.state == och.state


@pytest.mark.parametrize("log_trace", [18, 20])
def test_stark101_large_domains_vs_oracle(sp, orc, ctx, log_trace):
    """the whole prover against the oracle's NTT tier at 2^21 and 2^23 domains (0.8 s / 3 s of CPU), message by
    message, and the verifier accepts the GPU proof."""
    q = 3
    ch, och = sp.Channel(P), orc.Channel(P)
    sp.stark101_prove(ctx, ch, 3141592, log_trace, 3, q)
    orc.stark101_prove(och, a1=3141592, log_trace=log_trace, log_blowup=3, num_queries=q, literal=False)
    assert ch.state == och.state and ch.proof == och.proof
    claimed = int.from_bytes(ch.proof[0][40:48], "big")
    ok, why = sp.stark101_verify(ch.proof_flat(), claimed, log_trace, 3, q)
    assert ok, why


# ---------------------------------------------------------------- other fields
# moduli that stress the arithmetic: just below 2^32 (every add can wrap), just above 2^31, tiny, low two-adicity
@pytest.mark.parametrize("modulus,gen", [(998244353, 3), (2013265921, 31), (17, 3), (7, 3), (257, 3),
                                         (4293918721, 19), (4294967291, 2), (3489660929, 3), (2281701377, 3), (469762049, 3)])
def test_other_moduli(sp, orc, modulus, gen):
    c = sp.Context(modulus, gen, 0)
    try:
        adic = c.two_adicity
        log_n = min(adic, 16)
        w = orc.root_of_unity(log_n, modulus, gen)
        assert c.root_of_unity(log_n) == w
        coeffs = orc.synthetic_column(1, 1 << max(log_n - 2, 0), modulus)
        off = gen
        assert np.array_equal(c.coset_evaluate(coeffs, log_n, off), orc.coset_evaluate(coeffs, log_n, off, w, modulus))
        a = orc.synthetic_column(2, 500, modulus)
        assert np.array_equal(c.batch_inverse(a), orc.batch_inverse(a, modulus))
        ch, och = sp.Channel(modulus), orc.Channel(modulus)
        pr = sp.fri_commit(c, coeffs, sp.CosetFri(c, off, log_n), ch)
        opr = orc.fri_commit_fast(coeffs, log_n, off, w, och, modulus)
        sp.decommit_fri(2, (1 << log_n) - 1, pr, ch)
        orc.decommit_fri(2, (1 << log_n) - 1, opr, och)
        assert ch.proof == och.proof and ch.state == och.state
    finally:
        c.close()


# ---------------------------------------------------------------- BASELINE.json's full sizes
def test_cfg2_lde_commit_2e20(sp, orc, ctx):
    """configs[1]: single-column coset LDE (2^17 coefficients, seed 42) + Merkle commit at a 2^20 domain."""
    log_n = 20
    c = orc.synthetic_column(42, 1 << 17)
    ev = ctx.coset_evaluate(c, log_n, 5)
    want = orc.coset_evaluate(c, log_n, 5, orc.root_of_unity(log_n), P)
    assert np.array_equal(ev, want)
    assert sp.MerkleTree.new(ctx, ev).root_bytes() == orc.merkle_root_only(want)


@pytest.mark.timeout(600)
def test_cfg3_fri_commit_2e24_exact(sp, orc, ctx):
    """configs[2] at full size against the oracle's NTT tier: every root, the final constant, and the
    transcript after 32 query openings, bit for bit; plus size-independent properties on the layers."""
    log_n, log_deg, q = 24, 21, 32
    c = orc.synthetic_poly_exact_degree(43, 1 << log_deg)
    ch, och = sp.Channel(P), orc.Channel(P)
    pr = sp.fri_commit(ctx, c, sp.CosetFri(ctx, 5, log_n), ch)
    opr = orc.fri_commit_fast(c, log_n, 5, orc.root_of_unity(log_n), och, P)
    assert pr.num_layers == opr.num_layers == 22
    assert sum(pr.layer_len(k) for k in range(22)) == 33554424
    for k in range(22):
        assert pr.tree(k).root() == opr.tree(k).root_hex(), k
    assert ch.state == och.state
    assert np.array_equal(pr.final_poly(), opr.final_poly())
    # sampled rows of every layer
    rng = np.random.default_rng(7)
    for k in range(22):
        n = pr.layer_len(k)
        off = int(rng.integers(0, max(n - 64, 1)))
        assert np.array_equal(pr.layer(k, off, min(64, n)), opr.layer(k)[off:off + min(64, n)]), k
    # the last layer is the constant polynomial
    assert set(pr.layer(21).tolist()) == {int(pr.final_poly()[0])}
    sp.decommit_fri(q, (1 << log_n) - 1, pr, ch)
    orc.decommit_fri(q, (1 << log_n) - 1, opr, och)
    assert ch.state == och.state and ch.proof_size() == och.proof_size()
    assert hashlib.sha256(ch.proof_flat()).hexdigest() == hashlib.sha256(och.proof_flat()).hexdigest()


def test_fold_property_at_full_size(sp, orc, ctx):
    """encode -> fold consistency without the oracle's big arrays: e'[i] recomputed from e[i], e[i+n/2]
    with the reference's field operators at sampled indices of a 2^22 layer."""
    log_n = 22
    c = orc.synthetic_poly_exact_degree(11, 1 << 19)
    pr, _ = sp.fri_begin(ctx, c, log_n, 5)
    beta = 123456789
    pr.fold(beta)
    n, w = 1 << log_n, orc.root_of_unity(log_n)
    inv2 = orc.fe_inverse(2, P)
    rng = np.random.default_rng(3)
    for i in [0, 1, n // 2 - 1] + [int(x) for x in rng.integers(0, n // 2, 20)]:
        a, b = int(pr.layer(0, i, 1)[0]), int(pr.layer(0, i + n // 2, 1)[0])
        d = orc.fe_mul(5, orc.fe_pow(w, i, P), P)
        want = orc.fe_add(orc.fe_mul(orc.fe_add(a, b, P), inv2, P),
                          orc.fe_mul(orc.fe_mul(beta, orc.fe_sub(a, b, P), P), orc.fe_inverse(orc.fe_mul(2, d, P), P), P), P)
        assert int(pr.layer(1, i, 1)[0]) == want, i
    # openings of the folded layer verify against its root
    t = pr.tree(1)
    for idx in (0, 12345, n // 2 - 1):
        assert orc.merkle_verify(t.root_bytes(), n // 2, idx, int(pr.layer(1, idx, 1)[0]), t.get_authentication_path(idx))


def test_fri_maximum_domain_2e30_verifies(sp, orc):
    """The largest FRI domain this field admits (2^30 points, 28 layers, 2^31 leaves hashed): the verifier
    (Merkle paths of every layer + fold consistency, fri_verify.rs:12-177 completed) accepts the transcript, and a
    flipped bit is rejected.  Size-independent property; the CPU oracle does not go this far."""
    c2 = sp.Context()
    try:
        log_n, q = 30, 4
        coeffs = orc.synthetic_poly_exact_degree(30, 1 << (log_n - 3))
        ch = sp.Channel(P)
        pr = sp.fri_commit(c2, coeffs, sp.CosetFri(c2, 5, log_n), ch)
        assert pr.num_layers == 28 and pr.layer_len(0) == 1 << log_n and pr.layer_len(27) == 8
        sp.decommit_fri(q, (1 << log_n) - 1, pr, ch)
        flat = ch.proof_flat()
        ok, why = sp.verify_fri(flat, log_n, 5, q, (1 << log_n) - 1, log_n - 3)
        assert ok, why
        bad = bytearray(flat)
        bad[len(bad) // 2] ^= 0x10
        ok2, _ = sp.verify_fri(bytes(bad), log_n, 5, q, (1 << log_n) - 1, log_n - 3)
        assert not ok2
        # layer 0 against the reference's literal Horner evaluation (ops.rs:76-83) at a few points of the coset
        w = orc.root_of_unity(log_n)
        for i in (1, (1 << 29) + 3):                 # 2^27 Horner steps each on the host
            x = 5 * pow(w, i, P) % P
            assert int(pr.layer(0, i, 1)[0]) == orc.poly_evaluate(coeffs, x, P), i
        pr.free()
    finally:
        c2.close()


def test_lde_roundtrip_2e24(sp, orc, ctx):
    """interpolate(evaluate(c)) == c at 2^24 (idempotence), device resident."""
    log_n = 24
    c = orc.synthetic_column(5, 1 << log_n)
    v = ctx.upload(c)
    ev = ctx.coset_evaluate_dev(v, log_n, 5)
    back = ctx.coset_interpolate_dev(ev, 5)
    assert np.array_equal(back.download(), c)
    lde = ctx.coset_lde_dev(ev, 5, 1, 7)           # same polynomial on a 2x larger coset
    assert int(lde.download(0, 1)[0]) == orc.poly_evaluate(c, 7, P)    # the reference's Horner at the first coset point (2^24 steps)


# ---------------------------------------------------------------- error behaviour at the boundary
def test_argument_errors(sp, orc, ctx):
    v = ctx.upload(np.arange(10, dtype=np.uint64))
    with pytest.raises(sp.StarkError) as e:
        v.download(5, 10)                                    # range out of bounds
    assert e.value.code == 1
    with pytest.raises(sp.StarkError):
        ctx.coset_interpolate_dev(v, 5)                      # length is not a power of two
    t = sp.MerkleTree.new(ctx, np.arange(10, dtype=np.uint64))
    with pytest.raises(sp.StarkError):
        t.get_authentication_path(10)                        # no such leaf
    with pytest.raises(sp.StarkError):
        t.node(0, 0)                                         # level-0 digests are not stored
    other = sp.Context(998244353, 3, 0)
    try:
        with pytest.raises(sp.StarkError):
            other.batch_inverse_dev(v)                       # vector belongs to another context
    finally:
        other.close()
    # the library is still healthy after the errors
    assert sp.MerkleTree.new(ctx, [0, 1, 2]).root() == "07f15c470799d408152313c5b0e914969f00e954b6df62e826e235bd7d06d424"


def test_two_contexts_interleaved(sp, orc):
    """Contexts are independent (own stream, own reduction scratch and barrier counter)."""
    a, b = sp.Context(P, 5, 0), sp.Context(P, 5, 0)
    try:
        c1, c2 = orc.synthetic_poly_exact_degree(1, 1 << 12), orc.synthetic_poly_exact_degree(2, 1 << 13)
        p1, r1 = sp.fri_begin(a, c1, 15, 5)
        p2, r2 = sp.fri_begin(b, c2, 16, 5)
        for beta in (3, 5, 7, 11):
            p1.fold(beta); p2.fold(beta + 1)
        for pr, c, log_n, betas in ((p1, c1, 15, (3, 5, 7, 11)), (p2, c2, 16, (4, 6, 8, 12))):
            e = orc.coset_evaluate(c, log_n, 5, orc.root_of_unity(log_n), P)
            off, w = 5, orc.root_of_unity(log_n)
            for k, beta in enumerate(betas):
                e = orc.fri_fold_evals(e, beta, off, w, P)
                off, w = off * off % P, w * w % P
                assert np.array_equal(pr.layer(k + 1), e)
                assert pr.tree(k + 1).root_bytes() == orc.merkle_root_only(e)
    finally:
        a.close(); b.close()


@pytest.mark.parametrize("modulus,gen", [(2013265921, 31), (998244353, 3), (469762049, 3)])
def test_small_modulus_blowup8_path(sp, orc, modulus, gen):
    """p < 2^31: the blow-up-by-8 transform must not use lazily reduced ("weak") butterfly values, which need
    2^32 < 2p (found by tools/soak.py: evaluations were off by multiples of p folded wrongly)."""
    c = sp.Context(modulus, gen, 0)
    try:
        for log_n, nco in ((13, 14), (14, 31), (16, 602), (18, 1 << 15)):
            w = orc.root_of_unity(log_n, modulus, gen)
            coeffs = orc.synthetic_column(5 + log_n, nco, modulus)
            off = 720954169 % modulus
            assert np.array_equal(c.coset_evaluate(coeffs, log_n, off), orc.coset_evaluate(coeffs, log_n, off, w, modulus)), log_n
        ch, och = sp.Channel(modulus), orc.Channel(modulus)
        coeffs = orc.synthetic_column(9, 2731, modulus)
        pr = sp.fri_commit(c, coeffs, sp.CosetFri(c, 267101690 % modulus, 15), ch)
        sp.decommit_fri(2, (1 << 15) - 1, pr, ch)
        opr = orc.fri_commit_fast(coeffs, 15, 267101690 % modulus, orc.root_of_unity(15, modulus, gen), och, modulus)
        orc.decommit_fri(2, (1 << 15) - 1, opr, och)
        assert ch.state == och.state and ch.proof == och.proof
    finally:
        c.close()


@pytest.mark.parametrize("modulus,gen", [(3221225473, 5), (4293918721, 19), (998244353, 3)])
def test_blowup8_first_pass_ragged_coefficients(sp, orc, modulus, gen):
    """First pass of the blow-up-by-8 transform: a tile takes the lowest row digit x the TOP two row bits, so one thread
    reads FOUR adjacent coefficients with one 16-byte load and walks their scale / shift factors.  Coefficient counts
    around every multiple of 4, at the ends of the tables, below one tile row and above -- for the reference field (the
    kernels instantiated with the modulus as an immediate), a modulus close to 2^32 (run-time modulus, weak values) and one
    below 2^31 (strict values)."""
    c = sp.Context(modulus, gen, 0)
    try:
        for log_n in (13, 14, 17):
            w = orc.root_of_unity(log_n, modulus, gen)
            top = 1 << (log_n - 3)
            for nco in (1, 2, 3, 4, 5, 6, 7, 8, 9, 127, 129, 1021, 1023, 1024, top - 3, top - 1, top):
                coeffs = orc.synthetic_column(nco + log_n, nco, modulus)
                off = (5 + 977 * nco) % modulus or 7
                got = c.coset_evaluate(coeffs, log_n, off)
                assert np.array_equal(got, orc.coset_evaluate(coeffs, log_n, off, w, modulus)), (log_n, nco)
        # the LDE chain (natural iNTT -> first pass reads its unscaled coefficients with c0 = 1/n folded into the walk)
        for log_n in (10, 12, 15):
            vals = orc.synthetic_column(77 + log_n, 1 << log_n, modulus)
            w, W = orc.root_of_unity(log_n, modulus, gen), orc.root_of_unity(log_n + 3, modulus, gen)
            co = orc.coset_interpolate(vals, log_n, 1, w, modulus)
            want = orc.coset_evaluate(co, log_n + 3, 11, W, modulus)
            assert np.array_equal(c.coset_lde(vals, log_n, 1, 3, 11), want), log_n
    finally:
        c.close()


@pytest.mark.parametrize("modulus,gen", [(4293918721, 19), (3489660929, 3)])
def test_large_modulus_lde_and_fold_paths(sp, orc, modulus, gen):
    """The blow-up-8 LDE kernel (weak/lazy arithmetic) and the fused fold with p close to 2^32."""
    c = sp.Context(modulus, gen, 0)
    try:
        log_t, log_n = 13, 16
        w_t, w_n = orc.root_of_unity(log_t, modulus, gen), orc.root_of_unity(log_n, modulus, gen)
        col = orc.synthetic_column(3, 1 << log_t, modulus)
        col[:8] = [modulus - 1, modulus - 2, 0, 1, modulus - 1, 0, modulus - 1, modulus - 1]
        coef = orc.coset_interpolate(col, log_t, 1, w_t, modulus)
        want = orc.coset_evaluate(coef, log_n, gen, w_n, modulus)
        assert np.array_equal(c.coset_lde(col, log_t, 1, 3, gen), want)
        assert np.array_equal(c.coset_evaluate(coef, log_n, gen), want)
        pr, _ = sp.fri_begin(c, coef, log_n, gen)
        beta = modulus - 1
        pr.fold(beta)
        assert np.array_equal(pr.layer(1), orc.fri_fold_evals(want, beta, gen, w_n, modulus))
    finally:
        c.close()


def test_context_destroyed_before_its_handles(sp, orc):
    """a C or Rust caller may drop the context first.  stark_ctx_destroy with handles outstanding only marks
    the context; the last handle released tears it down -- no use-after-free, and the handles stay usable until then."""
    import ctypes as C
    L = sp.lib()
    h = C.c_void_p()
    assert L.stark_ctx_create(P, 5, 0, C.byref(h)) == 0
    vals = orc.synthetic_column(3, 1 << 12)
    v, t = C.c_void_p(), C.c_void_p()
    assert L.stark_vec_upload(h, vals.ctypes.data_as(C.c_void_p), vals.size, C.byref(v)) == 0
    assert L.stark_merkle_commit_dev(h, v, C.byref(t)) == 0
    L.stark_ctx_destroy(h)                                   # handles outstanding: deferred
    out = np.zeros(16, dtype=np.uint64)
    assert L.stark_vec_download(v, 0, 16, out.ctypes.data_as(C.c_void_p)) == 0 and np.array_equal(out, vals[:16])
    root = np.zeros(32, dtype=np.uint8)
    assert L.stark_merkle_root(t, root.ctypes.data_as(C.c_void_p)) == 0 and root.tobytes() == orc.merkle_root_only(vals)
    L.stark_vec_destroy(v)
    L.stark_tree_destroy(t)                                  # last handle: the context goes with it


def test_oversized_log_n_is_rejected_before_any_read(sp, ctx):
    """stark_ntt / stark_intt / stark_coset_interpolate sized and uploaded before checking log_n."""
    import ctypes as C
    L = sp.lib()
    tiny = np.zeros(4, dtype=np.uint64)
    for fn in (L.stark_ntt, L.stark_intt):
        for bad in (31, 40, 64, 200):
            assert fn(ctx.h, tiny.ctypes.data_as(C.c_void_p), bad) != 0
    assert L.stark_coset_interpolate(ctx.h, tiny.ctypes.data_as(C.c_void_p), 31, 5, tiny.ctypes.data_as(C.c_void_p)) != 0
    v = ctx.upload(np.arange(64, dtype=np.uint64))
    with pytest.raises(sp.StarkError):
        ctx.pow_mul_dev(v, 8, 0, True, 0, 5, 1, 4)         # exponents up to 7*7 = 49 >= 2^4
    ctx.pow_mul_dev(v, 8, 0, True, 0, 5, 1, 6)             # 49 < 2^6: fine
