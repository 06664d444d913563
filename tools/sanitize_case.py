#!/usr/bin/env python
"""Smallest end-to-end case for compute-sanitizer: exercises every kernel once at small sizes."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
sp = importlib.import_module("stark-prover_b200")
from oracle import pyoracle as orc
P = sp.P_DEFAULT
ctx = sp.Context()
for n in (1, 3, 9, 100, 513, 4097, 70000):
    v = orc.synthetic_column(n, n)
    t = sp.MerkleTree.new(ctx, v)
    assert t.root() == orc.Tree(v).root_hex(), n
    assert t.get_authentication_path(n - 1) == orc.Tree(v).path(n - 1)
for log_n in (0, 3, 9, 10, 14, 17):
    a = orc.synthetic_column(log_n, 1 << log_n)
    assert np.array_equal(ctx.intt(ctx.ntt(a, log_n), log_n), a)
e = orc.synthetic_column(5, 1 << 10)
ctx.coset_lde(e, 10, 1, 3, 5)
a = orc.synthetic_column(6, 5000); a[7] = 0
assert np.array_equal(ctx.batch_inverse(a), orc.batch_inverse(a, P))
for log_n, log_deg in ((6, 6), (13, 10), (17, 14)):
    c = orc.synthetic_poly_exact_degree(log_n, 1 << log_deg)
    ch, och = sp.Channel(P), orc.Channel(P)
    pr = sp.fri_commit(ctx, c, sp.CosetFri(ctx, 5, log_n), ch)
    sp.decommit_fri(2, (1 << log_n) - 1, pr, ch)
    opr = orc.fri_commit_fast(c, log_n, 5, orc.root_of_unity(log_n), och, P)
    orc.decommit_fri(2, (1 << log_n) - 1, opr, och)
    assert ch.state == och.state
ch, och = sp.Channel(P), orc.Channel(P)
sp.stark101_prove(ctx, ch)
orc.stark101_prove(och)
assert ch.state == och.state
v = ctx.upload(orc.synthetic_column(1, 1 << 12))
ctx.ntt_batch_dev(v, 6)
ctx.close()
print("sanitize case ok")
