#!/usr/bin/env python
"""Randomised differential soak of the HOST side (no GPU needed): the library's Channel against the C oracle's and the
hashlib restatement over random send / receive sequences, Merkle verification of oracle paths with random corruptions,
and the FRI / STARK verifiers on oracle transcripts with random bit flips.   python tools/soak_host.py [seconds] [seed]"""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
sp = importlib.import_module("stark-prover_b200")
from oracle import pyoracle as orc

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
P = sp.P_DEFAULT
counts, fails = {}, 0
def check(kind, ok, detail):
    global fails
    counts[kind] = counts.get(kind, 0) + 1
    if not ok:
        fails += 1
        print(f"MISMATCH {kind}: {detail}", flush=True)

def flat(msgs): return b"".join(len(m).to_bytes(4, "little") + bytes(m) for m in msgs)

t_end = time.time() + budget
while time.time() < t_end:
    kind = int(rng.integers(0, 4))
    if kind == 0:                                   # channel op sequences
        m = int(rng.choice([P, 2013265921, 17, 4294967291]))
        a, b, c = sp.Channel(m), orc.Channel(m), orc.PyChannel(m)
        ok, log = True, []
        a.send(b"seed"); b.send(b"seed"); c.send(b"seed")
        for _ in range(int(rng.integers(1, 40))):
            op = int(rng.integers(0, 3))
            if op == 0:
                n = int(rng.choice([0, 1, 8, 31, 32, 33, 64, 511, 512, 513, 1023, 1024, 1025, 2047, 2048, 2049, int(rng.integers(0, 6000))]))
                msg = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
                a.send(msg); b.send(msg); c.send(msg); log.append(("send", n))
            elif op == 1:
                x, y, z = a.receive_random_field_element(), b.receive_random_field_element(), c.receive_random_field_element()
                ok = ok and x == y == z; log.append(("rfe",))
            else:
                lo = int(rng.integers(0, 1 << int(rng.integers(1, 62))))
                hi = lo + int(rng.integers(0, 1 << int(rng.integers(1, 62))))
                show = bool(rng.integers(0, 2))
                x, y, z = a.receive_random_int(lo, hi, show), b.receive_random_int(lo, hi, show), c.receive_random_int(lo, hi, show)
                ok = ok and x == y == z and 0 <= x <= hi - lo; log.append(("rri", lo, hi, show))   # channel.rs:72: (state + min) % range, not shifted back
            ok = ok and a.state == b.state == c.state
        ok = ok and a.proof == b.proof == c.proof and a.proof_size() == b.proof_size() == c.proof_size()
        ok = ok and a.compressed_proof_size() == b.compressed_proof_size() and a.compressed_proof == b.compressed_proof
        check("channel", ok, f"modulus={m} ops={log[:12]}...")
    elif kind == 1:                                 # Merkle paths: accept the real one, reject any single-bit change
        n = int(rng.integers(1, 3000))
        vals = orc.synthetic_column(int(rng.integers(1, 1 << 30)), n)
        t = orc.Tree(vals)
        idx = int(rng.integers(0, n))
        path, root = t.path(idx), t.root()
        ok = sp.merkle_validate(root, n, idx, int(vals[idx]), path)
        if path:
            bad = bytearray(path); bad[int(rng.integers(0, len(bad)))] ^= 1 << int(rng.integers(0, 8))
            ok = ok and not sp.merkle_validate(root, n, idx, int(vals[idx]), bytes(bad))
            ok = ok and not sp.merkle_validate(root, n, idx, int(vals[idx]), path[:-32])
        ok = ok and not sp.merkle_validate(root, n, idx, (int(vals[idx]) + 1) % P, path)
        if n > 1:
            ok = ok and not sp.merkle_validate(root, n, (idx + 1) % n, int(vals[idx]), path) or int(vals[(idx + 1) % n]) == int(vals[idx])
        check("merkle_validate", ok, f"n={n} idx={idx}")
    elif kind == 2:                                 # FRI verifier on oracle transcripts + one random bit flip
        log_n = int(rng.integers(1, 11)); log_d = int(rng.integers(0, log_n + 1))
        nco = int(rng.integers(1, (1 << log_d) + 1)); off = int(rng.integers(1, P)); q = int(rng.integers(1, 4))
        c = orc.synthetic_poly_exact_degree(int(rng.integers(1, 1 << 30)), nco)
        ch = orc.Channel(P)
        pr = orc.fri_commit_fast(c, log_n, off, orc.root_of_unity(log_n), ch, P)
        orc.decommit_fri(q, (1 << log_n) - 1, pr, ch)
        msgs = ch.proof
        good, why = sp.verify_fri(flat(msgs), log_n, off, q, (1 << log_n) - 1, log_d)
        ok = good
        k = int(rng.integers(0, len(msgs)))
        if len(msgs[k]):
            bad = [bytearray(x) for x in msgs]
            bad[k][int(rng.integers(0, len(bad[k])))] ^= 1 << int(rng.integers(0, 8))
            rej, why2 = sp.verify_fri(flat(bad), log_n, off, q, (1 << log_n) - 1, log_d)
            ok = ok and not rej
        check("verify_fri", ok, f"log_n={log_n} coeffs={nco} off={off} q={q} why={why!r} flipped message {k}")
    else:                                           # STARK verifier on oracle transcripts + bit flip + wrong claim
        log_t, log_b = int(rng.integers(2, 9)), int(rng.integers(1, 4))
        a1, q = int(rng.integers(0, P)), int(rng.integers(1, 4))
        ch = orc.Channel(P)
        orc.stark101_prove(ch, a1, log_t, log_b, sp.G_DEFAULT, q, literal=False)
        claimed = int(orc.fibsq_trace(a1, (1 << log_t) - 1)[(1 << log_t) - 2])
        msgs = ch.proof
        good, why = sp.stark101_verify(flat(msgs), claimed, log_t, log_b, q)
        ok = good and not sp.stark101_verify(flat(msgs), (claimed + 1) % P, log_t, log_b, q)[0]
        k = int(rng.integers(0, len(msgs)))
        bad = [bytearray(x) for x in msgs]
        bad[k][int(rng.integers(0, len(bad[k])))] ^= 1 << int(rng.integers(0, 8))
        ok = ok and not sp.stark101_verify(flat(bad), claimed, log_t, log_b, q)[0]
        check("stark101_verify", ok, f"log_trace={log_t} log_blowup={log_b} a1={a1} q={q} why={why!r} flipped message {k}")
print(f"soak_host: {sum(counts.values())} cases in {budget:.0f} s {counts}, mismatches: {fails} (seed {seed})")
sys.exit(1 if fails else 0)
