#!/usr/bin/env python
"""Randomised differential soak against the CPU oracle (test infrastructure, not product):
    python tools/soak.py [seconds] [seed] [threads]
Random tree sizes, transform sizes / offsets, polynomial degrees, inverse inputs with zero patterns and a few moduli;
every result must equal the oracle's bit for bit.  Prints one line per failure and a summary; exit code 1 on any mismatch."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
sp = importlib.import_module("stark-prover_b200")
from oracle import pyoracle as orc

import threading

MODULI = [(3221225473, 5), (2013265921, 31), (998244353, 3), (4293918721, 19), (469762049, 3)]
lock = threading.Lock()
counts, fails = {}, [0]


def run(budget: float, seed: int, tag: str = "") -> None:
    """One worker: its own contexts (one per modulus), its own random stream."""
    rng = np.random.default_rng(seed)
    ctxs = {}

    def ctx_for(m, g):
        if m not in ctxs:
            ctxs[m] = sp.Context(m, g, 0)
        return ctxs[m]

    def check(kind, ok, detail):
        with lock:
            counts[kind] = counts.get(kind, 0) + 1
            if not ok:
                fails[0] += 1
                print(f"MISMATCH {tag}{kind}: {detail}", flush=True)

    t_end = time.time() + budget
    while time.time() < t_end:
        m, g = MODULI[int(rng.integers(0, len(MODULI)))]
        ctx = ctx_for(m, g)
        two = (m - 1 & -(m - 1)).bit_length() - 1
        kind = int(rng.integers(0, 14))
        if kind == 0:                                   # Merkle: ragged sizes, every level's ends + random nodes + paths
            n = int(rng.integers(1, 1 << int(rng.integers(1, 19))))
            vals = orc.synthetic_column(int(rng.integers(1, 1 << 30)), n, m)
            t, ot = sp.MerkleTree.new(ctx, vals), orc.Tree(vals)
            ok = t.root() == ot.root_hex()
            for l in range(1, t.depth + 1):
                w = -(-n // (1 << l))
                for j in {0, w - 1, int(rng.integers(0, w))}:
                    ok = ok and t.node(l, j) == ot.node(l, j)
            for idx in {0, n - 1, int(rng.integers(0, n))}:
                ok = ok and t.get_authentication_path(idx) == ot.path(idx)
            check("merkle", ok, f"n={n} modulus={m}")
            t.free()
        elif kind == 1:                                 # coset evaluation / interpolation round trip against the oracle
            log_n = int(rng.integers(0, min(two, 18) + 1))
            log_d = int(rng.integers(0, log_n + 1))
            nco = int(rng.integers(1, (1 << log_d) + 1))
            off = int(rng.integers(1, m))
            w = orc.root_of_unity(log_n, m, g)
            c = orc.synthetic_column(int(rng.integers(1, 1 << 30)), nco, m)
            ev = ctx.coset_evaluate(c, log_n, off)
            want = orc.coset_evaluate(c, log_n, off, w, m)
            back = ctx.coset_interpolate(ev, log_n, off)
            ok = np.array_equal(ev, want) and np.array_equal(back[:nco], c) and not back[nco:].any()
            check("coset_ntt", ok, f"log_n={log_n} coeffs={nco} offset={off} modulus={m}")
        elif kind == 2:                                 # batched inverse / quotient with zero patterns
            n = int(rng.integers(1, 1 << int(rng.integers(1, 17))))
            a = orc.synthetic_column(int(rng.integers(1, 1 << 30)), n, m)
            z = rng.random(n) < float(rng.choice([0.0, 0.01, 0.5, 0.99]))
            a[z] = 0
            got = ctx.batch_inverse(a)
            ok = np.array_equal(got, orc.batch_inverse(a, m))
            num = orc.synthetic_column(int(rng.integers(1, 1 << 30)), n, m)
            q = ctx.quotient_pointwise(num, a)
            inv = orc.batch_inverse(a, m)
            want = np.array([int(x) * int(y) % m for x, y in zip(num[:2000], inv[:2000])], dtype=np.uint64)
            ok = ok and np.array_equal(q[:2000], want)
            check("inverse", ok, f"n={n} zeros={int(z.sum())} modulus={m}")
        elif kind == 3:                                 # FRI commit + openings: whole transcript
            log_n = int(rng.integers(1, min(two, 15) + 1))
            log_d = int(rng.integers(0, log_n + 1))
            nco = int(rng.integers(1, (1 << log_d) + 1))
            off = int(rng.integers(1, m))
            c = orc.synthetic_column(int(rng.integers(1, 1 << 30)), nco, m)
            if rng.random() < 0.3:
                c[-int(rng.integers(1, nco + 1)):] = 0  # trailing zeros: Polynomial::new trims (may become the zero polynomial)
            q = int(rng.integers(0, 4))
            ch, och = sp.Channel(m), orc.Channel(m)
            try:
                pr = sp.fri_commit(ctx, c, sp.CosetFri(ctx, off, log_n), ch)
                sp.decommit_fri(q, (1 << log_n) - 1, pr, ch)
                gpu_err = None
            except sp.StarkError as e:
                gpu_err = e
            try:
                opr = orc.fri_commit_fast(c, log_n, off, orc.root_of_unity(log_n, m, g), och, m)
                orc.decommit_fri(q, (1 << log_n) - 1, opr, och)
                cpu_err = None
            except Exception as e:
                cpu_err = e
            ok = (gpu_err is None) == (cpu_err is None) and (gpu_err is not None or (ch.state == och.state and ch.proof == och.proof))
            check("fri", ok, f"log_n={log_n} coeffs={nco} offset={off} q={q} modulus={m} gpu_err={gpu_err} cpu_err={cpu_err}")
        elif kind == 5:                                 # LDE of evaluations: blow-ups 1..16 (8 takes its own kernel), both offsets random
            log_t = int(rng.integers(0, min(two, 17) + 1))
            log_b = int(rng.integers(0, min(4, two - log_t) + 1))
            off_in, off_out = int(rng.integers(1, m)), int(rng.integers(1, m))
            ev = orc.synthetic_column(int(rng.integers(1, 1 << 30)), 1 << log_t, m)
            got = ctx.coset_lde(ev, log_t, off_in, log_b, off_out)
            coef = orc.coset_interpolate(ev, log_t, off_in, orc.root_of_unity(log_t, m, g), m)
            want = orc.coset_evaluate(coef, log_t + log_b, off_out, orc.root_of_unity(log_t + log_b, m, g), m)
            check("coset_lde", np.array_equal(got, want), f"log_t={log_t} log_blowup={log_b} off_in={off_in} off_out={off_out} modulus={m}")
        elif kind == 6:                                 # natural-order NTT / iNTT and the coset domain
            log_n = int(rng.integers(0, min(two, 18) + 1))
            w = orc.root_of_unity(log_n, m, g)
            a = orc.synthetic_column(int(rng.integers(1, 1 << 30)), 1 << log_n, m)
            f = ctx.ntt(a, log_n)
            ok = np.array_equal(f, orc.ntt(a, log_n, w, m)) and np.array_equal(ctx.intt(f, log_n), a)
            off = int(rng.integers(1, m))
            ok = ok and np.array_equal(ctx.coset_domain(log_n, off), orc.coset_domain(off, w, 1 << log_n, m))
            check("ntt_domain", ok, f"log_n={log_n} offset={off} modulus={m}")
        elif kind == 7:                                 # four-step transform, all ranks emulated on this GPU (both exchange styles)
            mg = importlib.import_module("stark-prover_b200.multi_gpu")
            world = int(rng.choice([1, 2, 4, 8]))
            log_n = int(rng.integers(max(10, 2 * (world.bit_length() - 1) + 4), min(two, 17) + 1))
            nco = int(rng.integers(1, (1 << log_n) + 1))
            off = int(rng.integers(1, m))
            c = orc.synthetic_column(int(rng.integers(1, 1 << 30)), nco, m)
            want = orc.coset_evaluate(c, log_n, off, orc.root_of_unity(log_n, m, g), m)
            got = np.concatenate([b.download() for b in mg.four_step_lde_emulated(sp, ctx, c, log_n, off, world)])
            ok, which = np.array_equal(got, want), "all-to-all style"
            if ok:
                try:
                    blocks = mg.four_step_p2p_emulated(sp, ctx, c, log_n, off, world)
                    ok, which = np.array_equal(np.concatenate([b.download() for b in blocks]), want), "peer-memory style"
                    for b in blocks: b.free()
                except sp.StarkError as e:              # documented size limit of the peer-memory kernels (>= 32 rows and columns per rank)
                    ok = "every rank needs" in str(e)
                    which = f"peer-memory style raised {e}"
            check("four_step", ok, f"{which}: log_n={log_n} coeffs={nco} world={world} offset={off} modulus={m}")
        elif kind == 8:                                 # column-batched transforms (the four-step building block)
            log_m = int(rng.integers(1, min(two, 12) + 1))
            batch = int(rng.integers(1, 100))
            a = orc.synthetic_column(int(rng.integers(1, 1 << 30)), batch << log_m, m)
            v = ctx.upload(a)
            inv = bool(rng.integers(0, 2))
            ctx.ntt_batch_dev(v, log_m, inv)
            w = orc.root_of_unity(log_m, m, g)
            want = np.concatenate([(orc.intt if inv else orc.ntt)(a[i << log_m:(i + 1) << log_m], log_m, w, m) for i in range(batch)])
            check("ntt_batch", np.array_equal(v.download(), want), f"log_m={log_m} batch={batch} inverse={inv} modulus={m}")
            v.free()
        elif kind == 9:                                 # occasionally large: 2^19..2^22 trees (root + a path) and transforms
            if rng.random() < 0.85:
                continue
            n = int(rng.integers(1 << 19, (1 << 22) + 1))
            vals = orc.synthetic_column(int(rng.integers(1, 1 << 30)), n, m)
            t = sp.MerkleTree.new(ctx, vals)
            ok = t.root_bytes() == orc.merkle_root_only(vals)
            idx = int(rng.integers(0, n))
            ok = ok and sp.merkle_validate(t.root_bytes(), n, idx, int(vals[idx]), t.get_authentication_path(idx))
            t.free()
            log_n = int(rng.integers(19, min(two, 22) + 1))
            nco = int(rng.integers(1, (1 << log_n) + 1))
            off = int(rng.integers(1, m))
            c = orc.synthetic_column(int(rng.integers(1, 1 << 30)), nco, m)
            ok = ok and np.array_equal(ctx.coset_evaluate(c, log_n, off), orc.coset_evaluate(c, log_n, off, orc.root_of_unity(log_n, m, g), m))
            check("large", ok, f"leaves={n} log_n={log_n} coeffs={nco} offset={off} modulus={m}")
        elif kind == 10:                                # FRI: one batched opening == per-index openings; the verifier accepts the transcript
            log_n = int(rng.integers(2, min(two, 14) + 1))
            nco = int(rng.integers(2, (1 << int(rng.integers(1, log_n + 1))) + 1))
            off = int(rng.integers(1, m))
            c = orc.synthetic_poly_exact_degree(int(rng.integers(1, 1 << 30)), nco, m)
            q = int(rng.integers(1, 5))
            ch = sp.Channel(m)
            pr = sp.fri_commit(ctx, c, sp.CosetFri(ctx, off, log_n), ch)
            idxs = [int(x) for x in rng.integers(0, 1 << log_n, 3)]
            ok = pr.open(idxs) == b"".join(pr.open([i]) for i in idxs)
            sp.decommit_fri(q, (1 << log_n) - 1, pr, ch)
            good, why = sp.verify_fri(ch.proof_flat(), log_n, off, q, (1 << log_n) - 1, (nco - 1).bit_length(), m, g)
            check("fri_open_verify", ok and good, f"log_n={log_n} coeffs={nco} offset={off} q={q} modulus={m} verifier={why!r}")
            pr.free()
        elif kind == 11:                                # the sharded prover's code path with one rank == the plain prover
            mg = importlib.import_module("stark-prover_b200.multi_gpu")
            ctx = ctx_for(*MODULI[0])
            log_t, log_b = int(rng.integers(7, 13)), 3
            a1, q = int(rng.integers(0, MODULI[0][0])), int(rng.integers(1, 3))
            ch, ch1 = sp.Channel(MODULI[0][0]), sp.Channel(MODULI[0][0])
            mg.stark101_prove_multi(sp, ctx, ch, a1, log_t, log_b, q, 0, 1)
            sp.stark101_prove(ctx, ch1, a1, log_t, log_b, q)
            check("stark101_multi_world1", ch.state == ch1.state and ch.proof == ch1.proof, f"log_trace={log_t} a1={a1} q={q}")
        elif kind == 12:                                # C-level group of one rank, the large FRI layers hashed in leaf ranges (forced) at a random threshold
            ctx = ctx_for(*MODULI[0])
            m0 = MODULI[0][0]
            log_n = int(rng.integers(10, 17))
            nco = int(rng.integers(2, (1 << (log_n - 1)) + 1))
            q = int(rng.integers(1, 4))
            os.environ["STARK_MG_FRI_SHARD_FORCE"] = "1"
            os.environ["STARK_MG_FRI_SHARD_MIN_LOG"] = str(int(rng.integers(6, log_n + 1)))
            try:
                c = orc.synthetic_poly_exact_degree(int(rng.integers(1, 1 << 30)), nco, m0)
                grp = sp.MultiGpu(ctx, 0, 1)
                ch, och = sp.Channel(m0), orc.Channel(m0)
                f = grp.fri_commit(ctx.upload(c), log_n, 5, ch, int(rng.integers(0, 2)))
                grp.decommit_fri(f, q, (1 << log_n) - 1, ch)
                opr = orc.fri_commit_fast(c, log_n, 5, orc.root_of_unity(log_n), och, m0)
                orc.decommit_fri(q, (1 << log_n) - 1, opr, och)
                check("mg_fri_leaf_range_layers", ch.state == och.state and ch.proof == och.proof,
                      f"log_n={log_n} coeffs={nco} q={q} min_log={os.environ['STARK_MG_FRI_SHARD_MIN_LOG']}")
                f.free(); grp.close()
            finally:
                os.environ.pop("STARK_MG_FRI_SHARD_FORCE", None); os.environ.pop("STARK_MG_FRI_SHARD_MIN_LOG", None)
        elif kind == 13:                                # FRI layers by value, streamed to the host during the commit
            log_n = int(rng.integers(2, min(two, 18) + 1))
            nco = int(rng.integers(1, (1 << log_n) + 1))
            off = int(rng.integers(1, m))
            c = orc.synthetic_column(int(rng.integers(1, 1 << 30)), nco, m)
            buf = np.full(2 << log_n, 0xABCDABCDABCDABCD, dtype=np.uint64)
            ch, ch1 = sp.Channel(m), sp.Channel(m)
            pr = sp.fri_commit(ctx, c, sp.CosetFri(ctx, off, log_n), ch, layers_out=buf)
            p1 = sp.fri_commit(ctx, c, sp.CosetFri(ctx, off, log_n), ch1)
            ok, o = ch.state == ch1.state and pr.num_layers == p1.num_layers, 0
            for k in range(p1.num_layers):
                ln = p1.layer_len(k)
                ok = ok and pr.layer_host_offset(k) == o and np.array_equal(buf[o:o + ln], p1.layer(k))
                o += ln
            ok = ok and bool(np.all(buf[o:] == np.uint64(0xABCDABCDABCDABCD)))
            check("fri_layers_by_value", ok, f"log_n={log_n} coeffs={nco} offset={off} modulus={m}")
            pr.free(); p1.free()
        else:                                           # kind 4: the build-defined prover + its verifier (default field only)
            ctx = ctx_for(*MODULI[0])
            log_t, log_b = int(rng.integers(2, 13)), int(rng.integers(1, 5))
            a1, q = int(rng.integers(0, MODULI[0][0])), int(rng.integers(1, 4))
            ch, och = sp.Channel(MODULI[0][0]), orc.Channel(MODULI[0][0])
            sp.stark101_prove(ctx, ch, a1, log_t, log_b, q)
            orc.stark101_prove(och, a1, log_t, log_b, sp.G_DEFAULT, q, literal=False)
            claimed = int(orc.fibsq_trace(a1, (1 << log_t) - 1)[(1 << log_t) - 2])
            okv, why = sp.stark101_verify(ch.proof_flat(), claimed, log_t, log_b, q)
            check("stark101", ch.state == och.state and ch.proof == och.proof and okv, f"log_trace={log_t} log_blowup={log_b} a1={a1} q={q} verifier={why}")


if __name__ == "__main__":
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 12345
    threads = int(sys.argv[3]) if len(sys.argv) > 3 else 1          # > 1: concurrent workers, separate contexts on one GPU
    if threads == 1:
        run(budget, seed)
    else:
        ws = [threading.Thread(target=run, args=(budget, seed + 1000 * k, f"[worker {k}] ")) for k in range(threads)]
        for w in ws: w.start()
        for w in ws: w.join()
    print(f"soak: {sum(counts.values())} cases in {budget:.0f} s on {threads} thread(s) {counts}, mismatches: {fails[0]} (seed {seed})")
    sys.exit(1 if fails[0] else 0)
