"""Generates tests/golden/sharded.json: the CPU oracle's answers for the two sharded BASELINE configs at FULL size, so
that `bench.py --gpus N` can assert its multi-GPU results inside the run (the oracle itself takes minutes here):
  cfg4  64 columns x 2^22 rows (seeds 100+c), trace domain <g>, LDE on the coset 5*<h> of size 2^25: the 64 Merkle roots
  cfg5  one polynomial of degree 2^23-1 (seed 43) on the coset 5*<w> of size 2^26: root of layer 0, every later root,
        the final constant and the transcript after fri_commit + 8 openings
  prove5  the FibonacciSq prover (stark101_prove) with a 2^23-1 row trace, LDE/FRI domain 2^26, 3 queries: the transcript
Run from the repo root:  python tests/golden/make_golden_sharded.py   (~10 minutes on 8 cores, ~12 GB of RAM)"""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as o  # noqa: E402

P = o.P_DEFAULT


def cfg4(log_rows=22, n_cols=64, log_blowup=3):
    roots = []
    wr, wl = o.root_of_unity(log_rows, P), o.root_of_unity(log_rows + log_blowup, P)
    for c in range(n_cols):
        col = o.synthetic_column(100 + c, 1 << log_rows, P)
        coef = o.coset_interpolate(col, log_rows, 1, wr, P)
        lde = o.coset_evaluate(coef, log_rows + log_blowup, 5, wl, P)
        roots.append(o.merkle_root_only(lde).hex())
        print("cfg4 column", c, roots[-1][:16], flush=True)
    return {"log_rows": log_rows, "n_cols": n_cols, "log_blowup": log_blowup, "seed_base": 100, "offset_in": 1, "offset_out": 5,
            "roots": roots, "roots_sha256": hashlib.sha256("".join(roots).encode()).hexdigest()}


def cfg5(log_n=26, queries=8):
    c = o.synthetic_poly_exact_degree(43, 1 << (log_n - 3), P)
    ch = o.Channel(P)
    pr = o.fri_commit_fast(c, log_n, 5, o.root_of_unity(log_n, P), ch, P)
    state_commit = ch.state
    o.decommit_fri(queries, (1 << log_n) - 1, pr, ch)
    return {"log_n": log_n, "seed": 43, "offset": 5, "queries": queries, "num_layers": pr.num_layers,
            "roots": [pr.tree(k).root_hex() for k in range(pr.num_layers)],
            "final_poly": [int(x) for x in pr.final_poly()], "state_after_commit": state_commit, "final_state": ch.state,
            "proof_size": ch.proof_size(), "proof_sha256": hashlib.sha256(ch.proof_flat()).hexdigest()}


def prove(log_trace=23, log_blowup=3, queries=3, a1=3141592):
    ch = o.Channel(P)
    o.stark101_prove(ch, a1, log_trace, log_blowup, o.G_DEFAULT, queries, literal=False)
    return {"a1": a1, "log_trace": log_trace, "log_blowup": log_blowup, "queries": queries, "statement": ch.proof[0].hex(),
            "trace_root": ch.proof[1].decode(), "final_state": ch.state, "proof_size": ch.proof_size(), "n_messages": len(ch.proof),
            "proof_sha256": hashlib.sha256(ch.proof_flat()).hexdigest()}


def main():
    o.build()
    o.set_num_threads(len(os.sched_getaffinity(0)))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sharded.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    out["_comment"] = __doc__
    which = sys.argv[1:] or ["cfg5", "cfg5_small", "prove5", "prove5_small", "cfg4", "cfg4_small"]
    for w in which:
        t0 = time.time()
        if w == "cfg4":
            out[w] = cfg4()
        elif w == "cfg4_small":
            out[w] = cfg4(16, 16)
        elif w == "cfg5":
            out[w] = cfg5()
        elif w == "cfg5_small":
            out[w] = cfg5(22)
        elif w == "prove5":
            out[w] = prove()
        elif w == "prove5_small":
            out[w] = prove(15)
        print(w, "done in", round(time.time() - t0, 1), "s", flush=True)
        json.dump(out, open(path, "w"), indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
