"""Generates tests/golden/transcripts.json: transcripts of the oracle's LITERAL tier (Horner / Lagrange /
long division — the reference's own algorithms) at sizes where it finishes in seconds, cross-checked
against the hashlib twin.  Run from the repo root:  python tests/golden/make_golden.py
The reference (Rust nightly, un-vendored crates) cannot run in this image, so these pin the restatement,
not a run of the reference (see oracle/stark_oracle.h, PINNING)."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as o  # noqa: E402

P = o.P_DEFAULT


def sha(b: bytes) -> str:
    return hashlib.sha256(b).hexdigest()


def fri_case(seed, log_n, log_deg, offset, queries):
    n = 1 << log_n
    w = o.root_of_unity(log_n, P)
    c = o.synthetic_poly_exact_degree(seed, 1 << log_deg, P)
    d = o.coset_domain(offset, w, n, P)
    ch = o.Channel(P)
    pr = o.fri_commit_literal(c, d, ch, P)
    o.decommit_fri(queries, n - 1, pr, ch)
    # independent re-derivation of roots and transcript with hashlib
    pch = o.PyChannel(P)
    for k in range(pr.num_layers):
        root = o.py_merkle_root_hex(pr.layer(k).tolist())
        assert root == pr.tree(k).root_hex()
        pch.send(root.encode())
        if k + 1 < pr.num_layers:
            pch.receive_random_field_element()
    fin = pr.final_poly()
    pch.send(int(fin[0] if len(fin) else 0).to_bytes(8, "big"))
    return {
        "seed": seed, "log_n": log_n, "log_deg": log_deg, "offset": offset, "queries": queries,
        "num_layers": pr.num_layers,
        "roots": [pr.tree(k).root_hex() for k in range(pr.num_layers)],
        "layer_sha256": [sha(pr.layer(k).astype("<u8").tobytes()) for k in range(pr.num_layers)],
        "final_poly": [int(x) for x in fin],
        "state_after_commit": pch.state,
        "final_state": ch.state,
        "proof_size": ch.proof_size(),
        "proof_sha256": sha(ch.proof_flat()),
    }


def main():
    out = {"_comment": __doc__, "modulus": P, "fri": [], "stark101": {}}
    out["fri"].append(fri_case(43, 10, 7, 5, 3))
    out["fri"].append(fri_case(44, 12, 9, 5, 4))
    out["fri"].append(fri_case(45, 6, 6, 3, 2))      # blowup 1: folds down to a one-point layer (length == 1 branch)
    ch = o.Channel(P)
    o.stark101_prove(ch, literal=True)
    out["stark101"] = {"a1": 3141592, "log_trace": 10, "log_blowup": 3, "queries": 3,
                       "final_state": ch.state, "proof_size": ch.proof_size(),
                       "compressed_proof_size": ch.compressed_proof_size(),
                       "proof_sha256": sha(ch.proof_flat()), "n_messages": len(ch.proof),
                       "statement": ch.proof[0].hex(), "first_root": ch.proof[1].decode()}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "transcripts.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
