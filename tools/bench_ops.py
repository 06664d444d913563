#!/usr/bin/env python
"""Per-operation device timing at BASELINE sizes (inputs resident in HBM, CUDA events on the library's stream):
NTT / iNTT / LDE / batch inverse / quotient / Merkle commit / one fused fold layer.
Algorithmic bytes per SURVEY.md 8(d) (8 B per element at the API, 32 B per digest) against MEASURED_PEAKS.json."""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
sp = importlib.import_module("stark-prover_b200")
from oracle import pyoracle as orc

P = sp.P_DEFAULT
ctx = sp.Context()
stream = torch.cuda.ExternalStream(ctx.stream)
try:
    hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    hbm = 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps=10):
    fn()
    tot = 0.0
    for _ in range(reps):
        with torch.cuda.stream(stream):
            flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        out = fn()
        e1.record(stream)
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
        if hasattr(out, "free"):
            out.free()
    return tot / reps


rows = []
def report(name, ms, alg_bytes, n):
    gbs = alg_bytes / (ms * 1e-3) / 1e9
    rows.append({"op": name, "n": n, "ms": ms, "algorithmic_GBps": gbs, "frac_of_hbm": gbs / hbm, "Melem_per_s": n / (ms * 1e-3) / 1e6})
    print(f"{name:34s} n=2^{int(np.log2(n)):2d}  {ms:8.3f} ms  {gbs:8.1f} GB/s algorithmic ({gbs / hbm:5.1%} of {hbm:.0f})  {n / ms / 1e3:9.1f} Melem/s")


for log_n in (20, 24):
    n = 1 << log_n
    v = ctx.upload(orc.synthetic_column(1, n))
    report("coset_evaluate (NTT, nat->nat)", timed(lambda: ctx.coset_evaluate_dev(v, log_n, 5)), 16 * n, n)
    report("coset_interpolate (iNTT)", timed(lambda: ctx.coset_interpolate_dev(v, 5)), 16 * n, n)
    report("batch_inverse", timed(lambda: ctx.batch_inverse_dev(v)), 16 * n, n)
    report("quotient_pointwise", timed(lambda: ctx.quotient_pointwise_dev(v, v)), 24 * n, n)
    report("merkle_commit (tree retained)", timed(lambda: sp.MerkleTree.new(ctx, v)), 8 * n + 32 * (2 * n - 1), n)
    v.free()
for log_t in (17, 21, 22):
    t = ctx.upload(orc.synthetic_column(2, 1 << log_t))
    N = 1 << (log_t + 3)
    report(f"coset_lde 2^{log_t} -> 2^{log_t + 3}", timed(lambda: ctx.coset_lde_dev(t, 1, 3, 5)), 8 * (1 << log_t) + 8 * N, N)
    t.free()
# one fused fold-and-hash layer: 2^24 -> 2^23
c = orc.synthetic_poly_exact_degree(43, 1 << 21)
pr, _ = sp.fri_begin(ctx, ctx.upload(c), 24, 5)
ms_total = 0.0
reps = 5
for it in range(reps + 1):
    p2, _ = sp.fri_begin(ctx, ctx.upload(c), 24, 5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); p2.fold(12345); e1.record(stream); torch.cuda.synchronize()
    if it > 0:                         # the first repetition grows the allocation pool
        ms_total += e0.elapsed_time(e1)
    p2.free()
n = 1 << 24
report("fri fold+commit layer 2^24->2^23", ms_total / reps, 8 * n + 8 * (n // 2) + 32 * (n - 1), n)
print(json.dumps({"hbm_gbs_peak": hbm, "ops": rows}))
