#!/usr/bin/env python
"""8 emulated ranks of the peer-memory four-step LDE at 2^26 on ONE GPU (for an ncu launch list: every rank's kernels have
the per-rank sizes of the real 8-GPU run, minus the NVLink hops)."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sp = importlib.import_module("stark-prover_b200")
mg = importlib.import_module("stark-prover_b200.multi_gpu")
synth = importlib.import_module("stark-prover_b200.synthetic")
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 26
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
flags = (int(sys.argv[3]) if len(sys.argv) > 3 else 1) != 0
ctx = sp.Context()
cv = ctx.upload(synth.synthetic_poly_exact_degree(43, 1 << (log_n - 3)))
for rep in range(2):
    ctx.sync(); t0 = time.perf_counter()
    blocks = mg.four_step_p2p_emulated(sp, ctx, cv, log_n, 5, world, flags=flags)
    ctx.sync(); dt = time.perf_counter() - t0
    for b in blocks:
        b.free()
    reps = 2 if flags else 1
    print(f"emulated {world} ranks, 2^{log_n}, flags={flags}: {dt * 1e3:.3f} ms for {reps} transforms of all ranks = {dt * 1e3 / reps / world:.3f} ms per rank per transform")
