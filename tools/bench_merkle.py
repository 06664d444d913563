#!/usr/bin/env python
"""Micro-benchmark of the Merkle commitment alone (kernel experiments): python tools/bench_merkle.py [log_n]"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
sp = importlib.import_module("stark-prover_b200")
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
ctx = sp.Context()
vals = (np.arange(1 << log_n, dtype=np.uint64) * np.uint64(2654435761)) % np.uint64(sp.P_DEFAULT)
v = ctx.upload(vals)
for _ in range(3):
    sp.MerkleTree.new(ctx, v).free()
ctx.set_timing(True); ctx.read_timing()
reps = 10
t0 = time.perf_counter()
for _ in range(reps):
    t = sp.MerkleTree.new(ctx, v); root = t.root(); t.free()
wall = (time.perf_counter() - t0) / reps * 1e3
kt = ctx.read_timing()
leaf, node = kt["merkle_leaf"]["ms"] / reps, kt["merkle_node"]["ms"] / reps
ops = (kt["merkle_leaf"]["units"] + kt["merkle_node"]["units"]) / reps
print(f"log_n={log_n} defs='{os.environ.get('STARK_NVCC_DEFS','')}' wall={wall:.3f} ms leaf={leaf:.3f} node={node:.3f} "
      f"algorithmic={ops / ((leaf + node) * 1e-3) / 1e12:.2f} Tint-op/s root={root[:16]}")
