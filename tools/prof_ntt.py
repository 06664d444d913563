#!/usr/bin/env python
"""Smallest program that runs the transform kernels once at 2^24 (for ncu -k regex:'nat_|lde8_'):
natural-order evaluate + interpolate, and the blow-up-by-8 LDE of a 2^21 column."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sp = importlib.import_module("stark-prover_b200")
synth = importlib.import_module("stark-prover_b200.synthetic")

ctx = sp.Context()
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
v = ctx.upload(synth.synthetic_column(1, 1 << log_n))
for _ in range(2):
    e = ctx.coset_evaluate_dev(v, log_n, 5)
    c = ctx.coset_interpolate_dev(e, 5)
    e.free(); c.free()
t = ctx.upload(synth.synthetic_column(2, 1 << (log_n - 3)))
for _ in range(2):
    ctx.coset_lde_dev(t, 1, 3, 5).free()
ctx.sync()
print("ok")
