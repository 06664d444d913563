#!/usr/bin/env python
"""Smallest program that runs batch_inverse_kernel at 2^24 elements, plain and with a numerator (the point-wise quotient),
for ncu -k regex:batch_inverse."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sp = importlib.import_module("stark-prover_b200")
synth = importlib.import_module("stark-prover_b200.synthetic")
ctx = sp.Context()
n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 24)
a, b = ctx.upload(synth.synthetic_column(1, n)), ctx.upload(synth.synthetic_column(2, n))
for _ in range(2):
    ctx.batch_inverse_dev(a).free()
    ctx.quotient_pointwise_dev(a, b).free()
ctx.sync()
print("ok")
