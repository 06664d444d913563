#!/usr/bin/env python
"""Pipe-mix micro-benchmark: what does FMA-pipe work cost next to a saturated ALU pipe?  python tools/exp_pipe_mix.py"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sp = importlib.import_module("stark-prover_b200")
ctx = sp.Context()
names = ["3 ALU (SHF, LOP3, IADD3)", "+ IMAD", "+ mul.wide + mad (rotate on the FMA pipe)", "+ mul.hi (IMAD.HI)", "+ 2 IMAD",
         "no ALU work: 3 IMAD", "no ALU work: 3 IMAD.HI", "no ALU work: 3 IMAD.WIDE"]
t = ctx.measure_pipe_mix()
for n, v in zip(names, t):
    print(f"{n:48s} {v:6.2f} T steps/s   3 instr/step -> {3 * v:6.2f} T instr/s   relative to ALU alone {v / t[0]:.3f}")
