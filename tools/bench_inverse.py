#!/usr/bin/env python
"""Micro-benchmark of the batched inverse alone (kernel experiments): python tools/bench_inverse.py [log_n]"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
sp = importlib.import_module("stark-prover_b200")
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
ctx = sp.Context()
n = 1 << log_n
vals = (np.arange(1, n + 1, dtype=np.uint64) * np.uint64(2654435761)) % np.uint64(sp.P_DEFAULT)
v = ctx.upload(vals)
stream = torch.cuda.ExternalStream(ctx.stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(200):                    # clocks ramp up over the first few ms of work
    ctx.batch_inverse_dev(v).free()
tot, reps = 0.0, 30
for _ in range(reps):
    with torch.cuda.stream(stream):
        flush.zero_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); o = ctx.batch_inverse_dev(v); e1.record(stream)
    torch.cuda.synchronize()
    tot += e0.elapsed_time(e1)
    chk = int(o.download()[12345]); o.free()
ms = tot / reps
assert chk * int(vals[12345]) % sp.P_DEFAULT == 1
print(f"log_n={log_n} defs='{os.environ.get('STARK_NVCC_DEFS', '')}' batch_inverse {ms:.4f} ms  {n / ms / 1e6:.1f} Gelem/s  "
      f"{16 * n / ms / 1e6:.0f} GB/s algorithmic")
