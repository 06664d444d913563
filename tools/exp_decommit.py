import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
sp = importlib.import_module("stark-prover_b200")
from oracle import pyoracle as orc
P = sp.P_DEFAULT
ctx = sp.Context()
c = orc.synthetic_poly_exact_degree(43, 1 << 21)
ch = sp.Channel(P)
pr = sp.fri_commit(ctx, ctx.upload(c), sp.CosetFri(ctx, 5, 24), ch)
for _ in range(3): pr.open([123])
t0 = time.perf_counter()
recs = [pr.open([1000 * i + 7]) for i in range(32)]
t1 = time.perf_counter()
print(f"32 x open (GPU round trip + python): {(t1 - t0) * 1e3:.3f} ms  ({(t1 - t0) / 32 * 1e6:.1f} us each), {len(recs[0])} bytes each")
ch2 = sp.Channel(P); ch2.send(b"x")
t0 = time.perf_counter()
sp.decommit_fri(32, (1 << 24) - 1, pr, ch2)
t1 = time.perf_counter()
print(f"decommit_fri(32): {(t1 - t0) * 1e3:.3f} ms")
# host hashing alone: feed the same bytes to a channel message by message
ch3 = sp.Channel(P); ch3.send(b"x")
msgs = ch2.proof[1:]
t0 = time.perf_counter()
L = sp.lib()
for m in msgs:
    L.stark_channel_send(ch3.h, m, len(m))
t1 = time.perf_counter()
print(f"channel.send of the {len(msgs)} messages alone (python loop): {(t1 - t0) * 1e3:.3f} ms")
