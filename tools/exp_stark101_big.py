import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sp = importlib.import_module("stark-prover_b200")
from oracle import pyoracle as orc
P = sp.P_DEFAULT
ctx = sp.Context()
for log_trace, check in ((10, True), (18, True), (20, True), (23, False)):
    for rep in range(2):
        ch = sp.Channel(P)
        t0 = time.perf_counter()
        sp.stark101_prove(ctx, ch, 3141592, log_trace, 3, 3)
        dt = time.perf_counter() - t0
    line = f"log_trace={log_trace} domain=2^{log_trace + 3}: prove {dt * 1e3:.1f} ms, proof {ch.proof_size()} bytes, state {ch.state[:16]}"
    if check:
        och = orc.Channel(P)
        t0 = time.perf_counter()
        orc.stark101_prove(och, a1=3141592, log_trace=log_trace, log_blowup=3, num_queries=3, literal=False)
        line += f" | oracle {1e3 * (time.perf_counter() - t0):.0f} ms, identical={och.state == ch.state and och.proof == ch.proof}"
    print(line, flush=True)
