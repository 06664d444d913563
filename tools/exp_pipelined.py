"""Two (or more) independent FRI instances per GPU from separate host threads/contexts: the host-bound openings
of one overlap the device-bound commit of the other."""
import importlib, os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
sp = importlib.import_module("stark-prover_b200")
synth = importlib.import_module("stark-prover_b200.synthetic")
P, LOG_N, Q = sp.P_DEFAULT, 24, 32

def worker(k, reps, out):
    ctx = sp.Context()
    c = ctx.upload(synth.synthetic_poly_exact_degree(43 + k, 1 << (LOG_N - 3)))
    dom = sp.CosetFri(ctx, 5, LOG_N)
    def step():
        ch = sp.Channel(P)
        pr = sp.fri_commit(ctx, c, dom, ch)
        sp.decommit_fri(Q, (1 << LOG_N) - 1, pr, ch)
        pr.free()
        return ch.state
    for _ in range(3): step()
    barrier.wait()
    t0 = time.perf_counter()
    for _ in range(reps): st = step()
    out[k] = (time.perf_counter() - t0, st)
    ctx.close()

for nthreads in (1, 2, 3):
    barrier = threading.Barrier(nthreads)
    out = {}
    reps = 10
    ths = [threading.Thread(target=worker, args=(k, reps, out)) for k in range(nthreads)]
    t0 = time.perf_counter()
    for t in ths: t.start()
    for t in ths: t.join()
    wall = max(v[0] for v in out.values())
    print(f"{nthreads} concurrent instance(s): {wall / reps * 1e3:.2f} ms per round of {nthreads} -> {nthreads * reps * (1 << LOG_N) / wall / 1e6:.0f} Melem/s aggregate")
