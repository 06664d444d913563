#!/usr/bin/env python
"""Where the by-value FRI layers cost time (stark_fri_commit_to_host): the 2^24 headline step with the layers streamed to
pinned host memory, against the same step without them, and with the copy / the widening kernel switched off
(STARK_SINK_DEBUG, read once per process: this script re-runs itself).  One B200."""
import importlib
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def measure():
    import numpy as np
    import torch
    sp = importlib.import_module("stark-prover_b200")
    synth = importlib.import_module("stark-prover_b200.synthetic")
    P, log_n = 3221225473, 24
    ctx = sp.Context(P, sp.G_DEFAULT, 0)
    pin = torch.empty(1 << (log_n - 3), dtype=torch.int64).pin_memory()
    c = pin.numpy().view(np.uint64)
    c[:] = synth.synthetic_poly_exact_degree(43, 1 << (log_n - 3), P)
    keep = torch.empty(2 << log_n, dtype=torch.int64).pin_memory()
    buf = keep.numpy().view(np.uint64)
    dom = sp.CosetFri(ctx, 5, log_n)

    def step(layers, wait):
        ch = sp.Channel(P)
        t0 = time.perf_counter()
        pr = sp.fri_commit(ctx, c, dom, ch, layers_out=buf if layers else None, wait=wait)
        t1 = time.perf_counter()
        sp.decommit_fri(32, (1 << log_n) - 1, pr, ch)
        t2 = time.perf_counter()
        if layers:
            pr.layers_wait()
        t3 = time.perf_counter()
        pr.free()
        return (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3

    for name, layers, wait in (("resident", False, True), ("by value, complete on return", True, True), ("by value, wait after the openings", True, False)):
        for _ in range(3):
            step(layers, wait)
        best = None
        for _ in range(10):
            r = step(layers, wait)
            if best is None or sum(r) < sum(best):
                best = r
        print(f"  {name:36s} fri_commit {best[0]:.3f} ms  decommit_fri {best[1]:.3f} ms  layers_wait {best[2]:.3f} ms  total {sum(best):.3f} ms", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        measure()
    else:
        configs = [({}, "the product: every folded layer recomputed on the copy stream at the START of its tree, widened, cudaMemcpyAsync"),
                   ({"STARK_SINK_EARLY": "0"}, "a layer sent once the fused fold-and-hash launch has produced it (under the following layers)"),
                   ({"STARK_SINK_DEBUG": "1"}, "widening kernel only, no copy"),
                   ({"STARK_SINK_DEBUG": "2"}, "no kernel, no copy: API bookkeeping only")]
        # (a variant whose kernel wrote the u64 values straight into mapped pinned memory, 8 .. 296 CTAs, measured 13.5 .. 26.5 ms
        # per step against 9.3 ms: profiles/r02_by_value_modes.txt; removed)
        for env, what in configs:
            print((" ".join(f"{k}={v}" for k, v in env.items()) or "(default)") + ": " + what, flush=True)
            subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=dict(os.environ, **env), check=True)
