#!/usr/bin/env python
"""Two ranks (two processes, one GPU each), peer-memory four-step LDE at 2^26 a few times, for the NVLink byte counters of
the exchange kernels: rank 0 runs under `ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum -k regex:fourstep_`
(tools/prof_fourstep_launch.sh).  The plumbing between the ranks is a directory of files: a profiled rank may start and run
arbitrarily late, which rendezvous libraries do not forgive.
    python tools/prof_fourstep.py <rank> <world> <dir> [log_n]"""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
rank, world, box = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
log_n = int(sys.argv[4]) if len(sys.argv) > 4 else 26
os.makedirs(box, exist_ok=True)
counter = [0]


def _wait(path, timeout=600):
    t0 = time.time()
    while not os.path.exists(path):
        if time.time() - t0 > timeout:
            raise TimeoutError(path)
        time.sleep(0.002)


def gather(a):
    counter[0] += 1
    k = counter[0]
    tmp = os.path.join(box, f"g{k}_{rank}.tmp")
    np.asarray(a, dtype=np.uint8).tofile(tmp)
    os.rename(tmp, os.path.join(box, f"g{k}_{rank}.bin"))
    out = []
    for r in range(world):
        _wait(os.path.join(box, f"g{k}_{r}.bin"))
        out.append(np.fromfile(os.path.join(box, f"g{k}_{r}.bin"), dtype=np.uint8))
    return np.stack(out)


def barrier():
    gather(np.zeros(1, dtype=np.uint8))


sp = importlib.import_module("stark-prover_b200")
mg = importlib.import_module("stark-prover_b200.multi_gpu")
synth = importlib.import_module("stark-prover_b200.synthetic")
ctx = sp.Context(sp.P_DEFAULT, 5, rank)
cv = ctx.upload(synth.synthetic_poly_exact_degree(43, 1 << (log_n - 3)))
fs = mg.FourStepP2P(sp, ctx, log_n, rank, world, host_barriers=True, gather=gather, barrier=barrier)   # host barriers: the profiled rank is slow
for _ in range(3):
    fs.run(cv, 5)
ctx.sync()
tree = sp.MerkleTree.new(ctx, fs.block)
roots = gather(np.frombuffer(tree.root_bytes(), dtype=np.uint8))
root = mg.combine_subtree_roots([roots[r].tobytes() for r in range(world)])
if rank == 0:
    print("root", root.hex(), flush=True)
tree.free()
fs.close()
barrier()
ctx.close()
