#!/usr/bin/env python
"""Two ranks, peer-memory four-step LDE at 2^26 a few times (for NVLink byte counters of the exchange kernels: rank 0 runs
under `ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum -k regex:fourstep_`, see tools/prof_fourstep_launch.sh)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")                      # plumbing only: no NCCL kernels next to the profiled ones
sp = importlib.import_module("stark-prover_b200")
mg = importlib.import_module("stark-prover_b200.multi_gpu")
synth = importlib.import_module("stark-prover_b200.synthetic")
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 26
ctx = sp.Context(sp.P_DEFAULT, 5, local)
cv = ctx.upload(synth.synthetic_poly_exact_degree(43, 1 << (log_n - 3)))
fs = mg.FourStepP2P(sp, ctx, log_n, rank, world, host_barriers=True)     # host barriers: a profiled rank may be arbitrarily slow
for _ in range(3):
    fs.run(cv, 5)
ctx.sync()
tree = sp.MerkleTree.new(ctx, fs.block)
root, _ = mg.commit_leaf_ranges(tree.root_bytes, rank, world)
if rank == 0:
    print("root", root.hex(), flush=True)
tree.free()
fs.close()
dist.barrier()
dist.destroy_process_group()
ctx.close()
