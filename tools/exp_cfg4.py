import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
sp = importlib.import_module("stark-prover_b200")
from oracle import pyoracle as orc
ctx = sp.Context()
log_rows = 22
col = orc.synthetic_column(100, 1 << log_rows)
t = torch.empty(col.size, dtype=torch.int64).pin_memory(); pv = t.numpy().view(np.uint64); pv[:] = col
for name, src in (("pageable", col), ("pinned", pv), ("pageable", col), ("pinned", pv)):
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        v = ctx.upload(src); ctx.sync(); t1 = time.perf_counter()
        lde = ctx.coset_lde_dev(v, 1, 3, 5); ctx.sync(); t2 = time.perf_counter()
        tree = sp.MerkleTree.new(ctx, lde); r = tree.root_bytes(); t3 = time.perf_counter()
        tree.free(); lde.free(); v.free(); ctx.sync(); t4 = time.perf_counter()
        print(f"{name:9s} upload {1e3*(t1-t0):6.2f}  lde {1e3*(t2-t1):6.2f}  tree {1e3*(t3-t2):6.2f}  free {1e3*(t4-t3):6.2f} ms")
