#!/usr/bin/env python
"""Both ranks of the peer-memory four-step LDE in ONE process on TWO GPUs (peer access enabled between the devices, no
CUDA IPC: cudaIpcOpenMemHandle does not work under ncu), for the NVLink byte counters of the two exchange kernels:
    ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum,gpu__time_duration.sum -k regex:fourstep_ python tools/prof_fourstep_inproc.py
The kernels and the index algebra are exactly those of the multi-process run; the phases are host-synchronised here."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
sp = importlib.import_module("stark-prover_b200")
mg = importlib.import_module("stark-prover_b200.multi_gpu")
synth = importlib.import_module("stark-prover_b200.synthetic")
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 26
world = 2
assert torch.cuda.device_count() >= 2 and torch.cuda.can_device_access_peer(0, 1)
# torch enables peer access in both directions on the first cross-device copy
a = torch.ones(16, device="cuda:0"); b = a.to("cuda:1"); c = b.to("cuda:0"); torch.cuda.synchronize(0); torch.cuda.synchronize(1)
ctxs = [sp.Context(sp.P_DEFAULT, 5, r) for r in range(world)]
coeffs = synth.synthetic_poly_exact_degree(43, 1 << (log_n - 3))
cvs = [cx.upload(coeffs) for cx in ctxs]
n_loc = (1 << log_n) // world
rows = [cx.peer_alloc(n_loc)[0] for cx in ctxs]
blocks = [cx.peer_alloc(n_loc)[0] for cx in ctxs]
pr, pb = [v.device_ptr for v in rows], [v.device_ptr for v in blocks]
for rep in range(3):
    for r, cx in enumerate(ctxs):
        cx.fourstep_phase_a(cvs[r], log_n, 5, world, r, pr)          # returns after its peer stores are complete
    for r, cx in enumerate(ctxs):
        cx.fourstep_phase_c(rows[r], log_n, world, r, pb)
subs = [sp.MerkleTree.new(cx, blocks[r]).root_bytes() for r, cx in enumerate(ctxs)]
print("root", mg.combine_subtree_roots(subs).hex(), flush=True)
