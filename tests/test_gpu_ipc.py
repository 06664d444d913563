"""Two PROCESSES on cuda:0 drive the peer-memory four-step transform through real CUDA IPC handles
(cudaIpcGetMemHandle / cudaIpcOpenMemHandle work between processes on one device; NCCL does not): the peer-store kernels
and the device-side epoch hand-over of csrc/fourstep.cu run against memory that belongs to another process, as they do on
several GPUs.  The plumbing between the processes is torch.distributed with the gloo backend."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import importlib, os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["STARK_ROOT"])
sp = importlib.import_module("stark-prover_b200")
mg = importlib.import_module("stark-prover_b200.multi_gpu")
from oracle import pyoracle as orc
P = sp.P_DEFAULT
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
ctx = sp.Context(P, 5, 0)                       # both ranks on device 0
for log_n, log_deg, host_barriers in ((12, 9, False), (18, 15, False), (18, 18, True), (20, 17, False)):
    coeffs = orc.synthetic_poly_exact_degree(40 + log_n, 1 << log_deg)
    want = orc.coset_evaluate(coeffs, log_n, 5, orc.root_of_unity(log_n), P)
    fs = mg.FourStepP2P(sp, ctx, log_n, rank, world, host_barriers=host_barriers)
    cv = ctx.upload(coeffs)
    blk = (1 << log_n) // world
    for rep in range(3):                        # epochs 1..3 over the same buffers
        got = fs.run(cv, 5).download()
        assert np.array_equal(got, want[rank * blk:(rank + 1) * blk]), (log_n, rep, rank)
    tree = sp.MerkleTree.new(ctx, fs.block)
    root, subs = mg.commit_leaf_ranges(tree.root_bytes, rank, world)
    assert root == orc.merkle_root_only(want), (log_n, rank)
    tree.free()
    fs.close()
dist.barrier()
dist.destroy_process_group()
ctx.close()
print(f"rank {rank} ok")
'''


def test_four_step_over_cuda_ipc_two_processes(tmp_path):
    script = tmp_path / "ipc_worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, STARK_ROOT=ROOT, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=300)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"rank {r} ok" in out, out[-3000:]
