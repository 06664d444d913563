#!/usr/bin/env python
"""cfg5 fri_commit + openings on N GPUs against the size below which FRI layers are left to rank 0
(STARK_MG_FRI_SHARD_MIN_LOG; csrc/multi.cu: sharded_fri_layers).  Run under torchrun:
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/exp_shard_threshold.py
Prints one line per threshold: best-of-3 max-over-ranks ms, transcript checked against the committed golden."""
import importlib
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sp = importlib.import_module("stark-prover_b200")
    synth = importlib.import_module("stark-prover_b200.synthetic")
    P = 3221225473
    g5 = json.load(open(os.path.join(ROOT, "tests", "golden", "sharded.json")))["cfg5"]
    ctx = sp.Context(P, sp.G_DEFAULT, local)
    mg = sp.MultiGpu.from_torch(ctx)
    log_n, q = g5["log_n"], g5["queries"]
    cvec = ctx.upload(synth.synthetic_poly_exact_degree(g5["seed"], 1 << (log_n - 3), P))

    def fri():
        ch = sp.Channel(P) if rank == 0 else None
        f = mg.fri_commit(cvec, log_n, g5["offset"], ch, 1)
        mg.decommit_fri(f, q, (1 << log_n) - 1, ch)
        return f, ch

    for min_log in (25, 23, 22, 21, 20, 19, 18, 17):
        os.environ["STARK_MG_FRI_SHARD_MIN_LOG"] = str(min_log)
        best = None
        for rep in range(4):
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            f, ch = fri()
            torch.cuda.synchronize()
            ms = torch.tensor([(time.perf_counter() - t0) * 1e3], device=f"cuda:{local}")
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            if rank == 0:
                assert ch.state == g5["final_state"], f"transcript differs from the golden at min_log {min_log}"
            f.free()
            if rep:
                best = float(ms) if best is None else min(best, float(ms))
        if rank == 0:
            print(f"shard layers of >= 2^{min_log} leaves: {best:.3f} ms (wall, max over {world} ranks, best of 3)", flush=True)
    mg.close()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
