#!/usr/bin/env python
"""Real multi-GPU check + timing (run under torchrun, one rank per GPU):
  torchrun --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py [--full]
cfg4: 64 columns x 2^22 rows (LDE 2^25) sharded by column, roots all-gathered (small sizes unless --full)
cfg5: four-step LDE of one 2^26-point column (2^22 unless --full) + per-rank subtree + root gather.
Every result is compared with the single-GPU path (and the CPU oracle at the small sizes)."""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
P = 3221225473


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--cols", type=int, default=None)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sp = importlib.import_module("stark-prover_b200")
    mg = importlib.import_module("stark-prover_b200.multi_gpu")
    from oracle import pyoracle as orc
    ctx = sp.Context(P, 5, local)
    out = {"world": world}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- cfg4: column-parallel LDE + commit
    log_rows, n_cols = (22, 64) if args.full else (16, 16)
    if args.cols:
        n_cols = args.cols
    def pinned(a):                       # host columns live in pinned memory, like a prover's trace buffers would
        t = torch.empty(a.size, dtype=torch.int64).pin_memory()
        v = t.numpy().view(np.uint64)
        v[:] = a
        return t, v
    cols_keep = {c: pinned(orc.synthetic_column(100 + c, 1 << log_rows)) for c in mg.shard_columns(n_cols, rank, world)}
    cols = {c: kv[1] for c, kv in cols_keep.items()}
    commit = mg.gpu_column_committer(sp, ctx, lambda c: cols[c], 3, 1, 5)
    mg.commit_columns(min(n_cols, world), commit if rank < n_cols else (lambda c: b"\0" * 32), rank, world)   # warm-up
    barrier(); t0 = time.perf_counter()
    roots = mg.commit_columns(n_cols, commit, rank, world)
    barrier(); dt = time.perf_counter() - t0
    out["cfg4"] = {"columns": n_cols, "log_rows": log_rows, "log_lde": log_rows + 3, "seconds": dt,
                   "Melem_per_s": n_cols * (1 << (log_rows + 3)) / dt / 1e6, "root0": roots[0].hex()}
    # the same with three contexts (streams) per GPU: columns pipelined over three host threads
    ctxs = [ctx, sp.Context(P, 5, local), sp.Context(P, 5, local)]
    commits = [mg.gpu_column_committer(sp, cx, lambda c: cols[c], 3, 1, 5) for cx in ctxs]
    mg.commit_columns(n_cols, commits, rank, world)            # warm-up
    barrier(); t0 = time.perf_counter()
    roots3 = mg.commit_columns(n_cols, commits, rank, world)
    barrier(); dt3 = time.perf_counter() - t0
    assert roots3 == roots
    out["cfg4"].update({"seconds_3_streams": dt3, "Melem_per_s_3_streams": n_cols * (1 << (log_rows + 3)) / dt3 / 1e6})
    for cx in ctxs[1:]:
        cx.close()
    if rank == 0 and not args.full:
        for c in (0, n_cols - 1):
            col = orc.synthetic_column(100 + c, 1 << log_rows)
            coef = orc.coset_interpolate(col, log_rows, 1, orc.root_of_unity(log_rows), P)
            lde = orc.coset_evaluate(coef, log_rows + 3, 5, orc.root_of_unity(log_rows + 3), P)
            assert roots[c] == orc.merkle_root_only(lde), f"cfg4 column {c} root differs from the oracle"
    # every rank must hold identical roots
    if world > 1:
        h = torch.tensor(list(roots[-1]), dtype=torch.uint8, device="cuda")
        hs = [torch.empty_like(h) for _ in range(world)]
        dist.all_gather(hs, h)
        assert all(torch.equal(hs[0], x) for x in hs)

    # ---------------- cfg5: four-step LDE + leaf-range commit
    log_n = 26 if args.full else 22
    coeffs = orc.synthetic_poly_exact_degree(43, 1 << (log_n - 3), P)
    cvec = ctx.upload(coeffs)                                                  # every rank holds the coefficients (N/8 elements)
    blk = mg.four_step_lde(sp, ctx, cvec, log_n, 5, rank, world)              # warm-up
    blk.free()
    barrier(); t0 = time.perf_counter()
    blk = mg.four_step_lde(sp, ctx, cvec, log_n, 5, rank, world)
    ctx.sync(); barrier(); t_lde = time.perf_counter() - t0
    t0 = time.perf_counter()
    tree = sp.MerkleTree.new(ctx, blk)
    root, subs = mg.commit_leaf_ranges(tree.root_bytes, rank, world)
    barrier(); t_commit = time.perf_counter() - t0
    out["cfg5"] = {"log_domain": log_n, "lde_seconds": t_lde, "commit_seconds": t_commit, "root": root.hex(),
                   "Melem_per_s": (1 << log_n) / (t_lde + t_commit) / 1e6}
    if rank == 0:
        # single-GPU answer for the same column
        ref = sp.MerkleTree.new(ctx, ctx.coset_evaluate_dev(cvec, log_n, 5))
        assert ref.root_bytes() == root, "cfg5 root differs from the single-GPU path"
        if log_n <= 22:
            want = orc.coset_evaluate(coeffs, log_n, 5, orc.root_of_unity(log_n), P)
            n_blk = (1 << log_n) // world
            assert np.array_equal(blk.download(), want[:n_blk]), "cfg5 block differs from the oracle"
        print(json.dumps(out), flush=True)
    # ---------------- cfg5, exchanges over peer memory (CUDA IPC + NVLink P2P stores from the kernels)
    fs = mg.FourStepP2P(sp, ctx, log_n, rank, world)
    fs.run(cvec, 5)                                                            # warm-up
    barrier(); t0 = time.perf_counter()
    pblk = fs.run(cvec, 5)
    barrier(); t_p2p = time.perf_counter() - t0
    assert np.array_equal(pblk.download(0, 4096), blk.download(0, 4096)) and np.array_equal(pblk.download(len(pblk) - 4096, 4096), blk.download(len(blk) - 4096, 4096))
    ptree = sp.MerkleTree.new(ctx, pblk)
    proot, _ = mg.commit_leaf_ranges(ptree.root_bytes, rank, world)
    assert proot == root, "peer-memory four-step differs from the NCCL path"
    if rank == 0:
        print(json.dumps({"world": world, "cfg5_p2p": {"log_domain": log_n, "lde_seconds_p2p": t_p2p, "lde_seconds_nccl": t_lde,
                                                       "Melem_per_s_p2p": (1 << log_n) / t_p2p / 1e6}}), flush=True)
    ptree.free()
    fs.close()

    # ---------------- cfg5 continued: the whole FRI commit + openings with the sharded layer 0
    log_f = 24 if args.full else 18
    cf = orc.synthetic_poly_exact_degree(43, 1 << (log_f - 3), P)
    ch = sp.Channel(P)
    barrier(); t0 = time.perf_counter()
    mp = mg.fri_commit_multi(sp, ctx, cf, log_f, 5, ch, rank, world)
    mg.decommit_fri_multi(sp, mp, 8, (1 << log_f) - 1, ch, rank, world)
    barrier(); t_fri = time.perf_counter() - t0
    if rank == 0:
        ch1 = sp.Channel(P)
        pr1 = sp.fri_commit(ctx, cf, sp.CosetFri(ctx, 5, log_f), ch1)
        sp.decommit_fri(8, (1 << log_f) - 1, pr1, ch1)
        assert ch.state == ch1.state and ch.proof == ch1.proof, "sharded-layer-0 transcript differs from the single-GPU transcript"
        print(json.dumps({"world": world, "cfg5_fri": {"log_domain": log_f, "seconds": t_fri, "transcript_state": ch.state,
                                                       "identical_to_single_gpu": True}}), flush=True)
    barrier()

    # ---------------- cfg5, whole prover: FibonacciSq trace of 2^23 - 1 rows (2^15 - 1 unless --full), domain 8x
    log_t = 23 if args.full else 15
    mg.stark101_prove_multi(sp, ctx, sp.Channel(P), 3141592, log_t, 3, 3, rank, world)      # warm-up: twiddles, pools, NCCL
    chm, phases = sp.Channel(P), {}
    barrier(); t0 = time.perf_counter()
    mg.stark101_prove_multi(sp, ctx, chm, 3141592, log_t, 3, 3, rank, world, timings=phases)
    barrier(); t_prove = time.perf_counter() - t0
    if rank == 0:
        sp.stark101_prove(ctx, sp.Channel(P), 3141592, log_t, 3, 3)
        ch1 = sp.Channel(P)
        t0 = time.perf_counter()
        sp.stark101_prove(ctx, ch1, 3141592, log_t, 3, 3)
        t_single = time.perf_counter() - t0
        assert chm.state == ch1.state and chm.proof == ch1.proof, "multi-GPU prover transcript differs from the single-GPU transcript"
        print(json.dumps({"world": world, "cfg5_prove": {"log_trace": log_t, "log_domain": log_t + 3, "seconds": t_prove,
                                                         "single_gpu_seconds": t_single, "rank0_phase_seconds": {k: round(v, 5) for k, v in phases.items()},
                                                         "transcript_state": chm.state,
                                                         "identical_to_single_gpu": True}}), flush=True)
    barrier()
    if world > 1:
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
