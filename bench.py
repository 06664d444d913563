#!/usr/bin/env python
"""bench.py — LDE + Merkle + FRI-commit throughput (Melem/s) and prove ms at a 2^24 domain on N B200s.

A step = one pass of the hot path over one synthetic polynomial (BASELINE.json configs[2]):
coset evaluation of a degree-(2^21-1) polynomial on the 2^24-point domain 5*<w>, 22 Merkle commitments,
21 fused fold-and-hash layers driven by the host Fiat-Shamir channel, and the openings of 32 queries
(`fri_commit` + `decommit_fri`, reference src/fri/fri_commit.rs:72-179).

  value     device-timed (CUDA events on the library's stream), coefficients already resident in HBM
  e2e       the same step through the host-buffer C ABI: pinned u64 coefficients -> stark_fri_commit (H2D inside) ->
            roots/openings back in the host channel;  e2e_full additionally copies every FRI layer back by value, as the
            reference's FRIProof returns them (fri_commit.rs:9-13, :117-121)
  prove     stark101_prove (trace -> LDE -> commit -> composition -> FRI -> queries) at a 2^24 domain, oracle beside it
  N > 1     one process per GPU (torchrun).  `value`: every rank commits its own column (weak scaling, replicas: the FRI
            loop is not partitioned, north-star).  `sharded`: the two configs that DO shard, at fixed total size (strong
            scaling), asserted against committed oracle goldens inside the run:
              cfg4  64 columns x 2^22 rows -> LDE 2^25 + Merkle commit per column, column c -> rank c mod N, roots gathered
              cfg5  one 2^26-point column: four-step LDE (NCCL all-to-all and peer-memory stores), leaf-range commit,
                    fri_commit with the sharded layer 0 + 8 openings
  --impl reference   the CPU oracle port of the reference algorithm (the Rust crate cannot be built here), all host
            threads, on the SAME 2^24 workload
"""
from __future__ import annotations

import argparse
import hashlib
import importlib
import json
import os
import socket
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

P = 3221225473
OFFSET = 5
QUERIES = 32
METRIC = "lde_merkle_fri_commit_throughput"
UNIT = "Melem/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log-n", type=int, default=24, help="log2 of the LDE/FRI domain (headline: 24)")
    ap.add_argument("--log-blowup", type=int, default=3)
    ap.add_argument("--cpu-log-n", type=int, default=None, help="CPU arm domain (default: the same as --log-n)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-timing", action="store_true")
    ap.add_argument("--no-pipelined", action="store_true")
    ap.add_argument("--no-sharded", action="store_true", help="skip the cfg4 / cfg5 strong-scaling block")
    ap.add_argument("--no-prove", action="store_true")
    ap.add_argument("--sharded-small", action="store_true", help="cfg4 16 x 2^16, cfg5 2^22 (quick check of the sharded block)")
    ap.add_argument("--profile-mode", action="store_true",
                    help="for runs under ncu: exactly --warmup warm-up steps and --steps steps of the device-resident path, nothing else")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------- helpers
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md).
    The sampler runs from before the warm-up; only rows stamped inside [mark_start, mark_stop] are used."""

    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.rows, self.proc, self.t0, self.t1 = device, [], None, None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", os.environ.get("BENCH_SMI_MS", "10"), "-i", str(self.device)],
                                         stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def mark_start(self): self.t0 = time.time()
    def mark_stop(self): self.t1 = time.time()

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, pw = [], None, set(), []
        for ts, r in self.rows:
            if self.t0 is not None and not (self.t0 - 0.02 <= ts <= (self.t1 or ts) + 0.05):
                continue
            try:
                sm.append(float(r[1])); smax = float(r[2]); pw.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_min_mhz": sm[0] if sm else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "power_w_max": max(pw) if pw else None, "samples": len(sm)}


def measured_peaks() -> tuple[float, str]:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def latest_profile(suffix: str):
    pdir = os.path.join(ROOT, "profiles")
    try:
        f = sorted(x for x in os.listdir(pdir) if x.endswith(suffix))[-1]
        return json.load(open(os.path.join(pdir, f))), f"profiles/{f}"
    except Exception:
        return None, None


def ncu_traffic() -> tuple:
    """DRAM bytes per hashing launch from the committed ncu capture (profiles/rNN_traffic.json), or None."""
    d, f = latest_profile("_traffic.json")
    if not d:
        return None, None
    return d["hashing_dram_bytes_per_launch"], f"{f}: {d['hashing_dram_bytes_per_step'] / 1e9:.2f} GB over {d['hashing_launches']} launches per step"


def build_record(sp) -> dict:
    """Was the library this run loaded compiled on this machine, or did it travel with the tree?"""
    rec = {"library": os.path.relpath(sp.LIB_PATH, ROOT), "host": socket.gethostname()}
    try:
        rec["sha256_16"] = hashlib.sha256(open(sp.LIB_PATH, "rb").read()).hexdigest()[:16]
        info = json.load(open(sp.LIB_PATH.replace(".so", ".build.json")))
        try:
            boot = open("/proc/sys/kernel/random/boot_id").read().strip()
        except Exception:
            boot = ""
        rec.update({"built_on": info.get("host"), "built_at": info.get("time"), "nvcc": info.get("nvcc"),
                    "compiled_on_this_host": info.get("host") == rec["host"] and info.get("boot_id", "?") == boot})
        if not rec["compiled_on_this_host"]:
            rec["note"] = "prebuilt libstark_b200.so shipped with the tree (sm_100a cross-compiled by __graft_entry__.build())"
    except Exception as e:
        rec["note"] = f"no build record next to the library ({e})"
    return rec


def layer_sizes(log_n: int, log_deg: int) -> list[int]:
    return [1 << (log_n - k) for k in range(log_deg + 1)]


def algorithmic_counts(log_n: int, log_deg: int) -> dict:
    """bytes at 8 B/element, 32 B/digest; 1384 int-ops per SHA-256 compression."""
    sizes = layer_sizes(log_n, log_deg)
    comp = sum(3 * n - 2 for n in sizes)
    lde = 8 * (1 << log_deg) + 8 * (1 << log_n)
    trees = sum(8 * n + 32 * (2 * n - 1) for n in sizes)
    folds = sum(8 * n + 8 * (n // 2) for n in sizes[:-1])
    fused = lde + sum(32 * (2 * n - 1) for n in sizes) + sum(8 * n for n in sizes[:-1]) + sum(8 * n for n in sizes[1:])
    return {"compressions": comp, "int_ops": 1384 * comp, "bytes_per_op_sum": lde + trees + folds, "bytes_fused_min": fused}


# ------------------------------------------------------------------------------------------- reference arm
def literal_tier(orc, log_n_full, log_blowup):
    """the reference's LITERAL algorithm (Horner over every layer's domain, single thread like
    the reference) timed where it finishes in about a second, with the analytic extrapolation to the full workload."""
    ln = 13
    c = orc.synthetic_poly_exact_degree(43, 1 << (ln - log_blowup), P)
    dom = orc.coset_domain(OFFSET, orc.root_of_unity(ln, P), 1 << ln, P)
    horner = lambda l: sum((1 << (l - k)) * (1 << (l - log_blowup - k)) for k in range(l - log_blowup + 1))
    t0 = time.perf_counter()
    pr = orc.fri_commit_literal(c, dom, orc.Channel(P), P)
    dt = time.perf_counter() - t0
    del pr
    ns = dt * 1e9 / horner(ln)
    return {"measured": f"fri_commit, literal tier, 2^{ln} domain, 1 thread: {dt * 1e3:.1f} ms", "ns_per_horner_step": ns,
            "horner_steps_full_workload": horner(log_n_full),
            "extrapolated_seconds_full_workload": ns * 1e-9 * horner(log_n_full),
            "note": "upper-bounds the reference's speed: its field multiply is a u128 remainder (element.rs:106), the oracle's is u64"}


def cpu_step(orc, coeffs, log_n, queries):
    """The reference's fri_commit + decommit_fri on the CPU (oracle port, NTT tier, retained trees)."""
    ch = orc.Channel(P)
    pr = orc.fri_commit_fast(coeffs, log_n, OFFSET, orc.root_of_unity(log_n, P), ch, P)
    orc.decommit_fri(queries, (1 << log_n) - 1, pr, ch)
    return ch


def cpu_time_steps(orc, log_n, log_blowup, warm, steps=None, budget_s=None):
    """ms per CPU step at a 2^log_n domain: `steps` timed steps, or as many as fit `budget_s` (at least 2)."""
    coeffs = orc.synthetic_poly_exact_degree(43, 1 << (log_n - log_blowup), P)
    for _ in range(warm):
        cpu_step(orc, coeffs, log_n, QUERIES)
    n, t0, state = 0, time.perf_counter(), None
    while True:
        state = cpu_step(orc, coeffs, log_n, QUERIES).state
        n += 1
        el = time.perf_counter() - t0
        if (steps is not None and n >= steps) or (steps is None and n >= 2 and el >= budget_s) or el > 120.0:
            break
    return el / n * 1e3, n, state


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyoracle as orc
    orc.build()
    orc.set_num_threads(len(os.sched_getaffinity(0)))      # torchrun exports OMP_NUM_THREADS=1: use every host core we may run on
    log_n = args.cpu_log_n or args.log_n
    log_deg = log_n - args.log_blowup
    ms, n_steps, state = cpu_time_steps(orc, log_n, args.log_blowup, max(args.warmup, 1), steps=args.steps)
    val = (1 << log_n) / (ms * 1e-3) / 1e6
    cores = orc.num_threads()
    same = log_n == args.log_n
    sample = (f"{n_steps} x fri_commit+decommit_fri at a 2^{log_n} domain (degree 2^{log_deg}-1, blowup {1 << args.log_blowup}, {QUERIES} queries)"
              + (": the full workload of the B200 arm" if same else f": 1/{1 << (args.log_n - log_n)} of the 2^{args.log_n} workload per step")
              + f"; oracle NTT tier, OpenMP, SHA-NI={bool(orc.lib().or_sha256_accel_active())}")
    extra = {}
    if same and log_n > 20:                               # the cache-resident figure of round 1, kept for comparison
        ms20, n20, _ = cpu_time_steps(orc, 20, args.log_blowup, 1, budget_s=3.0)
        extra = {"sample_2e20": {"value": (1 << 20) / (ms20 * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms20, "steps": n20}}
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": n_steps,
            "warmup": max(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"cfg3: fri_commit + decommit_fri, degree 2^{log_deg}-1 polynomial on the 2^{log_n} coset 5*<w>, "
                                   f"p=3221225473, {log_deg + 1} layers, {QUERIES} queries" + ("" if same else f" (bounded CPU sample of the 2^{args.log_n} workload)"),
                       "log_domain": args.log_n, "sample_log_domain": log_n, "blowup": 1 << args.log_blowup, "queries": QUERIES,
                       "same_config_as_b200_arm": same},
            "cpu_baseline": dict({"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                                  "literal": literal_tier(orc, args.log_n, args.log_blowup)}, **extra),
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "transcript_state": state,
            "note": "the reference is Rust nightly + un-vendored crates and cannot be built in this image; this is the C oracle port of its "
                    "algorithm (NTT tier: same bits as the literal Horner tier, which is O(N*d) and cannot reach this size)"}
    emit(line)


# ------------------------------------------------------------------------------------------- B200 arm
def pipelined_throughput(sp, synth, log_n, log_deg, n, device, instances=3, reps=5):
    out, barrier = {}, threading.Barrier(instances)

    def worker(k):
        ctx = sp.Context(P, sp.G_DEFAULT, device)
        c = ctx.upload(synth.synthetic_poly_exact_degree(143 + k, 1 << log_deg, P))
        dom = sp.CosetFri(ctx, OFFSET, log_n)

        def step():
            ch = sp.Channel(P)
            pr = sp.fri_commit(ctx, c, dom, ch)
            sp.decommit_fri(QUERIES, n - 1, pr, ch)
            pr.free()
        for _ in range(3):
            step()
        barrier.wait()
        t0 = time.perf_counter()
        for _ in range(reps):
            step()
        out[k] = time.perf_counter() - t0
        ctx.close()

    ths = [threading.Thread(target=worker, args=(k,)) for k in range(instances)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    wall = max(out.values())
    return {"instances": instances, "value": instances * reps * n / wall / 1e6, "unit": UNIT, "ms_per_instance": wall / (instances * reps) * 1e3,
            "timing": "host wall clock around synchronous calls (every call ends in a stream sync)",
            "note": "independent polynomials committed concurrently on one GPU from separate host threads / contexts / streams"}


class Timer:
    """CUDA events on the library's stream around a host-driven phase; max over ranks."""

    def __init__(self, torch, dist, stream, world, local):
        self.torch, self.dist, self.stream, self.world, self.local = torch, dist, stream, world, local

    def sync_all(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def run(self, fn):
        torch = self.torch
        self.sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        out = fn()
        e1.record(self.stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), out

    def max_over_ranks(self, ms: float) -> float:
        if self.world == 1:
            return ms
        t = self.torch.tensor([ms], device=f"cuda:{self.local}")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    @staticmethod
    def _release(out):
        for o in (out if isinstance(out, tuple) else (out,)):
            if hasattr(o, "free"):
                o.free()

    def best(self, fn, reps: int, keep=None):
        """min over `reps` of the max-over-ranks time; returns (ms, last result); earlier results are freed."""
        best, last = None, None
        for _ in range(reps):
            if last is not None:
                self._release(last)
            ms, out = self.run(fn)
            ms = self.max_over_ranks(ms)
            best = ms if best is None else min(best, ms)
            last = out
        return best, last


def hash_roofline(achieved, mix_peak, note):
    """Leaf-dominated launches execute fewer instructions than the algorithmic 1384 per compression (the leaf block is
    specialised: ~75 fewer; the padding block of a parent hash has a precomputed schedule), so achieved / peak can pass 1;
    the fraction is capped and the executed-instruction view (ncu: ALU pipe cycles active) printed beside it."""
    pipe, src = latest_profile("_pipe_util.json")
    raw = achieved / mix_peak if mix_peak else None
    return {"bound": "int", "achieved": achieved, "peak": mix_peak, "unit": "Tint-op/s per GPU",
            "frac": min(raw, 1.0) if raw is not None else None, "frac_uncapped": raw,
            "executed_alu_pipe_pct": (pipe or {}).get("merkle_subtree_kernel<VALUES>"), "executed_alu_pipe_pct_source": src, "note": note}


def sharded_block(args, sp, synth, ctx, tm: Timer, rank, world, local, mix_peak):
    """The two BASELINE configs that shard, at fixed total size, asserted against the committed oracle goldens."""
    import numpy as np
    import torch
    gold_path = os.path.join(ROOT, "tests", "golden", "sharded.json")
    gold = json.load(open(gold_path))
    g4, g5 = (gold["cfg4_small"], gold["cfg5_small"]) if args.sharded_small else (gold["cfg4"], gold["cfg5"])
    mg = sp.MultiGpu.from_torch(ctx)
    out = {"world": world, "scaling": "strong", "goldens": "tests/golden/sharded.json (CPU oracle, tests/golden/make_golden_sharded.py)",
           "timing": "CUDA events on the library's stream around each phase, max over ranks, best of the repetitions"}

    # ---------------- cfg4: 64 columns x 2^22 rows, column c -> rank c mod N
    log_rows, n_cols, log_b = g4["log_rows"], g4["n_cols"], g4["log_blowup"]
    n_rows, N4 = 1 << log_rows, 1 << (log_rows + log_b)
    mine = list(range(rank, n_cols, world))
    pin = torch.empty((max(len(mine), 1), n_rows), dtype=torch.int64).pin_memory()
    pin_np = pin.numpy().view(np.uint64)
    cols = {}
    for k, c in enumerate(mine):
        pin_np[k, :] = synth.synthetic_column(g4["seed_base"] + c, n_rows, P)
        cols[c] = pin_np[k]
    commit = lambda: mg.commit_columns(cols, n_cols, log_rows, g4["offset_in"], log_b, g4["offset_out"])
    commit()                                                            # warm-up: twiddles, pool blocks
    reps4 = 3 if world > 1 else 2
    ms4, roots = tm.best(commit, reps4)
    assert [r.hex() for r in roots] == g4["roots"], "cfg4: the gathered roots differ from the oracle's goldens"
    comp4 = n_cols * (3 * N4 - 2)
    out["cfg4"] = {"workload": f"{n_cols} columns x 2^{log_rows} rows -> coset LDE 2^{log_rows + log_b} + Merkle commit per column, column c -> rank c mod N, "
                               f"{n_cols} roots all-gathered", "ms": ms4, "Melem_per_s": n_cols * N4 / (ms4 * 1e-3) / 1e6,
                   "roots_equal_golden": True, "roots_sha256": g4["roots_sha256"], "columns_per_gpu": len(mine),
                   "h2d_bytes_per_gpu": len(mine) * n_rows * 8, "collective": f"ncclAllGather of {32 * -(-n_cols // world)} bytes per rank",
                   "roofline": hash_roofline(1384.0 * comp4 / world / (ms4 * 1e-3) / 1e12, mix_peak,
                                             "hashing int-ops of the whole phase (upload, LDE and root gather included in the time) per GPU")}
    del cols, pin, pin_np

    # ---------------- cfg5: one 2^26-point column
    log_n, q5 = g5["log_n"], g5["queries"]
    N5, blk = 1 << log_n, (1 << log_n) // world
    cvec = ctx.upload(synth.synthetic_poly_exact_degree(g5["seed"], 1 << (log_n - 3), P))
    exch_bytes = 2 * 4 * blk * (world - 1) // world                      # two exchanges, (G-1)/G of the rank's 4*N/G bytes each
    c5 = {"workload": f"degree 2^{log_n - 3}-1 polynomial on the 2^{log_n} coset 5*<w>: four-step LDE, leaf-range commit, fri_commit with the sharded "
                      f"layer 0 + {q5} openings", "exchange_bytes_per_gpu_per_transform": exch_bytes}
    for name, transport in (("lde_nccl_all_to_all", 0), ("lde_peer_memory", 1)):
        run = lambda: mg.fourstep_lde(cvec, log_n, g5["offset"], transport)
        run().free()                                                    # warm-up (allocations, IPC mappings, NCCL channels)
        ms, b = tm.best(run, 5, keep=lambda v: None)
        tree, root, subs = mg.commit_leaf_ranges(b)
        assert root.hex() == g5["roots"][0], f"cfg5 ({name}): the root of the distributed column differs from the oracle's golden"
        tree.free()
        c5[name] = {"ms": ms, "Melem_per_s": N5 / (ms * 1e-3) / 1e6, "root_equals_golden": True,
                    "exchange_GBps_per_gpu": (exch_bytes / (ms * 1e-3) / 1e9) if world > 1 else None,
                    "note": "GB/s = exchanged bytes / WHOLE transform time (both NTT phases included), a lower bound on the link rate"}
    blkv = mg.fourstep_lde(cvec, log_n, g5["offset"], 1)

    def commit_ranges():
        t, root, subs = mg.commit_leaf_ranges(blkv)
        t.free()
        return root
    commit_ranges()
    ms_c, root = tm.best(commit_ranges, 3)
    assert root.hex() == g5["roots"][0]
    compl = 3 * N5 - 2
    c5["leaf_range_commit"] = {"ms": ms_c, "Melem_per_s": N5 / (ms_c * 1e-3) / 1e6, "root_equals_golden": True,
                               "roofline": hash_roofline(1384.0 * compl / world / (ms_c * 1e-3) / 1e12, mix_peak, "one tree of 2^%d leaves per GPU" % (log_n - (world.bit_length() - 1)))}
    blkv.free()

    def fri():
        ch = sp.Channel(P) if rank == 0 else None
        f = mg.fri_commit(cvec, log_n, g5["offset"], ch, 1)
        mg.decommit_fri(f, q5, N5 - 1, ch)
        return f, ch
    f, ch = fri()
    f.free()
    ms_f, (f, ch) = tm.best(fri, 3, keep=None)
    if rank == 0:
        pr = f.proof
        assert pr.num_layers == g5["num_layers"] and [pr.tree(k).root() for k in range(1, pr.num_layers)] == g5["roots"][1:]
        assert ch.state == g5["final_state"] and hashlib.sha256(ch.proof_flat()).hexdigest() == g5["proof_sha256"], \
            "cfg5: the transcript of the sharded fri_commit + decommit_fri differs from the oracle's golden"
    f.free()
    c5["fri_commit_and_openings"] = {"ms": ms_f, "Melem_per_s": N5 / (ms_f * 1e-3) / 1e6, "transcript_equals_golden": True,
                                     "transcript_state": g5["final_state"], "layers": g5["num_layers"], "queries": q5,
                                     "note": "four-step LDE + leaf-range commit of layer 0; at N > 1 layer 0 is all-gathered, every rank replicates "
                                             "the folds (HBM-bound, cheap) and the layers of >= 2^21 leaves are hashed in leaf ranges too (32-byte "
                                             "subtree roots gathered, beta broadcast); the smaller layers, the channel and their openings stay on "
                                             "rank 0; one exchange per query for every leaf-range opening"}
    if world > 1:
        # the same call with the leaf hashing of layers >= 1 left on rank 0 (round 1 / first half of round 2), for comparison
        os.environ["STARK_MG_FRI_SHARD"] = "0"
        try:
            f, ch = fri()
            f.free()
            ms_f0, (f, ch) = tm.best(fri, 2, keep=None)
            if rank == 0:
                assert ch.state == g5["final_state"], "cfg5: transcript with the layers >= 1 on rank 0 differs from the golden"
            f.free()
        finally:
            del os.environ["STARK_MG_FRI_SHARD"]
        c5["fri_commit_and_openings"]["ms_with_layers_ge1_on_rank0"] = ms_f0
    # ---------------- cfg5, the whole prover: FibonacciSq trace of 2^23 - 1 rows, domain 2^26, 3 queries
    gp = gold.get("prove5_small" if args.sharded_small else "prove5")
    if gp:
        def prove_all():
            chp = sp.Channel(P) if rank == 0 else None
            mg.stark101_prove(chp, gp["a1"], gp["log_trace"], gp["log_blowup"], gp["queries"], 1)
            return chp
        prove_all()
        ms_p, chp = tm.best(prove_all, 2)
        if rank == 0:
            assert chp.state == gp["final_state"] and hashlib.sha256(chp.proof_flat()).hexdigest() == gp["proof_sha256"], \
                "cfg5: the transcript of the sharded prover differs from the oracle's golden"
        c5["prove"] = {"ms": ms_p, "what": f"stark_mg_stark101_prove: trace of 2^{gp['log_trace']}-1 rows (sequential recurrence on the host, replicated), "
                                           f"four-step LDE, leaf-range commitments of f and CP, composition on the local range, large FRI layers hashed in leaf ranges, "
                                           f"{gp['queries']} queries", "transcript_equals_golden": True, "transcript_state": gp["final_state"]}
    # ---------------- the same workloads on ONE GPU, measured by rank 0 in this run (the other ranks wait): strong-scaling reference
    tm.sync_all()
    if world > 1:
        single = {}
        if rank == 0:
            solo = Timer(torch, None, tm.stream, 1, local)
            g1 = sp.MultiGpu(ctx, 0, 1)
            ev = lambda: ctx.coset_evaluate_dev(cvec, log_n, g5["offset"])
            ev().free()
            single["lde_ms"], e = solo.best(ev, 3, keep=lambda v: None)
            mk = lambda: sp.MerkleTree.new(ctx, e)
            mk().free()
            single["commit_ms"], t = solo.best(mk, 2, keep=lambda v: None)
            assert t.root_bytes().hex() == g5["roots"][0]
            t.free(); e.free()

            def fri1():
                ch1 = sp.Channel(P)
                pr1 = sp.fri_commit(ctx, cvec, sp.CosetFri(ctx, g5["offset"], log_n), ch1)
                sp.decommit_fri(q5, N5 - 1, pr1, ch1)
                return pr1, ch1
            fri1()[0].free()
            single["fri_ms"], (pr1, ch1) = solo.best(fri1, 2, keep=None)
            assert ch1.state == g5["final_state"]
            pr1.free()
            if gp:
                def prove1():
                    c1 = sp.Channel(P)
                    sp.stark101_prove(ctx, c1, gp["a1"], gp["log_trace"], gp["log_blowup"], gp["queries"])
                    return c1
                prove1()
                single["prove_ms"], c1 = solo.best(prove1, 2)
                assert c1.state == gp["final_state"]
            # one cfg4 column on one GPU x n_cols (columns are independent: the single-GPU time of the whole config)
            k1 = min(n_cols, 8)                                   # enough columns for the upload / hashing overlap to show
            colp = torch.empty((k1, n_rows), dtype=torch.int64).pin_memory()
            colp_np = colp.numpy().view(np.uint64)
            for c in range(k1):
                colp_np[c, :] = synth.synthetic_column(g4["seed_base"] + c, n_rows, P)
            some = lambda: g1.commit_columns({c: colp_np[c] for c in range(k1)}, k1, log_rows, g4["offset_in"], log_b, g4["offset_out"])
            some()
            ms1, r1 = solo.best(some, 3)
            assert [r.hex() for r in r1] == g4["roots"][:k1]
            single["cfg4_ms"] = ms1 * n_cols / k1
            single["cfg4_note"] = f"{n_cols}/{k1} x the measured time of {k1} columns ({ms1:.3f} ms) on one GPU (columns are independent)"
            g1.close()
        tm.sync_all()
        if rank == 0:
            out["single_gpu_reference"] = single
            out["cfg4"]["speedup_vs_1gpu"] = single["cfg4_ms"] / ms4
            out["cfg4"]["efficiency"] = single["cfg4_ms"] / ms4 / world
            c5["lde_peer_memory"]["speedup_vs_1gpu"] = single["lde_ms"] / c5["lde_peer_memory"]["ms"]
            c5["lde_nccl_all_to_all"]["speedup_vs_1gpu"] = single["lde_ms"] / c5["lde_nccl_all_to_all"]["ms"]
            c5["leaf_range_commit"]["speedup_vs_1gpu"] = single["commit_ms"] / ms_c
            c5["leaf_range_commit"]["efficiency"] = single["commit_ms"] / ms_c / world
            c5["fri_commit_and_openings"]["speedup_vs_1gpu"] = single["fri_ms"] / ms_f
            if gp:
                c5["prove"]["speedup_vs_1gpu"] = single["prove_ms"] / c5["prove"]["ms"]
    out["cfg5"] = c5
    cvec.free()
    mg.close()
    return out


def prove_block(sp, ctx, tm: Timer, log_n, log_blowup, with_oracle: bool):
    """`prove ms at a 2^24 domain` (BASELINE.json metric): stark101_prove, the build-defined FibonacciSq prover of DESIGN.md
    cfg1 scaled to a 2^(log_n - log_blowup) - 1 row trace, 3 queries; transcript asserted equal to the CPU oracle's."""
    log_trace, q = log_n - log_blowup, 3

    def prove():
        ch = sp.Channel(P)
        sp.stark101_prove(ctx, ch, 3141592, log_trace, log_blowup, q)
        return ch
    prove()
    ms, ch = tm.best(prove, 3)
    t0 = time.perf_counter()
    prove()
    wall = (time.perf_counter() - t0) * 1e3
    out = {"what": f"stark101_prove: FibonacciSq trace of 2^{log_trace}-1 rows, LDE/FRI domain 2^{log_n}, {q} queries (trace generation on the host included)",
           "ms": ms, "wall_ms": wall, "transcript_state": ch.state, "proof_bytes": ch.proof_size()}
    if with_oracle:
        from oracle import pyoracle as orc
        orc.build()
        orc.set_num_threads(len(os.sched_getaffinity(0)))
        och = orc.Channel(P)
        t0 = time.perf_counter()
        orc.stark101_prove(och, 3141592, log_trace, log_blowup, sp.G_DEFAULT, q, literal=False)
        out["oracle_ms"] = (time.perf_counter() - t0) * 1e3
        out["oracle_cores"] = orc.num_threads()
        assert och.state == ch.state and och.proof_flat() == ch.proof_flat(), "stark101_prove: GPU transcript differs from the oracle's"
        out["transcript_equals_oracle"] = True
        claimed = int.from_bytes(ch.proof[0][40:48], "big")
        ok, why = sp.stark101_verify(ch.proof_flat(), claimed, log_trace, log_blowup, q)
        assert ok, why
        out["verifier_accepts"] = True
    return out


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sp = importlib.import_module("stark-prover_b200")
    synth = importlib.import_module("stark-prover_b200.synthetic")

    log_n, log_deg = args.log_n, args.log_n - args.log_blowup
    n = 1 << log_n
    ctx = sp.Context(P, sp.G_DEFAULT, local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local)
    tm = Timer(torch, dist, stream, world, local)
    coeffs = synth.synthetic_poly_exact_degree(43 + rank, 1 << log_deg, P)
    pinned = torch.empty(1 << log_deg, dtype=torch.int64).pin_memory()
    pinned_np = pinned.numpy().view(np.uint64)
    pinned_np[:] = coeffs
    dev_coeffs = ctx.upload(coeffs)
    domain = sp.CosetFri(ctx, OFFSET, log_n)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")   # > 126 MB L2
    roots_out = torch.zeros(world, 32, dtype=torch.uint8, device=f"cuda:{local}") if world > 1 else None

    per_rank_ms = []

    def step(src, layers_out=None, wait=True):
        ch = sp.Channel(P)
        pr = sp.fri_commit(ctx, src, domain, ch, layers_out=layers_out, wait=wait)
        sp.decommit_fri(QUERIES, n - 1, pr, ch)
        if layers_out is not None and not wait:
            pr.layers_wait()                     # the last layer copies ran under the openings
        return pr, ch

    def gather_roots(pr):
        if world == 1:
            return
        mine = torch.frombuffer(bytearray(pr.tree(0).root_bytes()), dtype=torch.uint8).to(f"cuda:{local}")
        dist.all_gather_into_tensor(roots_out.view(-1), mine)

    def timed(src, steps, flush_l2=True, after=None, layers_out=None, wait=True):
        """max-over-ranks device time per step (ms) and the last step's artefacts.  `after(pr)`: extra work inside the
        timed region (the layers copied back by value after the step); `layers_out`: the layers streamed to that pinned
        buffer during the commit (e2e_full)."""
        total = 0.0
        pr = ch = None
        for _ in range(steps):
            if pr is not None:
                pr.free()
            if flush_l2:
                with torch.cuda.stream(stream):
                    flush.zero_()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            pr, ch = step(src, layers_out, wait)
            if after is not None:
                after(pr)
            gather_roots(pr)
            if world > 1:
                stream.wait_stream(torch.cuda.current_stream())
            e1.record(stream)
            torch.cuda.synchronize()
            total += e0.elapsed_time(e1)
        ms = total / max(steps, 1)
        if world > 1:
            t = torch.tensor([ms], device=f"cuda:{local}")
            every = torch.zeros(world, device=f"cuda:{local}")
            dist.all_gather_into_tensor(every, t)
            per_rank_ms[:] = [round(float(x), 3) for x in every.tolist()]
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, pr, ch

    if args.profile_mode:
        # (main() has set STARK_OPEN_SERVER=0: the resident opening server talks to the host while it runs, which a profiler
        # that serialises and replays kernels cannot follow; under ncu the queries are one launch each, as before)
        timed(dev_coeffs, args.warmup, flush_l2=False)
        ms, pr, ch = timed(dev_coeffs, args.steps, flush_l2=False)
        emit({"profile_mode": True, "ms_per_step_under_profiler": ms, "launches": ctx.launch_count})
        pr.free()
        ctx.close()
        return
    # ---- warm-up, then the device-resident timed region
    # one sampler per job (rank 0's GPU): a polling nvidia-smi per rank perturbs the host-bound opening phase
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    timed(dev_coeffs, max(args.warmup, 3))
    l0 = ctx.launch_count
    sampler.mark_start()
    ms_dev, pr, ch = timed(dev_coeffs, args.steps)
    per_rank_dev = list(per_rank_ms)
    sampler.mark_stop()
    launches = (ctx.launch_count - l0) // args.steps
    clocks = sampler.stop()
    final_state = ch.state
    n_layers = pr.num_layers
    d2h = 32 * n_layers + 8 * n_layers + len(pr.open([0])) * QUERIES
    layer_lens = [pr.layer_len(k) for k in range(n_layers)]
    pr.free()

    # ---- where the step's wall time goes on the host side (one extra, untimed-for-the-metric step)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    chx = sp.Channel(P)
    prx = sp.fri_commit(ctx, dev_coeffs, domain, chx)
    t1 = time.perf_counter()
    sp.decommit_fri(QUERIES, n - 1, prx, chx)
    t2 = time.perf_counter()
    prx.free()
    host_breakdown = {"fri_commit_ms": (t1 - t0) * 1e3, "decommit_fri_ms": (t2 - t1) * 1e3}

    # ---- end to end: pinned host coefficients through the host-buffer ABI
    timed(pinned_np, 1)
    ms_e2e, pr2, ch2 = timed(pinned_np, args.steps)
    assert ch2.state == final_state, "device-resident and host-buffer paths disagree"
    pr2.free()
    # ---- end to end with FRIProof BY VALUE: every layer copied back as u64 (fri_commit.rs:117-121 returns Vec<Vec<FieldElement>>)
    layers_pin = torch.empty(sum(layer_lens), dtype=torch.int64).pin_memory()
    layers_np = layers_pin.numpy().view(np.uint64)

    def copy_layers_back(p):
        off = 0
        for k, ln in enumerate(layer_lens):
            p.layer(k, 0, ln, out=layers_np[off:off + ln])
            off += ln
    timed(pinned_np, 1, after=copy_layers_back)
    ms_full_after, pr4, _ = timed(pinned_np, max(2, min(args.steps, 5)), after=copy_layers_back)
    assert int(layers_np[0]) == int(pr4.layer(0, 0, 1)[0])
    pr4.free()
    # the same by-value result with the copies issued DURING the commit: stark_fri_commit_to_host widens and copies every layer
    # on a second, high-priority stream while the main stream hashes the following layers; complete when fri_commit returns
    check = hashlib.sha256(layers_np.tobytes()).hexdigest()
    layers_np[:] = 0
    timed(pinned_np, 1, layers_out=layers_np)
    ms_full_sync, pr5, ch5 = timed(pinned_np, max(2, min(args.steps, 5)), layers_out=layers_np)
    assert ch5.state == final_state, "the by-value path changed the transcript"
    assert hashlib.sha256(layers_np.tobytes()).hexdigest() == check, "streamed layers differ from stark_fri_layer_read"
    pr5.free()
    # ... and with the wait moved behind the openings (stark_fri_commit_to_host_async + stark_fri_layers_wait): the layers are
    # complete when the STEP ends, the tail of the copies runs under decommit_fri
    layers_np[:] = 0
    ms_full, pr5, ch5 = timed(pinned_np, max(2, min(args.steps, 5)), layers_out=layers_np, wait=False)
    assert ch5.state == final_state, "the by-value path changed the transcript"
    assert hashlib.sha256(layers_np.tobytes()).hexdigest() == check, "streamed layers differ from stark_fri_layer_read"
    pr5.free()
    del layers_pin, layers_np

    line = {"metric": METRIC, "value": world * n / (ms_dev * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "dtype_note": "u32 Montgomery field arithmetic (p = 3221225473 < 2^32; u64 at the ABI) and 32-bit SHA-256 words",
            "data": "synthetic",
            "config": {"workload": f"cfg3: fri_commit + decommit_fri, degree 2^{log_deg}-1 polynomial on the 2^{log_n} coset 5*<w>, "
                                   f"p=3221225473, {n_layers} layers, {QUERIES} queries; one column per GPU (replicas: the FRI loop is not partitioned)",
                       "log_domain": log_n, "blowup": 1 << args.log_blowup, "queries": QUERIES, "layers": n_layers,
                       "l2": "256 MiB buffer written between timed iterations; layer 0 + its tree = 576 MiB > L2",
                       "fri_commit_decommit_ms": ms_dev,
                       "fri_layers": "device-resident handles (stark_fri_layer_read copies ranges on demand); e2e_full returns all of them by value (streamed during the commit)"},
            "e2e": {"value": world * n / (ms_e2e * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": 8 << log_deg, "d2h_bytes_per_step": d2h},
            "e2e_full": {"value": world * n / (ms_full * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms_full,
                         "h2d_bytes_per_step": 8 << log_deg, "d2h_bytes_per_step": d2h + 8 * sum(layer_lens),
                         "note": "e2e plus every FRI layer in pinned host memory as u64 when the step ends, i.e. FRIProof.fri_layers by "
                                 "value: stark_fri_commit_to_host_async streams them on a second stream under the hashing of the following "
                                 "layers and the openings (stark_fri_layers_wait at the end of the step)",
                         "ms_per_step_layers_complete_when_fri_commit_returns": ms_full_sync,
                         "ms_per_step_copied_after_the_step": ms_full_after},
            "gpu_launches": int(launches), "clocks": clocks, "transcript_state": final_state,
            "ms_per_rank": per_rank_dev if world > 1 else None,
            "host_breakdown_ms": host_breakdown}

    # ---- per-kernel durations, live, CUDA events on the launching stream (separate instrumented steps; every rank runs
    # them so that the collectives inside `timed` stay matched, rank 0 reports)
    hbm_peak, peak_src = measured_peaks()
    alg = algorithmic_counts(log_n, log_deg)
    mix_peak = None
    if not args.no_kernel_timing:
        ctx.set_timing(True)
        ctx.read_timing()
        ksteps = max(2, min(args.steps, 5))
        ms_instr, pr3, _ = timed(dev_coeffs, ksteps)
        kt = ctx.read_timing()
        ctx.set_timing(False)
        pr3.free()
        alu_peak, mix_peak = ctx.measure_int_peak()
        hash_ms = (kt["merkle_leaf"]["ms"] + kt["merkle_node"]["ms"]) / ksteps
        hash_ops = (kt["merkle_leaf"]["units"] + kt["merkle_node"]["units"]) / ksteps
        hash_launches = (kt["merkle_leaf"]["launches"] + kt["merkle_node"]["launches"]) // ksteps
        achieved = hash_ops / (hash_ms * 1e-3) / 1e12
        pipe, pipe_src = latest_profile("_pipe_util.json")
        line["roofline"] = {
            "kernel": "merkle_subtree_kernel<VALUES|FOLD|DIGESTS> (fused fold + leaf hash, 3 levels per launch) + merkle_tail_kernel",
            "bound": "int", "achieved": achieved, "peak": mix_peak, "unit": "Tint-op/s", "frac": achieved / mix_peak,
            "peak_source": "stark_measure_int_peak on this GPU: register chains of SHF/LOP3/IADD3 (ALU pipe) and IMAD (FMA pipe) "
                           "issued 3:1, the mix of the algorithmic count (1024 rotate/logic instructions : 360 adds per "
                           f"compression); the ALU pipe alone peaks at {alu_peak:.1f}, and 1024 of the 1384 can only run there",
            "alu_pipe_peak": alu_peak, "frac_of_alu_pipe_peak": achieved / alu_peak,
            "executed_alu_pipe_pct": pipe, "executed_alu_pipe_pct_source": pipe_src,
            "traffic": ncu_traffic()[0], "traffic_source": ncu_traffic()[1],
            "launches_per_step": hash_launches, "kernel_ms_per_step": hash_ms,
            "share_of_step": hash_ms / ms_instr, "rank": rank,
            "algorithmic": "1384 int-ops per SHA-256 compression; leaf = 1, node = 2 compressions (SURVEY.md 8d)",
            "note": "achieved counts the ALGORITHMIC 1384 instructions per compression against the two-pipe issue peak; the kernel "
                    "executes fewer on the leaf launches (specialised leaf block) and more on node launches (1384 + 904 per parent); "
                    "`executed_alu_pipe_pct` is ncu's sm__pipe_alu_cycles_active per launch class, the executed-instruction view"}
        ntt_ms = kt["ntt"]["ms"] / ksteps
        ntt_gbs = kt["ntt"]["units"] / ksteps / (ntt_ms * 1e-3) / 1e9 if ntt_ms else None
        # the fused fold+hash launches also stream every layer once: algorithmic bytes of those launches
        leaf_bytes = alg["bytes_fused_min"] - (8 * (1 << log_deg) + 8 * n)
        ntt_ops = (1 << args.log_blowup) * ((1 << log_deg) // 2 * log_deg * 12 + 6 * (1 << log_deg))
        line["roofline_hbm"] = {
            "ntt": {"bound": "hbm", "achieved": ntt_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ntt_gbs / hbm_peak if ntt_gbs else None,
                    "kernel_ms_per_step": ntt_ms, "launches_per_step": kt["ntt"]["launches"] // ksteps,
                    "algorithmic": "8 B per coefficient read + 8 B per evaluation written (SURVEY.md 8d LDE n->N)",
                    # (n/2) log2 n butterflies x (mul-mod + add-mod + sub-mod); counted as 12 int-ops per butterfly
                    # (the round-1 instruction count; the carry-predicated butterfly of round 2 executes 9): in this 32-bit field
                    # the transform is integer-bound, not HBM-bound
                    "int": {"achieved": ntt_ops / (ntt_ms * 1e-3) / 1e12, "peak": mix_peak, "unit": "Tint-op/s",
                            "frac": ntt_ops / (ntt_ms * 1e-3) / 1e12 / mix_peak,
                            "algorithmic": "12 int-ops per butterfly, (n/2)*log2(n) butterflies per size-n transform, "
                                           "2^blowup transforms + 6 per point for the coset shift"} if ntt_ms else None},
            "fold_and_hash": {"bound": "hbm", "achieved": leaf_bytes / (hash_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                              "frac": leaf_bytes / (hash_ms * 1e-3) / 1e9 / hbm_peak,
                              "note": "same launches as `roofline`: integer-bound, HBM fraction shown for completeness"},
            "peak_source": peak_src}
        line["kernel_ms"] = {k: v["ms"] / ksteps for k, v in kt.items()}
        line["instrumented_ms_per_step"] = ms_instr
    line["algorithmic"] = alg
    line["build"] = build_record(sp)

    # ---- the configs that shard (strong scaling), every N including 1
    if not args.no_sharded:
        line["sharded"] = sharded_block(args, sp, synth, ctx, tm, rank, world, local, mix_peak)

    if rank == 0:
        # ---- prove ms at the headline domain (BASELINE.json metric), oracle beside it
        if not args.no_prove:
            try:
                line["prove"] = prove_block(sp, ctx, Timer(torch, None, stream, 1, local), log_n, args.log_blowup,
                                            with_oracle=(not args.no_cpu_baseline) and world == 1)
            except AssertionError:
                raise
            except Exception as e:
                line["prove"] = {"error": str(e)}
        # ---- saturated throughput: independent instances from separate host threads/contexts on the same GPU
        # (one instance's host-bound openings overlap another's device-bound commit); informational, `value` stays
        # the single-instance figure
        if world == 1 and not args.no_pipelined:
            try:
                line["pipelined"] = pipelined_throughput(sp, synth, log_n, log_deg, n, local, instances=3, reps=max(3, min(args.steps, 6)))
            except Exception as e:                                   # never let the extra measurement break the line
                line["pipelined"] = {"error": str(e)}
        # ---- CPU baseline beside it: the SAME workload, all host threads, 10-30 s of CPU work
        if not args.no_cpu_baseline and world == 1:
            from oracle import pyoracle as orc          # the CPU baseline leg is the only use of the oracle in this arm
            orc.build()
            orc.set_num_threads(len(os.sched_getaffinity(0)))
            cl = args.cpu_log_n or log_n
            ms_cpu, reps, cstate = cpu_time_steps(orc, cl, args.log_blowup, 1, budget_s=10.0)
            if cl == log_n:
                assert cstate == final_state, "the CPU oracle's transcript of the benchmark workload differs from the GPU's"
            line["cpu_baseline"] = {"value": (1 << cl) / (ms_cpu * 1e-3) / 1e6, "unit": UNIT, "cores": orc.num_threads(), "kind": "port",
                                    "ms_per_step": ms_cpu, "same_config": cl == log_n,
                                    "transcript_equals_gpu": (cstate == final_state) if cl == log_n else None,
                                    "sample": f"{reps} x fri_commit+decommit_fri at a 2^{cl} domain"
                                              + (" (the full workload)" if cl == log_n else f" (1/{1 << (log_n - cl)} of the workload)")
                                              + f", oracle NTT tier + OpenMP, SHA-NI={bool(orc.lib().or_sha256_accel_active())}",
                                    "literal": literal_tier(orc, log_n, args.log_blowup)}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The one JSON line, on the process's original stdout."""
    txt = json.dumps(line) + "\n"
    if _REAL_STDOUT is None:
        sys.stdout.write(txt); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, txt.encode())


def main():
    global _REAL_STDOUT
    args = parse()
    # stdout carries exactly one JSON line: anything a library prints there (NCCL's version banner, ...) goes to stderr
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.profile_mode:
        os.environ["STARK_OPEN_SERVER"] = "0"       # read once by the library when the first decommit runs
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
