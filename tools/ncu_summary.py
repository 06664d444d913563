#!/usr/bin/env python
"""Summarise ncu artefacts into profiles/ (tracked).  No GPU needed: reads what gpurun brought back.

  python tools/ncu_summary.py launches gpurun_out/launches_r1d.csv profiles/r01_launches.md
  python tools/ncu_summary.py full gpurun_out/prof_merkle_r1c.ncu-rep profiles/r01_merkle_full.md
  python tools/ncu_summary.py traffic gpurun_out/traffic_r1.csv profiles/r01_traffic.json
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__sass_inst_executed_op_shared_ld.sum"]


def short(name: str) -> str:
    return re.sub(r"\(.*", "", name).replace("void ", "").replace("starkb200::", "")


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    rows = [r for r in csv.DictReader(lines) if r["Metric Name"] == "gpu__time_duration.sum"]
    half = len(rows) // 2           # --profile-mode runs 1 warm-up step + 1 step: keep the second
    step = rows[half:]
    agg, tot = collections.OrderedDict(), 0.0
    for r in step:
        t = float(r["Metric Value"].replace(",", "")) / 1e3
        a = agg.setdefault(short(r["Kernel Name"]), [0, 0.0])
        a[0] += 1; a[1] += t; tot += t
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({src}): one step of `bench.py --profile-mode --steps 1 --warmup 1`\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none`; per-launch times are cold-cache and serialised,\n"
                "so compare SHARES with the live CUDA-event numbers in the bench JSON, not absolutes.\n\n")
        f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1]:.1f} | {v[1] / tot:.3f} |\n")
        f.write(f"| **all** | {len(step)} | {tot:.1f} | 1.000 |\n\n## launches of the step, in order\n\n| # | kernel | grid | block | us |\n|---:|---|---|---|---:|\n")
        for i, r in enumerate(step):
            f.write(f"| {i} | `{short(r['Kernel Name'])}` | {r['Grid Size']} | {r['Block Size']} | {float(r['Metric Value'].replace(',', '')) / 1e3:.1f} |\n")
    print("wrote", dst)


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src})\n\n`ncu --set full --clock-control none --import-source on`; raw page, selected metrics.\n")
        for r in rows[2:]:
            f.write(f"\n## `{short(r[hdr.index('Kernel Name')])}` grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"| {k} | {r[i]} | {units[i]} |\n")
    print("wrote", dst)


def traffic(src, dst):
    """DRAM bytes per kernel over the second (= measured) step of `bench.py --profile-mode --steps 1 --warmup 1` captured with
    --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum."""
    import json
    lines = [l for l in open(src) if not l.startswith("==")]
    per_id = collections.OrderedDict()
    for r in csv.DictReader(lines):
        d = per_id.setdefault(int(r["ID"]), {"kernel": short(r["Kernel Name"])})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r["Metric Unit"], 1.0)
    rows = list(per_id.values())
    step = rows[len(rows) // 2:]
    per_kernel = collections.OrderedDict()
    for r in step:
        k = per_kernel.setdefault(r["kernel"], {"launches": 0, "dram_read": 0.0, "dram_write": 0.0, "us": 0.0})
        k["launches"] += 1
        k["dram_read"] += r.get("dram__bytes_read.sum", 0.0)
        k["dram_write"] += r.get("dram__bytes_write.sum", 0.0)
        k["us"] += r.get("gpu__time_duration.sum", 0.0) / 1e3
    hashing = [v for k, v in per_kernel.items() if k.startswith("merkle_subtree_kernel") or k.startswith("merkle_tail_kernel")]
    n = sum(v["launches"] for v in hashing)
    b = sum(v["dram_read"] + v["dram_write"] for v in hashing)
    json.dump({"source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none over one "
                         f"step of bench.py --profile-mode ({src})",
               "hashing_launches": n, "hashing_dram_bytes_per_step": b, "hashing_dram_bytes_per_launch": b / max(n, 1),
               "per_kernel": per_kernel}, open(dst, "w"), indent=1)
    print("wrote", dst)


def pipe(srcs, dst):
    """ALU / FMA pipe utilisation per launch class from one or more --set full captures (comma-separated .ncu-rep files):
    the executed-instruction view printed next to bench.py's algorithmic roofline fractions.  Time-weighted per class."""
    import json
    CLASS = {"merkle_subtree_kernel<0>": "merkle_subtree_kernel<VALUES>", "merkle_subtree_kernel<1>": "merkle_subtree_kernel<FOLD>",
             "merkle_subtree_kernel<2>": "merkle_subtree_kernel<DIGESTS>"}
    acc = collections.OrderedDict()
    for src in srcs.split(","):
        raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr = rows[0]
        col = lambda name: hdr.index(name)
        for r in rows[2:]:
            k = short(r[col("Kernel Name")])
            k = CLASS.get(k, re.sub(r"<.*", "", k) if k.startswith(("merkle_tail", "fri_")) else k)
            t = float(r[col("gpu__time_duration.sum")].replace(",", ""))
            a = acc.setdefault(k, {"t": 0.0, "alu": 0.0, "fma": 0.0, "fmaheavy": 0.0, "lsu": 0.0, "issue": 0.0, "n": 0})
            a["t"] += t; a["n"] += 1
            a["alu"] += t * float(r[col("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active")])
            a["fma"] += t * float(r[col("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active")])
            a["issue"] += t * float(r[col("smsp__issue_active.avg.pct_of_peak_sustained_active")])
            # the integer multiplies (IMAD, half-rate IMAD.HI / IMAD.WIDE, and ptxas' VIADD / IMAD.IADD adds) all run on the
            # FMA-HEAVY half of the FMA pipe: the combined fma figure above averages it with the idle FMA-lite half
            a["fmaheavy"] += t * float(r[col("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed")])
            a["lsu"] += t * float(r[col("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed")])
    out = {"_source": f"ncu --set full --clock-control none: {srcs}; sm__pipe_alu_cycles_active / sm__pipe_fma_cycles_active / "
                      "smsp__issue_active .avg.pct_of_peak_sustained_active, time-weighted over the captured launches of each class"}
    for k, a in acc.items():
        out[k] = {"alu_pipe_pct": round(a["alu"] / a["t"], 1), "fma_pipe_pct": round(a["fma"] / a["t"], 1),
                  "fmaheavy_pipe_pct": round(a["fmaheavy"] / a["t"], 1), "lsu_wavefront_pct": round(a["lsu"] / a["t"], 1),
                  "issue_active_pct": round(a["issue"] / a["t"], 1), "launches_captured": a["n"]}
    json.dump(out, open(dst, "w"), indent=1)
    print("wrote", dst)


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic, "pipe": pipe}[sys.argv[1]](sys.argv[2], sys.argv[3])
