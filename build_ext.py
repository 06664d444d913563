"""Builds stark-prover_b200/libstark_b200.so (sm_100a) with nvcc, in-tree.  Used by __graft_entry__.build()."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "stark-prover_b200")
CSRC = os.path.join(PKG, "csrc")
OUT = os.path.join(PKG, "libstark_b200.so")
INFO = os.path.join(PKG, "libstark_b200.build.json")
BUILD = os.path.join(ROOT, "build")
SOURCES = ["api.cu", "merkle.cu", "ntt.cu", "fri.cu", "stark101.cu", "peaks.cu", "fourstep.cu", "verify.cu", "multi.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr"]


def _host_cxx() -> str | None:
    # the image's /opt/gcc wrapper lacks some spec files; prefer the distro compiler
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else None


def _boot_id() -> str:
    try:
        return open("/proc/sys/kernel/random/boot_id").read().strip()
    except Exception:
        return ""


def _deps(src: str) -> list[str]:
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".hpp", ".cuh", ".h"))]
    return [src, os.path.join(ROOT, "include", "stark_b200.h")] + hdrs


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        if os.path.exists(OUT) and not force:      # a box without the toolkit: use the library that travelled with the tree
            return OUT
        raise RuntimeError("nvcc not found and no prebuilt libstark_b200.so in the tree")
    os.makedirs(BUILD, exist_ok=True)
    ccbin = _host_cxx()
    extra = os.environ.get("STARK_NVCC_DEFS", "").split()     # e.g. "-DSTARK_SHA_ADDS_ON_FMA=0" for kernel experiments
    base = [nvcc] + NVCC_FLAGS + extra + (["-ccbin", ccbin] if ccbin else [])
    objs, jobs = [], []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        if not os.path.exists(src):
            continue
        obj = os.path.join(BUILD, s.replace(".cu", ".o"))
        objs.append(obj)
        newest = max(os.path.getmtime(d) for d in _deps(src))
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < newest:
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        r = subprocess.run(base + ["-c", src, "-o", obj], capture_output=True, text=True)
        log = os.path.join(BUILD, os.path.basename(obj) + ".log")
        with open(log, "w") as fh:
            fh.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
        list(ex.map(compile_one, jobs))
    if jobs or force or not os.path.exists(OUT):
        r = subprocess.run(base + ["-shared", "-o", OUT] + objs + ["-lcudart", "-ldl"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        # provenance of the binary (bench.py reports whether the library it ran was compiled on the box or travelled with the tree)
        import json, socket, time
        ver = subprocess.run([nvcc, "--version"], capture_output=True, text=True).stdout.strip().splitlines()
        with open(INFO, "w") as fh:
            json.dump({"host": socket.gethostname(), "boot_id": _boot_id(), "time": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()), "nvcc": ver[-2] if len(ver) >= 2 else "",
                       "flags": NVCC_FLAGS + extra, "recompiled": [os.path.basename(j[0]) for j in jobs]}, fh)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
