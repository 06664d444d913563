"""The C++ host-side mirror (include/stark101.hpp) over the C ABI: compiled with g++ against the shared
library; the host-only part runs anywhere, the device part on the GPU box against the oracle."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = 3221225473


def _build(sp) -> str:
    out = os.path.join(ROOT, "build", "test_stark101")
    src = os.path.join(ROOT, "tests", "cpp", "test_stark101.cpp")
    hdrs = [os.path.join(ROOT, "include", h) for h in ("stark101.hpp", "stark_b200.h")]
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(f) for f in [src, sp.LIB_PATH] + hdrs):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        libdir = os.path.dirname(sp.LIB_PATH)
        subprocess.check_call([cxx, "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-o", out,
                               "-L", libdir, "-l:libstark_b200.so", f"-Wl,-rpath,{libdir}"])
    return out


def test_cpp_mirror_host_side(sp):
    exe = _build(sp)
    r = subprocess.run([exe, "host"], capture_output=True, text=True)
    assert r.returncode == 0 and "host ok" in r.stdout, r.stderr


def test_cpp_mirror_panics_without_device(sp):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([_build(sp), "gpu", "8", "5", "43", "2"], capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("log_n,log_deg,seed,q", [(10, 7, 43, 3), (16, 13, 5, 4)])
def test_cpp_mirror_fri_matches_oracle(sp, orc, log_n, log_deg, seed, q):
    r = subprocess.run([_build(sp), "gpu", str(log_n), str(log_deg), str(seed), str(q)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = dict(line.split(" ", 1) for line in r.stdout.strip().splitlines() if not line.startswith("root "))
    roots = [line.split()[2] for line in r.stdout.splitlines() if line.startswith("root ")]
    c = orc.synthetic_poly_exact_degree(seed, 1 << log_deg, P)
    ch = orc.Channel(P)
    pr = orc.fri_commit_fast(c, log_n, 5, orc.root_of_unity(log_n), ch, P)
    assert got["state_after_commit"] == ch.state
    orc.decommit_fri(q, (1 << log_n) - 1, pr, ch)
    assert got["final_state"] == ch.state
    assert int(got["proof_size"]) == ch.proof_size() and int(got["compressed_proof_size"]) == ch.compressed_proof_size()
    assert roots == [pr.tree(k).root_hex() for k in range(pr.num_layers)]
    fin = pr.final_poly()
    assert got["final_poly_len"] == f"{len(fin)} value {int(fin[0]) if len(fin) else 0}"
