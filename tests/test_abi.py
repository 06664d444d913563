"""CPU tests of the boundary: the shared library loads, exports every symbol include/stark_b200.h
declares, refuses to run without a device (no CPU fallback), and the host-side Channel matches the
oracle.  No GPU compute is issued here."""
import os
import re
import subprocess

import pytest

P = 3221225473


def test_header_symbols_exported(sp):
    L = sp.lib()
    syms = sp.exported_symbols()
    assert len(syms) > 50
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/stark_b200.h but not exported"
    out = subprocess.check_output(["nm", "-D", "--defined-only", sp.LIB_PATH], text=True)
    exported = set(re.findall(r" T (stark\w*)", out))
    assert set(syms) <= exported
    # nothing but the declared API (plus C++ runtime) leaks with C linkage
    assert exported <= set(syms), exported - set(syms)


def test_header_compiles_as_c(sp, tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "stark_b200.h"\nint main(void){return STARK_OK;}\n')
    inc = os.path.dirname(sp.HEADER_PATH)
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.check_call([cc, "-std=c99", "-Wall", "-Werror", "-I", inc, "-c", str(src), "-o", str(tmp_path / "t.o")])


def test_sm100a_cubin_present(sp):
    out = subprocess.run(["cuobjdump", "-lelf", sp.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback(sp):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(sp.StarkError) as e:
        sp.Context(P, 5, 0)
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)


def test_ctx_rejects_bad_modulus(sp):
    for m in (0, 4, 15, (1 << 32) + 15, 0xFFFFFFFF00000001):
        with pytest.raises(sp.StarkError) as e:
            sp.Context(m, 0, 0)
        assert e.value.code == 3


def test_product_does_not_touch_oracle():
    """The product path must not import, link or call anything under oracle/."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "stark-prover_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "pyoracle" not in txt and "stark_oracle" not in txt and "or_fri" not in txt, f
    libs = subprocess.run(["ldd", os.path.join(pkg, "libstark_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in libs


def test_host_channel_matches_oracle(sp, orc, golden):
    c = golden["spec_anchors"]["channel"]
    ch = sp.Channel(P)
    ch.send(golden["spec_anchors"]["root_0_to_7"].encode())
    assert ch.state == c["after_send_root"]
    assert ch.receive_random_field_element() == c["field_element"]
    assert ch.receive_random_int(0, 8191, True) == c["random_int_0_8191"]
    assert ch.state == c["after_random_int"]
    assert ch.proof_size() == c["proof_size"] and ch.compressed_proof_size() == c["compressed_proof_size"]
    a, b = sp.Channel(P), orc.Channel(P)
    for step in range(60):
        if step % 3 == 0:
            m = bytes((step * 13 + i) & 255 for i in range(step % 70))
            a.send(m); b.send(m)
        elif step % 3 == 1:
            assert a.receive_random_field_element() == b.receive_random_field_element()
        else:
            assert a.receive_random_int(step, step * step + 9, step % 2 == 1) == b.receive_random_int(step, step * step + 9, step % 2 == 1)
        assert a.state == b.state
    assert a.proof == b.proof and a.proof_flat() == b.proof_flat()
    assert a.compressed_proof == b.compressed_proof


def test_host_channel_block_boundaries_and_reuse(sp):
    """Channel::send lays state || hex(message) || padding out in one buffer and hashes it in one call: every message
    length around the 64-byte block boundaries (with and without a previous state), long messages, and transcripts that
    reuse a recycled log arena, against hashlib (sha256 1.5.0 `digest` of the concatenated text, channel.rs:35-44)."""
    import hashlib
    lens = list(range(0, 70)) + [95, 96, 97, 127, 128, 129, 991, 992, 993, 1023, 1024, 1025, 5000, 70001]
    for rnd in range(3):                       # later rounds start on an arena returned by the previous Channel
        ch = sp.Channel(P)
        state = ""
        for n in lens:
            m = bytes((7 * n + 3 * i + rnd) & 255 for i in range(n))
            ch.send(m)
            state = hashlib.sha256((state + m.hex()).encode()).hexdigest()
            assert ch.state == state, (rnd, n)
        assert ch.proof == [bytes((7 * n + 3 * i + rnd) & 255 for i in range(n)) for n in lens]
        assert ch.proof_size() == sum(lens)
        del ch


def test_channel_receive_before_send_is_error(sp):
    ch = sp.Channel(P)
    with pytest.raises(sp.StarkError):
        ch.receive_random_int(0, 10)
