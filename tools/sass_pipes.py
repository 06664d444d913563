#!/usr/bin/env python3
"""Static pipe-load estimate of one kernel from its SASS (cuobjdump -sass): issue cycles each pipe of an SM
sub-partition spends per warp if every instruction of the kernel ran once.  ALU and FMA-heavy pipes are 16 lanes
wide (2 cycles per warp instruction); IMAD.HI / IMAD.WIDE run at half that rate (profiles/r01_pipe_mix.txt).
Usage: tools/sass_pipes.py build/ntt.o <substring of the mangled kernel name> [...]"""
import collections
import re
import subprocess
import sys

FMA_HALF = ("IMAD.HI", "IMAD.WIDE")
ALU = ("IADD3", "LOP3", "SHF", "LEA", "MOV", "SEL", "ISETP", "PRMT", "VIMNMX", "IABS", "PLOP3", "CS2R", "BMSK", "SGXT", "FLO", "POPC", "BREV", "IMNMX", "VABSDIFF")
LSU = ("LDS", "STS", "LDG", "STG", "LD", "ST", "LDL", "STL", "ATOM", "RED", "LDC")


def kernels(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    name, body = None, collections.defaultdict(list)
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m and name:
            ins = m.group(1).split()
            if ins[0].startswith("@"):
                ins = ins[1:]
            body[name].append(ins[0])
    return body


def main():
    obj, pats = sys.argv[1], sys.argv[2:]
    for name, ops in kernels(obj).items():
        if pats and not any(p in name for p in pats):
            continue
        fma = alu = lsu = other = 0
        hist = collections.Counter()
        for op in ops:
            base = op.split(".")[0]
            hist[op if base == "IMAD" else base] += 1
            if base in ("IMAD", "VIADD", "IDP", "FFMA", "FMUL", "FADD", "HFMA2"):
                fma += 4 if op.startswith(FMA_HALF) else 2
            elif base in ALU:
                alu += 2
            elif base in LSU:
                lsu += 1
            else:
                other += 1
        print(f"{name}\n  instr {len(ops)}  fma-heavy cycles {fma}  alu cycles {alu}  lsu instr {lsu}  other {other}")
        print("  " + "  ".join(f"{k}:{v}" for k, v in hist.most_common(16)))


if __name__ == "__main__":
    main()
