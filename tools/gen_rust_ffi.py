#!/usr/bin/env python
"""Generates the `extern "C"` block of shim/src/ffi.rs from include/stark_b200.h, prototype by prototype.
    python tools/gen_rust_ffi.py            prints the block
    python tools/gen_rust_ffi.py --write    rewrites the block between the BEGIN/END markers in shim/src/ffi.rs
tests/test_shim.py runs the same parser and fails on any drift between the header and the committed block."""
from __future__ import annotations

import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "stark_b200.h")
FFI = os.path.join(ROOT, "shim", "src", "ffi.rs")
BEGIN, END = "    // BEGIN GENERATED (tools/gen_rust_ffi.py)", "    // END GENERATED"

OPAQUE = ["stark_ctx", "stark_vec", "stark_tree", "stark_fri", "stark_channel", "stark_mg", "stark_mg_fri"]
SCALARS = {"int": "c_int", "unsigned": "c_uint", "unsigned int": "c_uint", "size_t": "usize", "uint64_t": "u64", "uint32_t": "u32",
           "uint8_t": "u8", "char": "c_char", "long long": "i64", "unsigned long long": "u64", "double": "f64", "void": "c_void"}


def prototypes(text: str) -> list[tuple[str, str, list[tuple[str, str]]]]:
    """[(return type, name, [(C type, parameter name)])] for every function declared in the header."""
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    text = "\n".join(l for l in text.splitlines() if not l.lstrip().startswith("#"))
    text = text.replace('extern "C" {', " ").replace("}", " ")
    out = []
    for stmt in text.split(";"):
        stmt = " ".join(stmt.split())
        m = re.match(r"^(.*?)\b(stark\w*)\s*\((.*)\)$", stmt)
        if not m or stmt.startswith("typedef"):
            continue
        ret, name, params = m.group(1).strip(), m.group(2), m.group(3).strip()
        plist = []
        if params and params != "void":
            for p in params.split(","):
                p = p.strip()
                arr = re.search(r"\[\s*\d*\s*\]$", p)
                if arr:
                    p = p[: arr.start()].strip()
                pm = re.match(r"^(.*?)(\w+)$", p)
                ctype, pname = pm.group(1).strip(), pm.group(2)
                if arr:
                    ctype += "*"
                plist.append((ctype, pname))
        out.append((ret, name, plist))
    return out


def rust_type(c: str) -> str:
    """`const uint64_t* const*` -> `*const *const u64`, `stark_ctx**` -> `*mut *mut stark_ctx`, ..."""
    c = c.replace("*", " * ").split()
    # walk from the right: every `*` (optionally followed by const) is a pointer level
    levels = []
    while c and (c[-1] == "*" or (c[-1] == "const" and len(c) >= 2 and c[-2] == "*")):
        const_ptr = c[-1] == "const"
        if const_ptr:
            c.pop()
        c.pop()
        levels.append(const_ptr)
    base_const = "const" in c
    base = " ".join(t for t in c if t != "const")
    rt = SCALARS.get(base, base if base in OPAQUE else None)
    if rt is None:
        raise ValueError(f"unmapped C type {base!r}")
    # levels[0] is the outermost pointer; the pointee of the innermost is `base`
    res = rt
    for depth, _ in enumerate(reversed(levels)):
        pointee_const = base_const if depth == 0 else list(reversed(levels))[depth - 1]
        res = ("*const " if pointee_const else "*mut ") + res
    return res


RESERVED = {"type", "ref", "in", "box", "move", "fn", "mod", "use", "self", "super", "where", "loop", "match", "impl", "trait"}


def rust_decl(ret: str, name: str, params: list[tuple[str, str]]) -> str:
    args = ", ".join(f"{(p + '_') if p in RESERVED else p}: {rust_type(t)}" for t, p in params)
    r = "" if ret == "void" else f" -> {rust_type(ret)}"
    return f"    pub fn {name}({args}){r};"


def block() -> str:
    protos = prototypes(open(HEADER).read())
    return "\n".join([BEGIN] + [rust_decl(*p) for p in protos] + [END])


if __name__ == "__main__":
    b = block()
    if "--write" in sys.argv:
        s = open(FFI).read()
        i, j = s.index(BEGIN), s.index(END) + len(END)
        open(FFI, "w").write(s[:i] + b + s[j:])
        print(f"rewrote the generated block of {FFI} ({b.count('pub fn')} functions)")
    else:
        print(b)
