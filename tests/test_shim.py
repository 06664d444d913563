"""Mechanical checks of the Rust shim (shim/), which this image cannot compile (no rustc / cargo):
  * the `extern "C"` block of shim/src/ffi.rs is exactly what tools/gen_rust_ffi.py derives from include/stark_b200.h
    -- any arity / type / name drift between the header and the Rust declarations fails here;
  * an independent reading of the committed block agrees with the header on names and argument counts, and with the
    symbols the built library exports;
  * every `ffi::stark_*` call in the shim names a declared function and passes the declared number of arguments;
  * no body is a placeholder."""
import importlib.util
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "shim")


def _gen():
    spec = importlib.util.spec_from_file_location("gen_rust_ffi", os.path.join(ROOT, "tools", "gen_rust_ffi.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _rust_files():
    out = []
    for d, _, fs in os.walk(SHIM):
        out += [os.path.join(d, f) for f in fs if f.endswith(".rs")]
    return sorted(out)


def _split_args(s: str) -> list[str]:
    """top-level comma split of an argument list (parentheses, brackets, braces and angle brackets nest)"""
    args, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{<":
            depth += 1
        elif ch in ")]}>":
            depth -= 1
        if ch == "," and depth == 0:
            args.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        args.append(cur.strip())
    return args


def _declared():
    """{name: n_args} from the committed extern block, parsed independently of the generator"""
    txt = open(os.path.join(SHIM, "src", "ffi.rs")).read()
    block = txt[txt.index('extern "C" {'):]
    block = block[: block.index("\n}\n")]
    decl = {}
    for m in re.finditer(r"pub fn (\w+)\((.*?)\)(?: -> [^;]+)?;", block):
        decl[m.group(1)] = len(_split_args(m.group(2)))
    return decl


def test_extern_block_matches_the_header():
    gen = _gen()
    txt = open(gen.FFI).read()
    i, j = txt.index(gen.BEGIN), txt.index(gen.END) + len(gen.END)
    assert txt[i:j] == gen.block(), "shim/src/ffi.rs is stale: run `python tools/gen_rust_ffi.py --write`"


def test_declarations_agree_with_header_and_library(sp):
    gen = _gen()
    protos = {name: len(params) for _, name, params in gen.prototypes(open(gen.HEADER).read())}
    decl = _declared()
    assert decl == protos
    assert sorted(decl) == sp.exported_symbols()
    L = sp.lib()
    assert not [s for s in decl if not hasattr(L, s)]


def test_type_mapping_spot_checks():
    gen = _gen()
    rt = gen.rust_type
    assert rt("const uint64_t*") == "*const u64" and rt("uint64_t*") == "*mut u64"
    assert rt("stark_ctx**") == "*mut *mut stark_ctx" and rt("const stark_tree*") == "*const stark_tree"
    assert rt("void* const*") == "*const *mut c_void" and rt("const uint8_t**") == "*mut *const u8"
    assert rt("const uint64_t* const*") == "*const *const u64" and rt("char*") == "*mut c_char"
    assert rt("unsigned") == "c_uint" and rt("size_t") == "usize" and rt("long long*") == "*mut i64"


def test_call_sites_exist_and_have_the_declared_arity():
    decl = _declared()
    calls = 0
    for path in _rust_files():
        src = open(path).read()
        src = re.sub(r"//[^\n]*", "", src)
        prefix = r"(?:ffi::)?" if path.endswith("ffi.rs") else r"ffi::"
        for m in re.finditer(prefix + r"(stark\w*)\s*\(", src):
            name = m.group(1)
            if path.endswith("ffi.rs") and src[max(0, m.start() - 7):m.start()].endswith("pub fn "):
                continue                                   # a declaration, not a call
            depth, k = 1, m.end()
            while depth:
                depth += {"(": 1, ")": -1}.get(src[k], 0)
                k += 1
            args = _split_args(src[m.end():k - 1])
            assert name in decl, f"{os.path.relpath(path, ROOT)}: {name} is not declared in ffi.rs"
            assert len(args) == decl[name], f"{os.path.relpath(path, ROOT)}: {name} called with {len(args)} arguments, declared with {decl[name]}"
            calls += 1
    assert calls >= 40


def test_no_placeholder_bodies_and_all_files_present():
    for rel in ("build.rs", "Cargo.toml.patch", "patches/element.rs.diff", "patches/lib.rs.diff", "src/ffi.rs", "src/merkle/mod.rs",
                "src/polynomial/gpu.rs", "src/fri/mod.rs", "src/fri/coset_fri.rs", "src/fri/fri_commit.rs", "src/fri/fri_verify.rs"):
        assert os.path.exists(os.path.join(SHIM, rel)), rel
    for path in _rust_files():
        src = open(path).read()
        assert "unimplemented!" not in src and "todo!" not in src, path
        code = re.sub(r"//[^\n]*", "", src)              # comments may hold half-open intervals
        code = re.sub(r'"(?:[^"\\]|\\.)*"', '""', code)
        assert code.count("{") == code.count("}") and code.count("(") == code.count(")") and code.count("[") == code.count("]"), \
            f"unbalanced delimiters in {path}"
    build = open(os.path.join(SHIM, "build.rs")).read()
    import build_ext
    for s in build_ext.SOURCES:
        assert f'"{s}"' in build, f"build.rs does not compile {s}"
    assert "compute_100a" in build


def test_reference_signatures_are_kept():
    """the pub fns the north-star names as the drop-in surface, with the reference's argument lists"""
    fri = open(os.path.join(SHIM, "src", "fri", "fri_commit.rs")).read()
    assert re.search(r"pub fn fri_commit<const M: u64>\(poly: Polynomial<M>, domain: &CosetFri<M>, channel: &mut Channel<M>\) -> FRIProof<M>", fri)
    assert re.search(r"pub fn decommit_fri_layers<const M: u64>\(index: usize, fri_layers: &\[Vec<FieldElement<M>>\], fri_merkles: &\[MerkleTree<M>\],\s*channel: &mut Channel<M>\)", fri)
    assert re.search(r"pub fn decommit_fri<const M: u64>\(num_queries: usize, max_index: usize, fri_layers: &\[Vec<FieldElement<M>>\], fri_merkles: &\[MerkleTree<M>\],\s*channel: &mut Channel<M>\)", fri)
    mk = open(os.path.join(SHIM, "src", "merkle", "mod.rs")).read()
    assert "pub fn new(data: Vec<FieldElement<MODULUS>>) -> Self" in mk and "pub fn root(&self) -> String" in mk
    assert "pub fn get_authentication_path(&self, idx: usize) -> Vec<u8>" in mk
    cf = open(os.path.join(SHIM, "src", "fri", "coset_fri.rs")).read()
    assert "pub fn generate_coset_domain(&self) -> Vec<FieldElement<M>>" in cf
    assert "pub fn new(offset: FieldElement<M>, omega: FieldElement<M>, domain_size: usize) -> Self" in cf
