"""ctypes loader for the CPU oracle (oracle/stark_oracle.c) plus a tiny hashlib twin.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this module.

The hashlib twin (`py_*` functions) restates the Merkle and channel rules a second time with
Python's own SHA-256, so that the C oracle's hashing is cross-checked by an independent
implementation (implication iii).
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libstark_oracle.so")

P_DEFAULT = 3221225473  # 3*2^30 + 1
G_DEFAULT = 5


def build(force: bool = False) -> str:
    """Compile oracle/stark_oracle.c -> oracle/libstark_oracle.so (gcc, OpenMP)."""
    src = os.path.join(_HERE, "stark_oracle.c")
    hdr = os.path.join(_HERE, "stark_oracle.h")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "libstark_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None

u64 = C.c_uint64
u64p = C.POINTER(C.c_uint64)
u8p = C.POINTER(C.c_uint8)
szt = C.c_size_t


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build()
    L = C.CDLL(_LIB_PATH)

    def sig(name, res, *args):
        f = getattr(L, name)
        f.restype = res
        f.argtypes = list(args)

    for n in ("or_fe_add", "or_fe_sub", "or_fe_mul", "or_fe_pow", "or_fe_div"):
        sig(n, u64, u64, u64, u64)
    for n in ("or_fe_new", "or_fe_neg", "or_fe_inverse"):
        sig(n, u64, u64, u64)
    sig("or_fe_from_i128", u64, C.c_int64, u64)
    sig("or_poly_trim", szt, u64p, szt)
    sig("or_poly_evaluate", u64, u64p, szt, u64, u64)
    for n in ("or_poly_add", "or_poly_sub", "or_poly_mul"):
        sig(n, szt, u64p, szt, u64p, szt, u64p, u64)
    sig("or_poly_div_rem", C.c_int, u64p, szt, u64p, szt, u64p, C.POINTER(szt), u64p, C.POINTER(szt), u64)
    sig("or_poly_from_roots", szt, u64p, szt, u64p, u64)
    sig("or_poly_interpolate", szt, u64p, u64p, szt, u64p, u64)
    sig("or_lagrange_basis", C.c_int, u64p, szt, u64p, u64)
    sig("or_sha256", None, C.c_char_p, szt, u8p)
    sig("or_sha256_set_accel", None, C.c_int)
    sig("or_sha256_accel_active", C.c_int)
    sig("or_merkle_new", C.c_void_p, u64p, szt)
    sig("or_merkle_free", None, C.c_void_p)
    sig("or_merkle_num_leaves", szt, C.c_void_p)
    sig("or_merkle_depth", szt, C.c_void_p)
    sig("or_merkle_root", None, C.c_void_p, u8p)
    sig("or_merkle_root_hex", None, C.c_void_p, C.c_char_p)
    sig("or_merkle_node", C.c_int, C.c_void_p, szt, szt, u8p)
    sig("or_merkle_path", szt, C.c_void_p, szt, u8p)
    sig("or_merkle_verify", C.c_int, u8p, szt, szt, u64, u8p, szt)
    sig("or_merkle_root_only", None, u64p, szt, u8p)
    sig("or_merkle_root_from_digests", None, u8p, szt, u8p)
    sig("or_channel_new", C.c_void_p, u64)
    sig("or_channel_free", None, C.c_void_p)
    sig("or_channel_send", None, C.c_void_p, C.c_char_p, szt)
    sig("or_channel_receive_random_field_element", u64, C.c_void_p)
    sig("or_channel_receive_random_int", u64, C.c_void_p, u64, u64, C.c_int)
    sig("or_channel_proof_size", szt, C.c_void_p)
    sig("or_channel_compressed_proof_size", szt, C.c_void_p)
    sig("or_channel_state", C.c_char_p, C.c_void_p)
    sig("or_channel_proof_len", szt, C.c_void_p)
    sig("or_channel_proof_msg", szt, C.c_void_p, szt, C.POINTER(u8p))
    sig("or_channel_compressed_len", szt, C.c_void_p)
    sig("or_channel_compressed_msg", szt, C.c_void_p, szt, C.POINTER(u8p))
    sig("or_channel_proof_flat", szt, C.c_void_p, u8p)
    sig("or_coset_domain", None, u64, u64, szt, u64p, u64)
    sig("or_next_fri_domain", None, u64p, szt, u64p, u64)
    sig("or_next_fri_polynomial", szt, u64p, szt, u64, u64p, u64)
    sig("or_root_of_unity", u64, u64, C.c_uint, u64)
    sig("or_ntt", C.c_int, u64p, C.c_uint, u64, u64)
    sig("or_intt", C.c_int, u64p, C.c_uint, u64, u64)
    sig("or_coset_evaluate", C.c_int, u64p, szt, C.c_uint, u64, u64, u64p, u64)
    sig("or_coset_interpolate", C.c_int, u64p, C.c_uint, u64, u64, u64p, u64)
    sig("or_batch_inverse", None, u64p, szt, u64)
    sig("or_fri_fold_evals", None, u64p, szt, u64, u64, u64, u64p, u64)
    sig("or_fri_commit_literal", C.c_void_p, u64p, szt, u64p, szt, C.c_void_p, u64)
    sig("or_fri_commit_fast", C.c_void_p, u64p, szt, C.c_uint, u64, u64, C.c_void_p, u64)
    sig("or_fri_commit_fast_rootonly", C.c_int, u64p, szt, C.c_uint, u64, u64, C.c_void_p, u64)
    sig("or_fri_free", None, C.c_void_p)
    sig("or_fri_num_layers", szt, C.c_void_p)
    sig("or_fri_layer_len", szt, C.c_void_p, szt)
    sig("or_fri_layer", u64p, C.c_void_p, szt)
    sig("or_fri_tree", C.c_void_p, C.c_void_p, szt)
    sig("or_fri_final_poly", szt, C.c_void_p, u64p)
    sig("or_decommit_fri_layers", None, szt, C.c_void_p, C.c_void_p)
    sig("or_decommit_fri", None, szt, szt, C.c_void_p, C.c_void_p)
    sig("or_fibsq_trace", None, u64, szt, u64p, u64)
    sig("or_stark101_prove", C.c_int, u64, C.c_uint, C.c_uint, u64, szt, C.c_int, C.c_void_p, u64)
    sig("or_num_threads", C.c_int)
    sig("or_set_num_threads", None, C.c_int)
    _lib = L
    return L


def _a(x) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(x, dtype=np.uint64))


def _p(a: np.ndarray):
    return a.ctypes.data_as(u64p)


def _b(a: np.ndarray):
    return a.ctypes.data_as(u8p)


# ---------------------------------------------------------------- field
def fe_add(a, b, M): return lib().or_fe_add(a, b, M)
def fe_sub(a, b, M): return lib().or_fe_sub(a, b, M)
def fe_mul(a, b, M): return lib().or_fe_mul(a, b, M)
def fe_pow(a, e, M): return lib().or_fe_pow(a, e, M)
def fe_div(a, b, M): return lib().or_fe_div(a, b, M)
def fe_new(a, M): return lib().or_fe_new(a, M)
def fe_neg(a, M): return lib().or_fe_neg(a, M)
def fe_inverse(a, M): return lib().or_fe_inverse(a, M)
def fe_from_int(v, M): return lib().or_fe_from_i128(v, M)


# ---------------------------------------------------------------- polynomial (literal tier)
def poly_trim(c):
    c = _a(c)
    return c[: lib().or_poly_trim(_p(c), len(c))].copy()


def poly_evaluate(c, x, M):
    c = _a(c)
    return lib().or_poly_evaluate(_p(c), len(c), x, M)


def _binop(name, a, b, M, outlen):
    a, b = _a(a), _a(b)
    out = np.zeros(max(outlen, 1), dtype=np.uint64)
    n = getattr(lib(), name)(_p(a), len(a), _p(b), len(b), _p(out), M)
    return out[:n].copy()


def poly_add(a, b, M): return _binop("or_poly_add", a, b, M, max(len(a), len(b)))
def poly_sub(a, b, M): return _binop("or_poly_sub", a, b, M, max(len(a), len(b)))
def poly_mul(a, b, M): return _binop("or_poly_mul", a, b, M, len(a) + len(b))


def poly_div_rem(a, b, M):
    a, b = _a(a), _a(b)
    q = np.zeros(max(len(a), 1) + 1, dtype=np.uint64)
    r = np.zeros(max(len(a), 1) + 1, dtype=np.uint64)
    ql, rl = szt(0), szt(0)
    rc = lib().or_poly_div_rem(_p(a), len(a), _p(b), len(b), _p(q), C.byref(ql), _p(r), C.byref(rl), M)
    if rc != 0:
        raise ZeroDivisionError("Division by zero polynomial")
    return q[: ql.value].copy(), r[: rl.value].copy()


def poly_from_roots(roots, M):
    roots = _a(roots)
    out = np.zeros(len(roots) + 1, dtype=np.uint64)
    n = lib().or_poly_from_roots(_p(roots), len(roots), _p(out), M)
    return out[:n].copy()


def poly_interpolate(xs, ys, M):
    xs, ys = _a(xs), _a(ys)
    if len(xs) != len(ys):
        raise ValueError("Mismatched x and y lengths")
    out = np.zeros(max(len(xs), 1), dtype=np.uint64)
    n = lib().or_poly_interpolate(_p(xs), _p(ys), len(xs), _p(out), M)
    if n == (1 << 64) - 1:
        raise ArithmeticError("Z(x) should be divisible by (x - x_i)")
    return out[:n].copy()


def lagrange_basis(xs, M):
    xs = _a(xs)
    n = len(xs)
    out = np.zeros((n, n), dtype=np.uint64)
    if lib().or_lagrange_basis(_p(xs), n, _p(out), M) != 0:
        raise ArithmeticError("Z(x) should be divisible by (x - x_i)")
    return out


# ---------------------------------------------------------------- sha / merkle
def sha256(msg: bytes) -> bytes:
    out = np.zeros(32, dtype=np.uint8)
    lib().or_sha256(msg, len(msg), _b(out))
    return out.tobytes()


class Tree:
    def __init__(self, leaves=None, handle=None, owned=True):
        self._owned = owned
        if handle is not None:
            self.h = handle
        else:
            leaves = _a(leaves)
            self.h = lib().or_merkle_new(_p(leaves), len(leaves))
            if not self.h:
                raise ValueError("called `Option::unwrap()` on a `None` value (empty tree)")

    def __del__(self):
        if getattr(self, "_owned", False) and getattr(self, "h", None):
            lib().or_merkle_free(self.h)
            self.h = None

    @property
    def n(self): return lib().or_merkle_num_leaves(self.h)
    @property
    def depth(self): return lib().or_merkle_depth(self.h)

    def root(self) -> bytes:
        out = np.zeros(32, dtype=np.uint8)
        lib().or_merkle_root(self.h, _b(out))
        return out.tobytes()

    def root_hex(self) -> str:
        buf = C.create_string_buffer(65)
        lib().or_merkle_root_hex(self.h, buf)
        return buf.value.decode()

    def node(self, level, j) -> bytes:
        out = np.zeros(32, dtype=np.uint8)
        if lib().or_merkle_node(self.h, level, j, _b(out)) != 0:
            raise IndexError((level, j))
        return out.tobytes()

    def path(self, idx) -> bytes:
        out = np.zeros(32 * (self.depth + 1), dtype=np.uint8)
        n = lib().or_merkle_path(self.h, idx, _b(out))
        return out[:n].tobytes()


def merkle_verify(root: bytes, n_leaves, idx, value, path: bytes) -> bool:
    r = np.frombuffer(root, dtype=np.uint8).copy()
    p = np.frombuffer(path, dtype=np.uint8).copy() if path else np.zeros(1, dtype=np.uint8)
    return bool(lib().or_merkle_verify(_b(r), n_leaves, idx, value, _b(p), len(path)))


def merkle_root_from_digests(digests: list) -> bytes:
    d = np.frombuffer(b"".join(digests), dtype=np.uint8).copy()
    out = np.zeros(32, dtype=np.uint8)
    lib().or_merkle_root_from_digests(_b(d), len(digests), _b(out))
    return out.tobytes()


def merkle_root_only(leaves) -> bytes:
    leaves = _a(leaves)
    out = np.zeros(32, dtype=np.uint8)
    lib().or_merkle_root_only(_p(leaves), len(leaves), _b(out))
    return out.tobytes()


# ---------------------------------------------------------------- channel
class Channel:
    def __init__(self, M=P_DEFAULT):
        self.M = M
        self.h = lib().or_channel_new(M)

    def __del__(self):
        if getattr(self, "h", None):
            lib().or_channel_free(self.h)
            self.h = None

    def send(self, msg: bytes): lib().or_channel_send(self.h, msg, len(msg))
    def receive_random_field_element(self): return lib().or_channel_receive_random_field_element(self.h)
    def receive_random_int(self, lo, hi, show=False): return lib().or_channel_receive_random_int(self.h, lo, hi, int(show))
    def proof_size(self): return lib().or_channel_proof_size(self.h)
    def compressed_proof_size(self): return lib().or_channel_compressed_proof_size(self.h)
    @property
    def state(self): return lib().or_channel_state(self.h).decode()

    def _msgs(self, n, getter):
        out = []
        for i in range(n):
            p = u8p()
            ln = getter(self.h, i, C.byref(p))
            out.append(bytes(C.cast(p, C.POINTER(C.c_uint8 * ln)).contents) if ln else b"")
        return out

    @property
    def proof(self): return self._msgs(lib().or_channel_proof_len(self.h), lib().or_channel_proof_msg)
    @property
    def compressed_proof(self): return self._msgs(lib().or_channel_compressed_len(self.h), lib().or_channel_compressed_msg)

    def proof_flat(self) -> bytes:
        n = lib().or_channel_proof_flat(self.h, None)
        out = np.zeros(max(n, 1), dtype=np.uint8)
        lib().or_channel_proof_flat(self.h, _b(out))
        return out[:n].tobytes()


# ---------------------------------------------------------------- domains / fast tier
def root_of_unity(log_n, M=P_DEFAULT, g=G_DEFAULT): return lib().or_root_of_unity(g, log_n, M)


def coset_domain(offset, omega, n, M):
    out = np.zeros(n, dtype=np.uint64)
    lib().or_coset_domain(offset, omega, n, _p(out), M)
    return out


def next_fri_domain(d, M):
    d = _a(d)
    out = np.zeros(len(d) // 2, dtype=np.uint64)
    lib().or_next_fri_domain(_p(d), len(d), _p(out), M)
    return out


def next_fri_polynomial(c, beta, M):
    c = _a(c)
    out = np.zeros((len(c) + 1) // 2 + 1, dtype=np.uint64)
    n = lib().or_next_fri_polynomial(_p(c), len(c), beta, _p(out), M)
    return out[:n].copy()


def ntt(a, log_n, omega, M):
    a = _a(a).copy()
    assert lib().or_ntt(_p(a), log_n, omega, M) == 0
    return a


def intt(a, log_n, omega, M):
    a = _a(a).copy()
    assert lib().or_intt(_p(a), log_n, omega, M) == 0
    return a


def coset_evaluate(c, log_n, offset, omega, M):
    c = _a(c)
    out = np.zeros(1 << log_n, dtype=np.uint64)
    assert lib().or_coset_evaluate(_p(c), len(c), log_n, offset, omega, _p(out), M) == 0
    return out


def coset_interpolate(e, log_n, offset, omega, M):
    e = _a(e)
    out = np.zeros(1 << log_n, dtype=np.uint64)
    assert lib().or_coset_interpolate(_p(e), log_n, offset, omega, _p(out), M) == 0
    return out


def batch_inverse(a, M):
    a = _a(a).copy()
    lib().or_batch_inverse(_p(a), len(a), M)
    return a


def fri_fold_evals(e, beta, offset, omega, M):
    e = _a(e)
    out = np.zeros(len(e) // 2, dtype=np.uint64)
    lib().or_fri_fold_evals(_p(e), len(e), beta, offset, omega, _p(out), M)
    return out


# ---------------------------------------------------------------- FRI
class FriProof:
    def __init__(self, h):
        if not h:
            raise ValueError("fri_commit failed (empty layer / unsupported field)")
        self.h = h

    def __del__(self):
        if getattr(self, "h", None):
            lib().or_fri_free(self.h)
            self.h = None

    @property
    def num_layers(self): return lib().or_fri_num_layers(self.h)

    def layer(self, k) -> np.ndarray:
        n = lib().or_fri_layer_len(self.h, k)
        p = lib().or_fri_layer(self.h, k)
        return np.ctypeslib.as_array(p, shape=(n,)).copy()

    def tree(self, k) -> Tree:
        t = Tree(handle=lib().or_fri_tree(self.h, k), owned=False)
        t._keep = self
        return t

    def final_poly(self):
        out = np.zeros(1, dtype=np.uint64)
        n = lib().or_fri_final_poly(self.h, _p(out))
        return out[:n].copy()


def fri_commit_literal(coeffs, domain, ch: Channel, M=P_DEFAULT) -> FriProof:
    c, d = _a(coeffs), _a(domain)
    return FriProof(lib().or_fri_commit_literal(_p(c), len(c), _p(d), len(d), ch.h, M))


def fri_commit_fast(coeffs, log_n, offset, omega, ch: Channel, M=P_DEFAULT) -> FriProof:
    c = _a(coeffs)
    return FriProof(lib().or_fri_commit_fast(_p(c), len(c), log_n, offset, omega, ch.h, M))


def fri_commit_fast_rootonly(coeffs, log_n, offset, omega, ch: Channel, M=P_DEFAULT) -> None:
    c = _a(coeffs)
    assert lib().or_fri_commit_fast_rootonly(_p(c), len(c), log_n, offset, omega, ch.h, M) == 0


def decommit_fri_layers(index, proof: FriProof, ch: Channel): lib().or_decommit_fri_layers(index, proof.h, ch.h)
def decommit_fri(num_queries, max_index, proof: FriProof, ch: Channel): lib().or_decommit_fri(num_queries, max_index, proof.h, ch.h)


def fibsq_trace(a1, rows, M=P_DEFAULT):
    out = np.zeros(rows, dtype=np.uint64)
    lib().or_fibsq_trace(a1, rows, _p(out), M)
    return out


def stark101_prove(ch: Channel, a1=3141592, log_trace=10, log_blowup=3, g=G_DEFAULT, num_queries=3,
                   literal=False, M=P_DEFAULT) -> None:
    rc = lib().or_stark101_prove(a1, log_trace, log_blowup, g, num_queries, int(literal), ch.h, M)
    if rc != 0:
        raise RuntimeError(f"or_stark101_prove failed rc={rc}")


def num_threads(): return lib().or_num_threads()
def set_num_threads(n): lib().or_set_num_threads(n)


# ---------------------------------------------------------------- synthetic inputs
def splitmix64(seed: int, n: int) -> np.ndarray:
    """n outputs of splitmix64 seeded with `seed` (vectorised; identical in C/CUDA/Python)."""
    with np.errstate(over="ignore"):
        k = np.arange(1, n + 1, dtype=np.uint64)
        z = np.uint64(seed) + k * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def synthetic_column(seed: int, n: int, M=P_DEFAULT) -> np.ndarray:
    return splitmix64(seed, n) % np.uint64(M)


def synthetic_poly_exact_degree(seed: int, n_coeffs: int, M=P_DEFAULT) -> np.ndarray:
    c = synthetic_column(seed, n_coeffs, M)
    if c[-1] == 0:
        c[-1] = 1
    return c


# ---------------------------------------------------------------- hashlib twin (independent cross-check)
def py_leaf(v: int) -> bytes:
    return hashlib.sha256(int(v).to_bytes(8, "big")).digest()          # merkle/mod.rs:13-16


def py_merkle_levels(values) -> list[list[bytes]]:
    lv = [[py_leaf(v) for v in values]]
    while len(lv[-1]) > 1:
        cur = lv[-1]
        nxt = [hashlib.sha256(cur[i] + cur[i + 1]).digest() for i in range(0, len(cur) - 1, 2)]
        if len(cur) & 1:
            nxt.append(cur[-1])                                          # lone node promoted (rs_merkle)
        lv.append(nxt)
    return lv


def py_merkle_root_hex(values) -> str:
    return py_merkle_levels(values)[-1][0].hex()


def py_merkle_path(values, idx) -> bytes:
    out = b""
    lv = py_merkle_levels(values)
    for level in lv[:-1]:
        sib = idx ^ 1
        if sib < len(level):
            out += level[sib]
        idx >>= 1
    return out


class PyChannel:
    """channel.rs restated with hashlib (state is a lowercase hex string)."""

    def __init__(self, M=P_DEFAULT):
        self.M, self.state, self.proof, self.compressed_proof = M, "", [], []

    def send(self, m: bytes):
        self.state = hashlib.sha256((self.state + m.hex()).encode()).hexdigest()    # :35-44
        self.proof.append(bytes(m))
        self.compressed_proof.append(bytes(m))

    def receive_random_int(self, lo, hi, show=False):
        num = (int(self.state, 16) + lo) % (hi - lo + 1)                             # :72
        self.state = hashlib.sha256(self.state.encode()).hexdigest()                 # :75-76
        if show:
            self.proof.append(num.to_bytes(8, "big"))
        return num

    def receive_random_field_element(self):
        num = self.receive_random_int(0, self.M - 1, False)
        self.proof.append(num.to_bytes(8, "big"))
        return num % self.M

    def proof_size(self): return sum(len(m) for m in self.proof)
    def compressed_proof_size(self): return sum(len(m) for m in self.compressed_proof)
