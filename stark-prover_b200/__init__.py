"""stark-prover_b200 — ctypes binding of libstark_b200.so (include/stark_b200.h).

The product is the CUDA library; this module is the Python host-side mirror used by the tests and the
benchmark.  Names follow the reference crate (RazorClient/Stark-prover): `MerkleTree.new/root`
(src/merkle/mod.rs:10-26), `Channel.send/receive_random_*` (src/channel/channel.rs:35-84),
`CosetFri.generate_coset_domain` (src/fri/coset_fri.rs:32-36), `fri_commit`, `decommit_fri_layers`,
`decommit_fri` (src/fri/fri_commit.rs:72-179), `Polynomial.evaluate/interpolate` over a coset
(src/polynomial/ops.rs:76-83, :239-241).

There is no CPU fallback: importing works anywhere (so the symbol table can be checked), but every
compute entry point needs a CUDA device and raises `StarkError` otherwise.  The package directory name
contains a hyphen; import it with `importlib.import_module("stark-prover_b200")` or through the
`stark_prover_b200` alias module at the repo root.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from typing import Iterable, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libstark_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "stark_b200.h")

P_DEFAULT = 3221225473          # 3 * 2^30 + 1, the STARK-101 field
G_DEFAULT = 5                   # generator of F_p^*

u64, u64p, u8p, szt, vp = C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint8), C.c_size_t, C.c_void_p


class StarkError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[{code}] {msg}")
        self.code = code


_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """Loads libstark_b200.so; fails loudly when it has not been built (python build_ext.py)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise StarkError(4, f"{LIB_PATH} is missing: build it with `python build_ext.py` (nvcc, sm_100a). "
                            "There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)

    def sig(name, res, *args):
        f = getattr(L, name)
        f.restype, f.argtypes = res, list(args)

    I = C.c_int
    sig("stark_last_error", C.c_char_p)
    sig("stark_version", C.c_char_p)
    sig("stark_ctx_create", I, u64, u64, I, C.POINTER(vp))
    sig("stark_ctx_destroy", None, vp)
    sig("stark_ctx_sync", I, vp)
    sig("stark_ctx_modulus", u64, vp)
    sig("stark_ctx_generator", u64, vp)
    sig("stark_ctx_root_of_unity", u64, vp, C.c_uint)
    sig("stark_ctx_two_adicity", C.c_uint, vp)
    sig("stark_ctx_launch_count", C.c_ulonglong, vp)
    sig("stark_ctx_stream", vp, vp)
    sig("stark_measure_int_peak", I, vp, C.POINTER(C.c_double))
    sig("stark_measure_pipe_mix", I, vp, C.POINTER(C.c_double))
    sig("stark_ctx_set_timing", I, vp, I)
    sig("stark_ctx_read_timing", I, vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_ulonglong))
    sig("stark_vec_upload", I, vp, vp, szt, C.POINTER(vp))
    sig("stark_vec_alloc", I, vp, szt, C.POINTER(vp))
    sig("stark_vec_download", I, vp, szt, szt, vp)
    sig("stark_vec_from_device", I, vp, vp, szt, C.POINTER(vp))
    sig("stark_ntt_batch_dev", I, vp, vp, C.c_uint, I)
    sig("stark_pow_mul_dev", I, vp, vp, szt, szt, I, szt, u64, u64, C.c_uint)
    sig("stark_peer_alloc", I, vp, szt, C.POINTER(vp), vp)
    sig("stark_peer_open", I, vp, vp, C.POINTER(vp))
    sig("stark_peer_close", I, vp, vp)
    sig("stark_fourstep_phase_a", I, vp, vp, C.c_uint, u64, C.c_uint, C.c_uint, C.POINTER(vp), C.POINTER(vp), C.c_uint32)
    sig("stark_fourstep_phase_c", I, vp, vp, C.c_uint, C.c_uint, C.c_uint, C.POINTER(vp), C.POINTER(vp), C.c_uint32)
    sig("stark_fourstep_wait", I, vp, vp, C.c_uint, C.c_uint, C.c_uint32)
    sig("stark_mg_unique_id", I, vp)
    sig("stark_mg_create", I, vp, vp, C.c_uint, C.c_uint, C.POINTER(vp))
    sig("stark_mg_adopt", I, vp, vp, C.c_uint, C.c_uint, C.POINTER(vp))
    sig("stark_mg_destroy", None, vp)
    sig("stark_mg_rank", C.c_uint, vp)
    sig("stark_mg_world", C.c_uint, vp)
    sig("stark_mg_barrier", I, vp)
    sig("stark_mg_commit_columns", I, vp, szt, C.POINTER(vp), C.c_uint, u64, C.c_uint, u64, vp, C.POINTER(vp), C.POINTER(vp))
    sig("stark_mg_fourstep_lde", I, vp, vp, C.c_uint, u64, I, C.POINTER(vp))
    sig("stark_mg_commit_leaf_ranges", I, vp, vp, C.POINTER(vp), vp, vp)
    sig("stark_mg_fri_commit", I, vp, vp, C.c_uint, u64, I, vp, C.POINTER(vp))
    sig("stark_mg_decommit_fri", I, vp, szt, szt, vp)
    sig("stark_mg_stark101_prove", I, vp, u64, C.c_uint, C.c_uint, szt, I, vp)
    sig("stark_mg_fri_proof", vp, vp)
    sig("stark_mg_fri_subtree", vp, vp)
    sig("stark_mg_fri_destroy", None, vp)
    sig("stark_vec_len", szt, vp)
    sig("stark_vec_device_ptr", vp, vp)
    sig("stark_vec_destroy", None, vp)
    sig("stark_ntt", I, vp, vp, C.c_uint)
    sig("stark_intt", I, vp, vp, C.c_uint)
    sig("stark_coset_evaluate", I, vp, vp, szt, C.c_uint, u64, vp)
    sig("stark_coset_interpolate", I, vp, vp, C.c_uint, u64, vp)
    sig("stark_coset_lde", I, vp, vp, C.c_uint, u64, C.c_uint, u64, vp)
    sig("stark_batch_inverse", I, vp, vp, szt)
    sig("stark_quotient_pointwise", I, vp, vp, vp, szt, vp)
    sig("stark_coset_domain", I, vp, C.c_uint, u64, vp)
    sig("stark_coset_evaluate_dev", I, vp, vp, C.c_uint, u64, C.POINTER(vp))
    sig("stark_coset_interpolate_dev", I, vp, vp, u64, C.POINTER(vp))
    sig("stark_coset_lde_dev", I, vp, vp, u64, C.c_uint, u64, C.POINTER(vp))
    sig("stark_batch_inverse_dev", I, vp, vp, C.POINTER(vp))
    sig("stark_quotient_pointwise_dev", I, vp, vp, vp, C.POINTER(vp))
    sig("stark_merkle_commit", I, vp, vp, szt, C.POINTER(vp))
    sig("stark_merkle_commit_dev", I, vp, vp, C.POINTER(vp))
    sig("stark_merkle_root", I, vp, vp)
    sig("stark_merkle_root_hex", I, vp, C.c_char_p)
    sig("stark_merkle_num_leaves", szt, vp)
    sig("stark_merkle_depth", szt, vp)
    sig("stark_merkle_open", I, vp, szt, vp, szt, C.POINTER(szt))
    sig("stark_merkle_node", I, vp, szt, szt, vp)
    sig("stark_tree_destroy", None, vp)
    sig("stark_channel_new", I, u64, C.POINTER(vp))
    sig("stark_channel_destroy", None, vp)
    sig("stark_channel_send", I, vp, C.c_char_p, szt)
    sig("stark_channel_receive_random_field_element", I, vp, C.POINTER(u64))
    sig("stark_channel_receive_random_int", I, vp, u64, u64, I, C.POINTER(u64))
    sig("stark_channel_proof_size", szt, vp)
    sig("stark_channel_compressed_proof_size", szt, vp)
    sig("stark_channel_state", C.c_char_p, vp)
    sig("stark_channel_proof_len", szt, vp)
    sig("stark_channel_proof_msg", szt, vp, szt, C.POINTER(u8p))
    sig("stark_channel_proof_flat", szt, vp, vp)
    sig("stark_channel_compressed_len", szt, vp)
    sig("stark_channel_compressed_msg", szt, vp, szt, C.POINTER(u8p))
    sig("stark_fri_begin", I, vp, vp, szt, C.c_uint, u64, C.POINTER(vp), vp)
    sig("stark_fri_begin_dev", I, vp, vp, C.c_uint, u64, C.POINTER(vp), vp)
    sig("stark_fri_degree", I, vp, C.POINTER(C.c_longlong))
    sig("stark_fri_fold", I, vp, u64, vp)
    sig("stark_fri_final", I, vp, C.POINTER(u64), C.POINTER(szt))
    sig("stark_fri_num_layers", szt, vp)
    sig("stark_fri_layer_len", szt, vp, szt)
    sig("stark_fri_layer_read", I, vp, szt, szt, szt, vp)
    sig("stark_fri_layer_tree", vp, vp, szt)
    sig("stark_fri_open", I, vp, vp, szt, vp, szt, C.POINTER(szt))
    sig("stark_fri_begin_external", I, vp, vp, C.c_uint, u64, vp, vp, C.POINTER(vp))
    sig("stark_fri_open_layers", I, vp, szt, vp, szt, vp, szt, C.POINTER(szt))
    sig("stark_fri_destroy", None, vp)
    sig("stark_fri_commit", I, vp, vp, szt, C.c_uint, u64, vp, C.POINTER(vp))
    sig("stark_fri_commit_dev", I, vp, vp, C.c_uint, u64, vp, C.POINTER(vp))
    sig("stark_fri_begin_to_host", I, vp, vp, szt, C.c_uint, u64, vp, szt, C.POINTER(vp), vp)
    sig("stark_fri_layers_wait", I, vp)
    sig("stark_fri_layer_host_offset", szt, vp, szt)
    sig("stark_fri_commit_to_host", I, vp, vp, szt, C.c_uint, u64, vp, vp, szt, C.POINTER(vp))
    sig("stark_fri_commit_to_host_async", I, vp, vp, szt, C.c_uint, u64, vp, vp, szt, C.POINTER(vp))
    sig("stark_decommit_fri_layers", I, vp, szt, vp)
    sig("stark_decommit_fri", I, vp, szt, szt, vp)
    sig("stark101_prove", I, vp, u64, C.c_uint, C.c_uint, szt, vp)
    sig("stark101_trace_poly", I, vp, u64, C.c_uint, C.POINTER(vp), C.POINTER(u64))
    sig("stark101_composition_range", I, vp, vp, szt, szt, C.POINTER(u64), u64, C.c_uint, C.c_uint, C.POINTER(vp))
    sig("stark_merkle_verify", I, vp, szt, szt, u64, vp, szt, C.POINTER(I))
    sig("stark_fri_verify", I, vp, szt, u64, u64, C.c_uint, u64, szt, szt, C.c_uint, C.POINTER(I), C.c_char_p)
    sig("stark101_verify", I, vp, szt, u64, u64, u64, C.c_uint, C.c_uint, szt, C.POINTER(I), C.c_char_p)
    _lib = L
    return L


def _check(rc: int) -> None:
    if rc != 0:
        raise StarkError(rc, lib().stark_last_error().decode(errors="replace"))


def _arr(x) -> np.ndarray:
    """Contiguous uint64 view/copy of x (pinned torch tensors pass through .numpy() untouched)."""
    if isinstance(x, np.ndarray) and x.dtype == np.uint64 and x.flags["C_CONTIGUOUS"]:
        return x
    return np.ascontiguousarray(np.asarray(x, dtype=np.uint64))


def _ptr(a: np.ndarray):
    return C.c_void_p(a.ctypes.data)


class Context:
    """One per (device, modulus).  `generator` fixes w_n = generator^((p-1)/n)."""

    def __init__(self, modulus: int = P_DEFAULT, generator: int = G_DEFAULT, device: int = 0):
        h = vp()
        _check(lib().stark_ctx_create(modulus, generator, device, C.byref(h)))
        self.h = h
        self.modulus = lib().stark_ctx_modulus(h)
        self.generator = lib().stark_ctx_generator(h)
        self._children = weakref.WeakSet()      # device objects that must die before the context

    def _adopt(self, child):
        self._children.add(child)

    def close(self):
        if getattr(self, "h", None):
            kids = list(self._children)
            for child in [c for c in kids if not isinstance(c, MultiGpu)] + [c for c in kids if isinstance(c, MultiGpu)]:
                child.free()                     # groups last: their proofs and buffers refer to them
            lib().stark_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self): _check(lib().stark_ctx_sync(self.h))
    def root_of_unity(self, log_n: int) -> int: return lib().stark_ctx_root_of_unity(self.h, log_n)
    @property
    def two_adicity(self) -> int: return lib().stark_ctx_two_adicity(self.h)
    @property
    def launch_count(self) -> int: return lib().stark_ctx_launch_count(self.h)
    @property
    def stream(self) -> int: return lib().stark_ctx_stream(self.h) or 0

    def measure_int_peak(self) -> tuple[float, float]:
        t = (C.c_double * 2)()
        _check(lib().stark_measure_int_peak(self.h, t))
        return t[0], t[1]

    def measure_pipe_mix(self) -> list[float]:
        t = (C.c_double * 8)()
        _check(lib().stark_measure_pipe_mix(self.h, t))
        return list(t)

    def set_timing(self, on: bool): _check(lib().stark_ctx_set_timing(self.h, int(on)))

    def read_timing(self) -> dict:
        ms, un, ln = (C.c_double * 4)(), (C.c_double * 4)(), (C.c_ulonglong * 4)()
        _check(lib().stark_ctx_read_timing(self.h, ms, un, ln))
        names = ["merkle_leaf", "merkle_node", "ntt", "other"]
        return {names[i]: {"ms": ms[i], "units": un[i], "launches": int(ln[i])} for i in range(4)}

    # ---- device vectors
    def upload(self, host) -> "Vec":
        a = _arr(host)
        h = vp()
        _check(lib().stark_vec_upload(self.h, _ptr(a), a.size, C.byref(h)))
        return Vec(self, h)

    def from_device(self, device_ptr: int, n: int) -> "Vec":
        """Copies n u32 values from caller-owned device memory (a torch tensor's data_ptr())."""
        h = vp()
        _check(lib().stark_vec_from_device(self.h, C.c_void_p(device_ptr), n, C.byref(h)))
        return Vec(self, h)

    def ntt_batch_dev(self, v: "Vec", log_m: int, inverse: bool = False) -> None:
        _check(lib().stark_ntt_batch_dev(self.h, v.h, log_m, int(inverse)))

    def pow_mul_dev(self, v: "Vec", inner_len: int, outer0: int, product: bool, inner_stride: int, base: int, c0: int,
                    log_table: int) -> None:
        _check(lib().stark_pow_mul_dev(self.h, v.h, inner_len, outer0, int(product), inner_stride, base, c0, log_table))

    # ---- peer-visible buffers and the four-step phases that store into them (multi_gpu.FourStepP2P)
    def peer_alloc(self, n: int) -> tuple["Vec", bytes]:
        h = vp()
        handle = np.zeros(64, dtype=np.uint8)
        _check(lib().stark_peer_alloc(self.h, n, C.byref(h), _ptr(handle)))
        return Vec(self, h), handle.tobytes()

    def peer_open(self, handle: bytes) -> int:
        p = vp()
        hb = np.frombuffer(handle, dtype=np.uint8).copy()
        _check(lib().stark_peer_open(self.h, _ptr(hb), C.byref(p)))
        return p.value

    def peer_close(self, ptr: int) -> None:
        _check(lib().stark_peer_close(self.h, C.c_void_p(ptr)))

    def fourstep_phase_a(self, coeffs: "Vec", log_n: int, offset: int, world: int, rank: int, peer_rows: Sequence[int],
                         peer_flags: Optional[Sequence[int]] = None, epoch: int = 0) -> None:
        """peer_flags None: returns when the peer stores are complete (host barrier before phase C).  With the peers' flag
        arrays: nothing is synchronised, the hand-over happens on the device (epoch = 1, 2, 3, ... per transform)."""
        arr = (vp * world)(*[C.c_void_p(p) for p in peer_rows])
        fl = (vp * world)(*[C.c_void_p(p) for p in peer_flags]) if peer_flags is not None else None
        _check(lib().stark_fourstep_phase_a(self.h, coeffs.h, log_n, offset, world, rank, arr, fl, epoch))

    def fourstep_phase_c(self, rows: "Vec", log_n: int, world: int, rank: int, peer_blocks: Sequence[int],
                         peer_flags: Optional[Sequence[int]] = None, epoch: int = 0) -> None:
        arr = (vp * world)(*[C.c_void_p(p) for p in peer_blocks])
        fl = (vp * world)(*[C.c_void_p(p) for p in peer_flags]) if peer_flags is not None else None
        _check(lib().stark_fourstep_phase_c(self.h, rows.h, log_n, world, rank, arr, fl, epoch))

    def fourstep_wait(self, own_flags: int, slot: int, world: int, epoch: int) -> None:
        _check(lib().stark_fourstep_wait(self.h, C.c_void_p(own_flags), slot, world, epoch))

    def zeros(self, n: int) -> "Vec":
        h = vp()
        _check(lib().stark_vec_alloc(self.h, n, C.byref(h)))
        return Vec(self, h)

    # ---- polynomial (host buffers)
    def ntt(self, a, log_n: int) -> np.ndarray:
        a = _arr(a).copy()
        assert a.size == 1 << log_n
        _check(lib().stark_ntt(self.h, _ptr(a), log_n))
        return a

    def intt(self, a, log_n: int) -> np.ndarray:
        a = _arr(a).copy()
        assert a.size == 1 << log_n
        _check(lib().stark_intt(self.h, _ptr(a), log_n))
        return a

    def coset_evaluate(self, coeffs, log_n: int, offset: int = 1, out: Optional[np.ndarray] = None) -> np.ndarray:
        c = _arr(coeffs)
        if out is None:
            out = np.empty(1 << log_n, dtype=np.uint64)
        _check(lib().stark_coset_evaluate(self.h, _ptr(c), c.size, log_n, offset, _ptr(out)))
        return out

    def coset_interpolate(self, evals, log_n: int, offset: int = 1) -> np.ndarray:
        e = _arr(evals)
        assert e.size == 1 << log_n
        out = np.empty(1 << log_n, dtype=np.uint64)
        _check(lib().stark_coset_interpolate(self.h, _ptr(e), log_n, offset, _ptr(out)))
        return out

    def coset_lde(self, evals, log_n: int, offset_in: int, log_blowup: int, offset_out: int) -> np.ndarray:
        e = _arr(evals)
        assert e.size == 1 << log_n
        out = np.empty(1 << (log_n + log_blowup), dtype=np.uint64)
        _check(lib().stark_coset_lde(self.h, _ptr(e), log_n, offset_in, log_blowup, offset_out, _ptr(out)))
        return out

    def batch_inverse(self, a) -> np.ndarray:
        a = _arr(a).copy()
        _check(lib().stark_batch_inverse(self.h, _ptr(a), a.size))
        return a

    def quotient_pointwise(self, num, den) -> np.ndarray:
        n, d = _arr(num), _arr(den)
        assert n.size == d.size
        out = np.empty(n.size, dtype=np.uint64)
        _check(lib().stark_quotient_pointwise(self.h, _ptr(n), _ptr(d), n.size, _ptr(out)))
        return out

    def coset_domain(self, log_n: int, offset: int = 1) -> np.ndarray:
        out = np.empty(1 << log_n, dtype=np.uint64)
        _check(lib().stark_coset_domain(self.h, log_n, offset, _ptr(out)))
        return out

    # ---- polynomial (device vectors)
    def coset_evaluate_dev(self, coeffs: "Vec", log_n: int, offset: int = 1) -> "Vec":
        h = vp()
        _check(lib().stark_coset_evaluate_dev(self.h, coeffs.h, log_n, offset, C.byref(h)))
        return Vec(self, h)

    def coset_interpolate_dev(self, evals: "Vec", offset: int = 1) -> "Vec":
        h = vp()
        _check(lib().stark_coset_interpolate_dev(self.h, evals.h, offset, C.byref(h)))
        return Vec(self, h)

    def coset_lde_dev(self, evals: "Vec", offset_in: int, log_blowup: int, offset_out: int) -> "Vec":
        h = vp()
        _check(lib().stark_coset_lde_dev(self.h, evals.h, offset_in, log_blowup, offset_out, C.byref(h)))
        return Vec(self, h)

    def batch_inverse_dev(self, a: "Vec") -> "Vec":
        h = vp()
        _check(lib().stark_batch_inverse_dev(self.h, a.h, C.byref(h)))
        return Vec(self, h)

    def quotient_pointwise_dev(self, num: "Vec", den: "Vec") -> "Vec":
        h = vp()
        _check(lib().stark_quotient_pointwise_dev(self.h, num.h, den.h, C.byref(h)))
        return Vec(self, h)


class Vec:
    """Device-resident Vec<FieldElement<M>> (canonical u32 values in HBM)."""

    def __init__(self, ctx: Context, h):
        self.ctx, self.h = ctx, h
        ctx._adopt(self)

    def __len__(self): return lib().stark_vec_len(self.h)
    @property
    def device_ptr(self) -> int: return lib().stark_vec_device_ptr(self.h) or 0

    def download(self, offset: int = 0, n: Optional[int] = None) -> np.ndarray:
        n = len(self) - offset if n is None else n
        out = np.empty(n, dtype=np.uint64)
        _check(lib().stark_vec_download(self.h, offset, n, _ptr(out)))
        return out

    @property
    def __cuda_array_interface__(self):
        # int32 view of the u32 values: torch moves the bits (permute / all_to_all), it never does arithmetic on them
        return {"shape": (len(self),), "typestr": "<i4", "data": (self.device_ptr, False), "version": 3}

    def free(self):
        if getattr(self, "h", None):
            lib().stark_vec_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class MerkleTree:
    """MerkleTree<M> (src/merkle/mod.rs:5-27): `new` hashes and builds, `root` is 64 lowercase hex chars."""

    def __init__(self, ctx: Context, h, owned: bool = True, keep=None):
        self.ctx, self.h, self._owned, self._keep = ctx, h, owned, keep
        if owned:
            ctx._adopt(self)

    @classmethod
    def new(cls, ctx: Context, data) -> "MerkleTree":
        h = vp()
        if isinstance(data, Vec):
            _check(lib().stark_merkle_commit_dev(ctx.h, data.h, C.byref(h)))
        else:
            a = _arr(data)
            _check(lib().stark_merkle_commit(ctx.h, _ptr(a), a.size, C.byref(h)))
        return cls(ctx, h)

    def root(self) -> str:
        buf = C.create_string_buffer(65)
        _check(lib().stark_merkle_root_hex(self.h, buf))
        return buf.value.decode()

    def root_bytes(self) -> bytes:
        out = np.zeros(32, dtype=np.uint8)
        _check(lib().stark_merkle_root(self.h, _ptr(out)))
        return out.tobytes()

    @property
    def num_leaves(self) -> int: return lib().stark_merkle_num_leaves(self.h)
    @property
    def depth(self) -> int: return lib().stark_merkle_depth(self.h)

    def get_authentication_path(self, idx: int) -> bytes:
        out = np.zeros(32 * (self.depth + 1), dtype=np.uint8)
        n = szt(0)
        _check(lib().stark_merkle_open(self.h, idx, _ptr(out), out.size, C.byref(n)))
        return out[: n.value].tobytes()

    def node(self, level: int, j: int) -> bytes:
        out = np.zeros(32, dtype=np.uint8)
        _check(lib().stark_merkle_node(self.h, level, j, _ptr(out)))
        return out.tobytes()

    def free(self):
        if self._owned and getattr(self, "h", None):
            lib().stark_tree_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Channel:
    """Channel<M> (src/channel/channel.rs:14-95)."""

    def __init__(self, modulus: int = P_DEFAULT):
        h = vp()
        _check(lib().stark_channel_new(modulus, C.byref(h)))
        self.h, self.modulus = h, modulus

    def send(self, msg: bytes): _check(lib().stark_channel_send(self.h, msg, len(msg)))

    def receive_random_field_element(self) -> int:
        v = u64(0)
        _check(lib().stark_channel_receive_random_field_element(self.h, C.byref(v)))
        return v.value

    def receive_random_int(self, lo: int, hi: int, show_in_proof: bool = False) -> int:
        v = u64(0)
        _check(lib().stark_channel_receive_random_int(self.h, lo, hi, int(show_in_proof), C.byref(v)))
        return v.value

    def proof_size(self) -> int: return lib().stark_channel_proof_size(self.h)
    def compressed_proof_size(self) -> int: return lib().stark_channel_compressed_proof_size(self.h)
    @property
    def state(self) -> str: return lib().stark_channel_state(self.h).decode()

    @property
    def proof(self) -> list[bytes]:
        out = []
        for i in range(lib().stark_channel_proof_len(self.h)):
            p = u8p()
            n = lib().stark_channel_proof_msg(self.h, i, C.byref(p))
            out.append(bytes(C.cast(p, C.POINTER(C.c_uint8 * n)).contents) if n else b"")
        return out

    @property
    def compressed_proof(self) -> list[bytes]:
        out = []
        for i in range(lib().stark_channel_compressed_len(self.h)):
            p = u8p()
            n = lib().stark_channel_compressed_msg(self.h, i, C.byref(p))
            out.append(bytes(C.cast(p, C.POINTER(C.c_uint8 * n)).contents) if n else b"")
        return out

    def proof_flat(self) -> bytes:
        n = lib().stark_channel_proof_flat(self.h, None)
        out = np.zeros(max(n, 1), dtype=np.uint8)
        lib().stark_channel_proof_flat(self.h, _ptr(out))
        return out[:n].tobytes()

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().stark_channel_destroy(self.h)
                self.h = None
        except Exception:
            pass


class CosetFri:
    """CosetFri<M> (src/fri/coset_fri.rs:9-51): D = { offset * omega^i }."""

    def __init__(self, ctx: Context, offset: int, log_size: int):
        self.ctx, self.offset, self.log_size = ctx, offset, log_size
        self.omega = ctx.root_of_unity(log_size)
        self.domain_size = 1 << log_size

    def generate_coset_domain(self) -> np.ndarray:
        return self.ctx.coset_domain(self.log_size, self.offset)

    def next(self) -> "CosetFri":
        """First half squared (src/fri/fri_commit.rs:18-24): offset^2, omega^2, half the size."""
        return CosetFri(self.ctx, self.offset * self.offset % self.ctx.modulus, self.log_size - 1)


class FriProof:
    """FRIProof (src/fri/fri_commit.rs:9-13): fri_layers, fri_merkles, final_poly."""

    def __init__(self, ctx: Context, h):
        self.ctx, self.h = ctx, h
        ctx._adopt(self)

    @property
    def num_layers(self) -> int: return lib().stark_fri_num_layers(self.h)
    def layer_len(self, k: int) -> int: return lib().stark_fri_layer_len(self.h, k)

    def layer(self, k: int, offset: int = 0, n: Optional[int] = None, out: Optional[np.ndarray] = None) -> np.ndarray:
        """fri_layers[k] by value (u64 per element, as fri_commit.rs:117-121 returns them); `out` may be pinned memory"""
        n = self.layer_len(k) - offset if n is None else n
        if out is None:
            out = np.empty(n, dtype=np.uint64)
        assert out.dtype == np.uint64 and out.size >= n and out.flags["C_CONTIGUOUS"]
        _check(lib().stark_fri_layer_read(self.h, k, offset, n, _ptr(out)))
        return out[:n]

    def layers_wait(self) -> None:
        """Blocks until every layer copy issued so far by a *_to_host commit has landed in the host buffer."""
        _check(lib().stark_fri_layers_wait(self.h))

    def layer_host_offset(self, k: int) -> int:
        """Element offset of layer k in the host buffer of a *_to_host commit (-1: none)."""
        o = lib().stark_fri_layer_host_offset(self.h, k)
        return -1 if o == (1 << (8 * C.sizeof(szt))) - 1 else int(o)

    def tree(self, k: int) -> MerkleTree:
        return MerkleTree(self.ctx, lib().stark_fri_layer_tree(self.h, k), owned=False, keep=self)

    @property
    def degree(self) -> int:
        d = C.c_longlong(0)
        _check(lib().stark_fri_degree(self.h, C.byref(d)))
        return d.value

    def fold(self, beta: int) -> bytes:
        out = np.zeros(32, dtype=np.uint8)
        _check(lib().stark_fri_fold(self.h, beta, _ptr(out)))
        return out.tobytes()

    def final_poly(self) -> np.ndarray:
        v, n = u64(0), szt(0)
        _check(lib().stark_fri_final(self.h, C.byref(v), C.byref(n)))
        return np.array([v.value] if n.value else [], dtype=np.uint64)

    def open(self, indices: Sequence[int], first_layer: int = 0) -> bytes:
        idx = _arr(list(indices))
        n = szt(0)
        _check(lib().stark_fri_open_layers(self.h, first_layer, _ptr(idx), idx.size, None, 0, C.byref(n)))
        out = np.zeros(max(n.value, 1), dtype=np.uint8)
        _check(lib().stark_fri_open_layers(self.h, first_layer, _ptr(idx), idx.size, _ptr(out), out.size, C.byref(n)))
        return out[: n.value].tobytes()

    def free(self):
        if getattr(self, "h", None):
            if not getattr(self, "_borrowed", False):
                lib().stark_fri_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def fri_begin(ctx: Context, coeffs, log_n: int, offset: int, layers_out: Optional[np.ndarray] = None) -> tuple[FriProof, bytes]:
    """Layer 0 only (fri_commit.rs:78-86 without the send): returns the proof object and the root.  `layers_out`: as in
    fri_commit -- layer 0 and every later fold's layer stream to that host buffer; call `layers_wait()` before reading."""
    h = vp()
    root = np.zeros(32, dtype=np.uint8)
    if layers_out is not None:
        assert not isinstance(coeffs, Vec) and layers_out.dtype == np.uint64 and layers_out.flags["C_CONTIGUOUS"]
        c = _arr(coeffs)
        _check(lib().stark_fri_begin_to_host(ctx.h, _ptr(c), c.size, log_n, offset, _ptr(layers_out), layers_out.size, C.byref(h), _ptr(root)))
        pr = FriProof(ctx, h)
        pr._keep_layers = layers_out
        return pr, root.tobytes()
    if isinstance(coeffs, Vec):
        _check(lib().stark_fri_begin_dev(ctx.h, coeffs.h, log_n, offset, C.byref(h), _ptr(root)))
    else:
        c = _arr(coeffs)
        _check(lib().stark_fri_begin(ctx.h, _ptr(c), c.size, log_n, offset, C.byref(h), _ptr(root)))
    return FriProof(ctx, h), root.tobytes()


def fri_begin_external(ctx: Context, coeffs: "Vec", log_n: int, offset: int, layer0: "Vec", root0: bytes) -> FriProof:
    """Adopts a layer 0 that was evaluated and hashed in leaf ranges on several GPUs (multi_gpu.py)."""
    h = vp()
    r = np.frombuffer(root0, dtype=np.uint8).copy()
    _check(lib().stark_fri_begin_external(ctx.h, coeffs.h, log_n, offset, layer0.h, _ptr(r), C.byref(h)))
    pr = FriProof(ctx, h)
    pr._layer0 = layer0            # the proof shares the vector's device memory
    return pr


def fri_commit(ctx: Context, poly, domain: CosetFri, channel: Channel, layers_out: Optional[np.ndarray] = None,
               wait: bool = True) -> FriProof:
    """fri_commit(poly, domain, &mut channel) (src/fri/fri_commit.rs:72-122).  With `layers_out` (uint64, >= 2^(log_size+1)
    elements always suffice; pinned memory keeps the copies asynchronous) every layer is also returned BY VALUE, layer k at
    `proof.layer_host_offset(k)`, copied on a second stream under the hashing of the following layers; complete on return,
    or -- with wait=False -- once `proof.layers_wait()` has been called (the last copies then run under the query phase)."""
    h = vp()
    if layers_out is not None:
        assert not isinstance(poly, Vec), "fri_commit(layers_out=...): host coefficients"
        assert layers_out.dtype == np.uint64 and layers_out.flags["C_CONTIGUOUS"]
        c = _arr(poly)
        fn = lib().stark_fri_commit_to_host if wait else lib().stark_fri_commit_to_host_async
        _check(fn(ctx.h, _ptr(c), c.size, domain.log_size, domain.offset, channel.h, _ptr(layers_out), layers_out.size, C.byref(h)))
        pr = FriProof(ctx, h)
        pr._keep_layers = layers_out
        return pr
    if isinstance(poly, Vec):
        _check(lib().stark_fri_commit_dev(ctx.h, poly.h, domain.log_size, domain.offset, channel.h, C.byref(h)))
    else:
        c = _arr(poly)
        _check(lib().stark_fri_commit(ctx.h, _ptr(c), c.size, domain.log_size, domain.offset, channel.h, C.byref(h)))
    return FriProof(ctx, h)


def decommit_fri_layers(index: int, proof: FriProof, channel: Channel) -> None:
    """src/fri/fri_commit.rs:137-165."""
    _check(lib().stark_decommit_fri_layers(proof.h, index, channel.h))


def decommit_fri(num_queries: int, max_index: int, proof: FriProof, channel: Channel) -> None:
    """src/fri/fri_commit.rs:168-179."""
    _check(lib().stark_decommit_fri(proof.h, num_queries, max_index, channel.h))


def stark101_prove(ctx: Context, channel: Channel, a1: int = 3141592, log_trace: int = 10, log_blowup: int = 3,
                   num_queries: int = 3) -> None:
    """Build-defined FibonacciSq prover."""
    _check(lib().stark101_prove(ctx.h, a1, log_trace, log_blowup, num_queries, channel.h))


def stark101_trace_poly(ctx: Context, a1: int = 3141592, log_trace: int = 10) -> tuple["Vec", int]:
    """(coefficients of the trace polynomial f as a device Vec of 2^log_trace entries, a_{T-2})."""
    out, last = C.c_void_p(), C.c_uint64(0)
    _check(lib().stark101_trace_poly(ctx.h, a1, log_trace, C.byref(out), C.byref(last)))
    return Vec(ctx, out), int(last.value)


def stark101_composition_range(ctx: Context, f_block: "Vec", start: int, count: int, alpha, last_value: int,
                               log_trace: int, log_blowup: int) -> "Vec":
    """CP on the points start .. start+count-1 of the coset; f_block = f on that range + the next 2*blowup points."""
    al = (C.c_uint64 * 3)(*[int(x) for x in alpha])
    out = C.c_void_p()
    _check(lib().stark101_composition_range(ctx.h, f_block.h, start, count, al, last_value, log_trace, log_blowup, C.byref(out)))
    return Vec(ctx, out)


def merkle_validate(root: bytes, n_leaves: int, idx: int, value: int, path: bytes) -> bool:
    """MerkleTree::validate (called at src/fri/fri_verify.rs:109, never defined in the reference)."""
    r = np.frombuffer(root, dtype=np.uint8).copy()
    p = np.frombuffer(path, dtype=np.uint8).copy() if path else np.zeros(1, dtype=np.uint8)
    ok = C.c_int(0)
    _check(lib().stark_merkle_verify(_ptr(r), n_leaves, idx, value, _ptr(p), len(path), C.byref(ok)))
    return bool(ok.value)


def verify_fri(proof_flat: bytes, log_n: int, offset: int, num_queries: int, max_index: int, log_degree_bound: int,
               modulus: int = P_DEFAULT, generator: int = G_DEFAULT) -> tuple[bool, str]:
    """verify_fri (src/fri/fri_verify.rs:12-177, completed): replays a flattened proof; host side, no GPU needed.
    `log_degree_bound`: the committed polynomial is claimed to have at most 2^log_degree_bound coefficients, so the
    proof may hold at most that many folds (the reference's `expected_num_layers`, fri_verify.rs:15)."""
    buf = np.frombuffer(proof_flat, dtype=np.uint8).copy()
    ok, reason = C.c_int(0), C.create_string_buffer(160)
    _check(lib().stark_fri_verify(_ptr(buf), buf.size, modulus, generator, log_n, offset, num_queries, max_index, log_degree_bound,
                                  C.byref(ok), reason))
    return bool(ok.value), reason.value.decode()


def stark101_statement(modulus: int, generator: int, log_trace: int, log_blowup: int, num_queries: int, claimed_last: int) -> bytes:
    """First transcript message of the FibonacciSq STARK: the public inputs, 8 big-endian bytes each."""
    return b"".join(int(v).to_bytes(8, "big") for v in (modulus, generator % modulus, log_trace, log_blowup, num_queries, claimed_last % modulus))


def stark101_verify(proof_flat: bytes, claimed_last: int, log_trace: int = 10, log_blowup: int = 3, num_queries: int = 3,
                    modulus: int = P_DEFAULT, generator: int = G_DEFAULT) -> tuple[bool, str]:
    """Verifier of the FibonacciSq STARK transcript (public input: the claimed a_{T-2}); host side."""
    buf = np.frombuffer(proof_flat, dtype=np.uint8).copy()
    ok, reason = C.c_int(0), C.create_string_buffer(160)
    _check(lib().stark101_verify(_ptr(buf), buf.size, modulus, generator, claimed_last, log_trace, log_blowup, num_queries,
                                 C.byref(ok), reason))
    return bool(ok.value), reason.value.decode()


class MgFri:
    """FRIProof whose layer 0 is spread over a MultiGpu group (stark_mg_fri)."""

    def __init__(self, mg: "MultiGpu", h):
        self.mg, self.h = mg, h
        mg.ctx._adopt(self)

    @property
    def proof(self) -> Optional[FriProof]:
        """rank 0: every layer's values and the small layers' trees, borrowed.  On the other ranks None, or -- when the
        large layers >= 1 are hashed in leaf ranges (more than one GPU) -- the replicated large layers"""
        h = lib().stark_mg_fri_proof(self.h)
        if not h:
            return None
        pr = FriProof.__new__(FriProof)
        pr.ctx, pr.h, pr._borrowed = self.mg.ctx, h, True
        return pr

    def free(self):
        if getattr(self, "h", None):
            lib().stark_mg_fri_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class MultiGpu:
    """One group of ranks, one GPU each (stark_mg): the C-level multi-GPU entry points of SURVEY.md 8(e).
    The library owns its NCCL communicator; only the 128-byte unique id travels through the host program."""

    def __init__(self, ctx: Context, rank: int, world: int, unique_id: Optional[bytes] = None):
        h = vp()
        uid = np.frombuffer(unique_id, dtype=np.uint8).copy() if unique_id is not None else np.zeros(128, dtype=np.uint8)
        _check(lib().stark_mg_create(ctx.h, _ptr(uid), rank, world, C.byref(h)))
        self.ctx, self.h, self.rank, self.world = ctx, h, rank, world
        ctx._adopt(self)

    @staticmethod
    def unique_id() -> bytes:
        out = np.zeros(128, dtype=np.uint8)
        _check(lib().stark_mg_unique_id(_ptr(out)))
        return out.tobytes()

    @classmethod
    def from_torch(cls, ctx: Context, group=None) -> "MultiGpu":
        """Bootstraps the group over an initialised torch.distributed process group (any backend): rank 0 makes the id,
        the others receive it."""
        import torch.distributed as dist
        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            return cls(ctx, 0, 1)
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=group)
        return cls(ctx, rank, world, box[0])

    def barrier(self): _check(lib().stark_mg_barrier(self.h))

    def commit_columns(self, columns, n_cols: int, log_rows: int, offset_in: int, log_blowup: int, offset_out: int, keep: bool = False):
        """columns: {c: uint64 array} or a list indexed by column holding at least this rank's columns (c % world == rank).
        Returns the roots of all columns (list of 32-byte strings); with keep=True also {c: (Vec, MerkleTree)} for the
        rank's own columns."""
        get = (lambda c: columns.get(c)) if isinstance(columns, dict) else (lambda c: columns[c] if c < len(columns) else None)
        arrs = {c: _arr(get(c)) for c in range(self.rank, n_cols, self.world)}
        ptrs = (vp * n_cols)(*[C.c_void_p(arrs[c].ctypes.data) if c in arrs else None for c in range(n_cols)])
        roots = np.zeros(n_cols * 32, dtype=np.uint8)
        ldes = (vp * n_cols)() if keep else None
        trees = (vp * n_cols)() if keep else None
        _check(lib().stark_mg_commit_columns(self.h, n_cols, ptrs, log_rows, offset_in, log_blowup, offset_out, _ptr(roots), ldes, trees))
        out = [roots[32 * c:32 * c + 32].tobytes() for c in range(n_cols)]
        if not keep:
            return out
        kept = {c: (Vec(self.ctx, C.c_void_p(ldes[c])), MerkleTree(self.ctx, C.c_void_p(trees[c]))) for c in arrs}
        return out, kept

    def fourstep_lde(self, coeffs: "Vec", log_n: int, offset: int, transport: int = 1) -> "Vec":
        """This rank's natural-order block of the evaluations; transport 0 = NCCL all-to-all, 1 = peer-memory stores.
        The block aliases a buffer the next transform of the same size overwrites."""
        h = vp()
        _check(lib().stark_mg_fourstep_lde(self.h, coeffs.h, log_n, offset, transport, C.byref(h)))
        return Vec(self.ctx, h)

    def commit_leaf_ranges(self, block: "Vec") -> tuple["MerkleTree", bytes, list[bytes]]:
        h = vp()
        root = np.zeros(32, dtype=np.uint8)
        subs = np.zeros(32 * self.world, dtype=np.uint8)
        _check(lib().stark_mg_commit_leaf_ranges(self.h, block.h, C.byref(h), _ptr(root), _ptr(subs)))
        return MerkleTree(self.ctx, h, keep=block), root.tobytes(), [subs[32 * r:32 * r + 32].tobytes() for r in range(self.world)]

    def fri_commit(self, coeffs: "Vec", log_n: int, offset: int, channel: Optional[Channel], transport: int = 1) -> MgFri:
        h = vp()
        _check(lib().stark_mg_fri_commit(self.h, coeffs.h, log_n, offset, transport, channel.h if channel is not None else None, C.byref(h)))
        return MgFri(self, h)

    def stark101_prove(self, channel: Optional[Channel], a1: int = 3141592, log_trace: int = 10, log_blowup: int = 3, num_queries: int = 3,
                       transport: int = 1) -> None:
        """The FibonacciSq prover over the group (collective; `channel` on rank 0 only): stark101_prove's transcript."""
        _check(lib().stark_mg_stark101_prove(self.h, a1, log_trace, log_blowup, num_queries, transport,
                                             channel.h if channel is not None else None))

    def decommit_fri(self, f: MgFri, num_queries: int, max_index: int, channel: Optional[Channel]) -> None:
        _check(lib().stark_mg_decommit_fri(f.h, num_queries, max_index, channel.h if channel is not None else None))

    def free(self):
        if getattr(self, "h", None):
            lib().stark_mg_destroy(self.h)
            self.h = None

    def close(self):
        self.free()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def exported_symbols() -> list[str]:
    """Function names declared in include/stark_b200.h (for the symbol-table test)."""
    import re
    txt = open(HEADER_PATH).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(stark\w*)\s*\(", txt)))
