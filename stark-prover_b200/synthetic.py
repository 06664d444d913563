"""Synthetic inputs of SURVEY.md 8(d): values = splitmix64(seed) % p, identical in C++/Python/CUDA (the reference's
own benches draw `next_u64() % modulus` from a seeded generator, benches/poly_ops.rs:27-30,44)."""
from __future__ import annotations

import numpy as np

P_DEFAULT = 3221225473


def splitmix64(seed: int, n: int) -> np.ndarray:
    """n outputs of splitmix64 seeded with `seed` (vectorised)."""
    with np.errstate(over="ignore"):
        k = np.arange(1, n + 1, dtype=np.uint64)
        z = np.uint64(seed) + k * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def synthetic_column(seed: int, n: int, modulus: int = P_DEFAULT) -> np.ndarray:
    return splitmix64(seed, n) % np.uint64(modulus)


def synthetic_poly_exact_degree(seed: int, n_coeffs: int, modulus: int = P_DEFAULT) -> np.ndarray:
    """Coefficients of a polynomial of exact degree n_coeffs - 1 (top coefficient forced non-zero)."""
    c = synthetic_column(seed, n_coeffs, modulus)
    if c[-1] == 0:
        c[-1] = 1
    return c
