"""The index algebra of the permutation-free natural-order transform (csrc/ntt.cu, ntt_natural), restated in numpy
and checked against the oracle's NTT: digits most significant first, in-place strided passes with the per-row
twiddle W^(K n_i), K the output index accumulated so far, and a last pass that writes X[K + 2^(log_n - r_m) k_m].
Runs without a GPU; the kernels themselves are compared with the oracle in test_gpu_parity.py."""
import numpy as np
import pytest

from oracle import pyoracle as orc

P = 3221225473


def small_dft(col, w_r, p):
    """size-R DFT of a python-int list with root w_r (R <= 64 here): X[k] = sum_n x[n] w_r^(n k)."""
    r = len(col)
    pw = [pow(w_r, e, p) for e in range(r)]
    return [sum(col[n] * pw[(n * k) % r] for n in range(r)) % p for k in range(r)]


def natural_transform_model(x, bits, w, p):
    log_n = sum(bits)
    n = 1 << log_n
    a = [int(v) for v in x]                      # the work array, in place until the last pass
    out = [0] * n
    hi = log_n
    prev = []                                    # (width, offset in the output index) of the digits already done
    off = 0
    for i, r in enumerate(bits):
        lo = hi - r
        last = i + 1 == len(bits)
        big_w = pow(w, 1 << lo, p)               # W = w_{2^(log_n - lo)}
        w_r = pow(w, n >> r, p)                  # root of the size-2^r transforms
        for high in range(1 << (log_n - hi)):    # address bits above this digit: [k1][k2]..
            k_acc, rem = 0, high
            for width, o in reversed(prev):      # the last processed digit sits lowest in `high`
                k_acc |= (rem & ((1 << width) - 1)) << o
                rem >>= width
            for low in range(1 << lo):
                base = (high << hi) | low
                col = [a[base + (t << lo)] * pow(big_w, k_acc * t, p) % p for t in range(1 << r)]
                res = small_dft(col, w_r, p)
                for k in range(1 << r):
                    if last:
                        out[k_acc + (k << (log_n - r))] = res[k]
                    else:
                        a[base + (k << lo)] = res[k]
        prev.append((r, off))
        off += r
        hi = lo
    return np.array(out, dtype=np.uint64)


@pytest.mark.parametrize("bits", [[5, 5], [6, 5], [4, 3, 4], [3, 3, 2, 3]])
def test_digit_split_model_matches_oracle(bits):
    log_n = sum(bits)
    x = orc.synthetic_column(17 + log_n, 1 << log_n)
    w = orc.root_of_unity(log_n)
    assert np.array_equal(natural_transform_model(x, bits, w, P), orc.ntt(x, log_n, w, P))
