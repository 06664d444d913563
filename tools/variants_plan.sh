#!/bin/bash
# pass-width plans of the transforms (STARK_NTT_PLAN, run-time): parity on the transform tests, then per-op timings
for plan in "$@"; do
  export STARK_NTT_PLAN="$plan"
  echo "== plan='$plan'"
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "natural_transform_every or coset_lde or lde_roundtrip or cfg2" 2>&1 | tail -1
  python tools/bench_ops.py 2>&1 | grep -v "^{" | grep "coset.*2^2[45]"
done
