#!/bin/bash
# field-arithmetic variants on the GPU box: rebuild everything with each -D set, time the per-op kernels.
for defs in "" "-DSTARK_MONT_WIDE=1"; do
  export STARK_NVCC_DEFS="$defs"
  touch stark-prover_b200/csrc/*.cu
  python build_ext.py > /dev/null 2>&1 || { echo "build failed for $defs"; continue; }
  echo "== defs='$defs'"
  python tools/bench_ops.py 2>&1 | grep -v "^{" | grep "2^24\|2^25"
  python tools/bench_inverse.py 24
done
