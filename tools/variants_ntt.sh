#!/bin/bash
# butterfly variants on the GPU box: rebuild with each -DSTARK_NTT_BFLY mask, check parity on the transform tests, time the per-op kernels.
for v in ${@:-0 2 3 7}; do
  export STARK_NVCC_DEFS="-DSTARK_NTT_BFLY=$v"
  touch stark-prover_b200/csrc/*.cu
  python build_ext.py > /dev/null 2>&1 || { echo "build failed for $v"; continue; }
  echo "== STARK_NTT_BFLY=$v"
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "natural_transform_every or coset_lde or large_modulus or blowup8" 2>&1 | tail -1
  python tools/bench_ops.py 2>&1 | grep -v "^{" | grep "coset.*2^2[45]"
done
export STARK_NVCC_DEFS=""
touch stark-prover_b200/csrc/*.cu
python build_ext.py > /dev/null 2>&1
