#!/bin/bash
# field / butterfly variants on the GPU box: rebuild with each set of -D flags, check parity on the transform tests, time the per-op kernels.
#   STARK_FIELD_CARRY  0 = compare + select corrections, 1 = carry-predicated corrections (field.cuh)
#   STARK_CORR_FMA     which predicated corrections run on the FMA pipe (bit 0 product, 1 add, 2 subtract)
#   STARK_NTT_REGCAP   registers per thread the transform kernels are compiled for
# usage: tools/variants_ntt.sh "<defs 1>" "<defs 2>" ...
if [ $# -eq 0 ]; then set -- "-DSTARK_FIELD_CARRY=0" "" "-DSTARK_NTT_REGCAP=48" "-DSTARK_CORR_FMA=0"; fi
for v in "$@"; do
  export STARK_NVCC_DEFS="$v"
  touch stark-prover_b200/csrc/*.cu
  python build_ext.py > /dev/null 2>&1 || { echo "build failed for $v"; continue; }
  echo "== defs='$v'"
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "natural_transform_every or coset_lde or large_modulus or blowup8 or other_moduli or batch_inverse or fold" 2>&1 | tail -1
  python tools/bench_ops.py 2>&1 | grep -v "^{" | grep "2^2[45]"
done
export STARK_NVCC_DEFS=""
touch stark-prover_b200/csrc/*.cu
python build_ext.py > /dev/null 2>&1
