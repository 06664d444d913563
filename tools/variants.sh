#!/bin/bash
# kernel-variant sweep on the GPU box: rebuild merkle.cu with different -D flags and time the commitment
for defs in "" "-DSTARK_SHA_ADDS_ON_FMA=0" "-DSTARK_SHA_ADDS_ON_FMA=2" "-DSTARK_MERKLE_MIN_BLOCKS=12" "-DSTARK_MERKLE_MIN_BLOCKS=14" "-DSTARK_SHA_ADDS_ON_FMA=2 -DSTARK_MERKLE_MIN_BLOCKS=12"; do
  export STARK_NVCC_DEFS="$defs"
  touch stark-prover_b200/csrc/merkle.cu
  python build_ext.py > /dev/null 2>&1 || { echo "build failed for $defs"; continue; }
  grep -A2 "subtree_kernelILi0" build/merkle.o.log | grep -o "Used [0-9]* registers" | head -1
  python tools/bench_merkle.py 24
done
