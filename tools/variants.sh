#!/bin/bash
# kernel-variant sweep on the GPU box: rebuild merkle.cu with different -D flags and time the commitment
for defs in "" "-DSTARK_MERKLE_THREADS=64" "-DSTARK_MERKLE_THREADS=256" "-DSTARK_MERKLE_THREADS=512" "-DSTARK_MERKLE_THREADS=256 -DSTARK_MERKLE_MIN_BLOCKS=3"; do
  export STARK_NVCC_DEFS="$defs"
  touch stark-prover_b200/csrc/merkle.cu
  python build_ext.py > /dev/null 2>&1 || { echo "build failed for $defs"; continue; }
  python tools/bench_merkle.py 24
done
