#!/bin/bash
for defs in "" "-DSTARK_SHA_SHR_ON_FMA=0"; do
  export STARK_NVCC_DEFS="$defs"
  touch stark-prover_b200/csrc/merkle.cu
  python build_ext.py > /dev/null 2>&1 || { echo "build failed for $defs"; continue; }
  python tools/bench_merkle.py 24
  python tools/bench_merkle.py 12
done
