#!/bin/bash
# Kernel-variant sweep on the GPU box (nvcc is in the image): rebuilds merkle.cu with each -D set and times the tree alone.
# Results of the round-1 sweeps: profiles/r01_variants.txt.
for defs in "" "-DSTARK_MERKLE_MIN_BLOCKS=9" "-DSTARK_MERKLE_THREADS=256 -DSTARK_MERKLE_MIN_BLOCKS=4" "-DSTARK_MERKLE_THREADS=64" \
            "-DSTARK_SHA_SHR_ON_FMA=1" "-DSTARK_SHA_ADDS_ON_FMA=0" "-DSTARK_SHA_SCHED_IADD3=0x5555u" "-DSTARK_SHA_SCHED_IADD3=0xFFFFu"; do
  export STARK_NVCC_DEFS="$defs"
  touch stark-prover_b200/csrc/merkle.cu
  python build_ext.py > /dev/null 2>&1 || { echo "build failed for $defs"; continue; }
  python tools/bench_merkle.py 24
  python tools/bench_merkle.py 24
done
