#!/bin/bash
# batch-inverse parameter sweep on the GPU box: elements per thread, CTA size, register cap.
for defs in "-DSTARK_INV_K=12 -DSTARK_INV_MINB=5" "-DSTARK_INV_K=12 -DSTARK_INV_MINB=6" "-DSTARK_INV_K=8 -DSTARK_INV_MINB=8" "-DSTARK_INV_K=12 -DSTARK_INV_THREADS=128 -DSTARK_INV_MINB=10" \
            "-DSTARK_INV_K=12 -DSTARK_INV_THREADS=128 -DSTARK_INV_MINB=12" "-DSTARK_INV_K=16 -DSTARK_INV_MINB=6" "-DSTARK_INV_K=10 -DSTARK_INV_MINB=6" "-DSTARK_INV_K=14 -DSTARK_INV_MINB=5" "-DSTARK_INV_K=12 -DSTARK_INV_THREADS=512 -DSTARK_INV_MINB=3"; do
  export STARK_NVCC_DEFS="$defs"
  touch stark-prover_b200/csrc/fri.cu
  python build_ext.py > /dev/null 2>&1 || { echo "build failed for $defs"; continue; }
  python tools/bench_inverse.py 24
done
