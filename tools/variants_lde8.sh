#!/bin/bash
# CTA-size variants of the blow-up-by-8 transform on the GPU box (threads per tile = 2R >> STARK_LDE8_TSHIFT)
for v in ${@:-0 1 2}; do
  export STARK_NVCC_DEFS="-DSTARK_LDE8_TSHIFT=$v"
  touch stark-prover_b200/csrc/ntt.cu
  python build_ext.py > /dev/null 2>&1 || { echo "build failed for $v"; continue; }
  echo "== STARK_LDE8_TSHIFT=$v"
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "coset_lde or coset_evaluate or large_modulus or blowup8 or cfg2" 2>&1 | tail -1
  python tools/bench_ops.py 2>&1 | grep -v "^{" | grep "coset_lde"
  python bench.py --steps 5 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench ms_per_step', d['ms_per_step'], 'ntt ms', d['kernel_ms']['ntt'])"
done
export STARK_NVCC_DEFS=""
touch stark-prover_b200/csrc/ntt.cu
python build_ext.py > /dev/null 2>&1
