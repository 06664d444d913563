#!/bin/bash
# where the latency-bound tail kernel takes over from the 3-levels-per-launch kernel (items = 256 * CTAs)
for defs in "-DSTARK_TAIL_MAX_CTAS=128" "-DSTARK_TAIL_MAX_CTAS=512" "-DSTARK_TAIL_MAX_CTAS=1024" "-DSTARK_TAIL_MAX_CTAS=2048" "-DSTARK_TAIL_MAX_CTAS=4096" "-DSTARK_TAIL_MAX_CTAS=16384"; do
  export STARK_NVCC_DEFS="$defs"
  touch stark-prover_b200/csrc/merkle.cu
  python build_ext.py > /dev/null 2>&1 || { echo "build failed for $defs"; continue; }
  python bench.py --no-cpu-baseline --no-pipelined --no-kernel-timing 2>/dev/null | python -c "
import json,sys,os; d=json.loads(sys.stdin.read()); print(os.environ.get('STARK_NVCC_DEFS'), 'ms_per_step', round(d['ms_per_step'],3), d.get('host_breakdown_ms'))"
done
