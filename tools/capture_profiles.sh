#!/bin/bash
# Round-end evidence on the GPU box: both bench arms, the ncu launch list + DRAM traffic of one step, ncu --set full of the
# hashing / transform kernels.  Summarise with tools/ncu_summary.py into profiles/.
set -x
mkdir -p gpurun_out
python bench.py --impl reference > gpurun_out/f_bench_reference.json 2> gpurun_out/f_bench_reference.err
python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_traffic.csv python bench.py --profile-mode --steps 1 --warmup 1 > gpurun_out/f_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"merkle_subtree|lde8_pass" --launch-skip 16 -c 12 -o gpurun_out/f_prof_full python bench.py --profile-mode --steps 1 --warmup 1 > gpurun_out/f_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"nat_" -c 3 -o gpurun_out/f_prof_nat python tools/prof_ntt.py > gpurun_out/f_ncu3.log 2>&1
for f in gpurun_out/f_ncu1.log gpurun_out/f_ncu2.log gpurun_out/f_ncu3.log; do tail -n 2 $f; done
