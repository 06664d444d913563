#!/bin/bash
# Round-end evidence on the GPU box: GPU tests, both bench arms, per-op timings, the ncu launch list + DRAM traffic of one
# step, the launch list of smoke(), ncu --set full of the hashing / transform / inverse kernels.  Everything lands in
# gpurun_out/<prefix>_*; summarise with tools/ncu_summary.py into profiles/.   usage: tools/capture_profiles.sh [prefix]
P=${1:-f}
set -x
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS} 2>&1 | tail -3) > gpurun_out/${P}_pytest.log
python bench.py --impl reference > gpurun_out/${P}_bench_reference.json 2> gpurun_out/${P}_bench_reference.err
python bench.py > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.err
python tools/bench_ops.py > gpurun_out/${P}_ops.txt 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/${P}_traffic.csv python bench.py --profile-mode --steps 1 --warmup 1 > gpurun_out/${P}_ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${P}_smoke_launches.csv python __graft_entry__.py --smoke > gpurun_out/${P}_ncu0.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"merkle_subtree|merkle_tail|lde8_pass" -c 15 -o gpurun_out/${P}_prof_full python bench.py --profile-mode --steps 1 --warmup 1 > gpurun_out/${P}_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"nat_" -c 6 -o gpurun_out/${P}_prof_nat python tools/prof_ntt.py > gpurun_out/${P}_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"batch_inverse" -c 2 -o gpurun_out/${P}_prof_inv python tools/prof_inverse.py > gpurun_out/${P}_ncu4.log 2>&1
for f in gpurun_out/${P}_ncu0.log gpurun_out/${P}_ncu1.log gpurun_out/${P}_ncu2.log gpurun_out/${P}_ncu3.log gpurun_out/${P}_ncu4.log; do tail -n 2 $f; done
cat gpurun_out/${P}_pytest.log
