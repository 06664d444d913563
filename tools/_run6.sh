mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/r2f_pytest.log
tail -2 gpurun_out/r2f_pytest.log
python tools/prof_fourstep_emulated.py > gpurun_out/r2f_fs_emulated.txt 2>&1; cat gpurun_out/r2f_fs_emulated.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2f_fs_launches.csv python tools/prof_fourstep_emulated.py > gpurun_out/r2f_fs_ncu.log 2>&1
python bench.py --no-cpu-baseline --no-prove --no-sharded --no-pipelined > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_traffic.csv python bench.py --profile-mode --steps 1 --warmup 1 > gpurun_out/r2f_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"merkle_subtree|lde8_pass|merkle_tail" --launch-skip 18 -c 14 -o gpurun_out/r2f_prof_full python bench.py --profile-mode --steps 1 --warmup 1 > gpurun_out/r2f_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"nat_" -c 3 -o gpurun_out/r2f_prof_nat python tools/prof_ntt.py > gpurun_out/r2f_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"batch_inverse" --launch-skip 2 -c 2 -o gpurun_out/r2f_prof_inv python tools/prof_inverse.py > gpurun_out/r2f_ncu4.log 2>&1
for f in gpurun_out/r2f_ncu1.log gpurun_out/r2f_ncu2.log gpurun_out/r2f_ncu3.log gpurun_out/r2f_ncu4.log; do tail -n 2 $f; done
ls -la gpurun_out/*.ncu-rep
