mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/r2g_pytest.log
tail -2 gpurun_out/r2g_pytest.log
python tools/prof_fourstep_emulated.py > gpurun_out/r2g_fs_emulated.txt 2>&1; cat gpurun_out/r2g_fs_emulated.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2g_fs_launches.csv python tools/prof_fourstep_emulated.py > gpurun_out/r2g_fs_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"merkle_subtree|lde8_pass|merkle_tail" --launch-skip 37 -c 15 -o gpurun_out/r2g_prof_full python bench.py --profile-mode --steps 1 --warmup 1 > gpurun_out/r2g_ncu2.log 2>&1
tail -n 2 gpurun_out/r2g_ncu2.log
