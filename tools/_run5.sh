mkdir -p gpurun_out
python tools/prof_fourstep_emulated.py > gpurun_out/r2e_fs_emulated.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2e_fs_launches.csv python tools/prof_fourstep_emulated.py > gpurun_out/r2e_fs_ncu.log 2>&1
(bash tools/variants_plan.sh "" "21:8,7,6;24:8,8,8;22:8,7,7" "21:6,7,8;24:9,8,7;22:7,7,8" "21:9,6,6;24:7,8,9;22:9,7,6" "21:6,6,9;24:6,9,9;22:6,8,8" 2>&1) > gpurun_out/r2e_plans.txt
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/r2e_pytest.log
tail -2 gpurun_out/r2e_pytest.log; cat gpurun_out/r2e_fs_emulated.txt
