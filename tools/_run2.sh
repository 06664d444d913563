mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/r2b_pytest.log
tail -3 gpurun_out/r2b_pytest.log
(timeout 600 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err); echo "bench rc=$?"; tail -5 gpurun_out/r2b_bench.err
(timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2b_bench_ref.json 2> gpurun_out/r2b_bench_ref.err); echo "ref rc=$?"
