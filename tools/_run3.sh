mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2c_env.txt
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2c_bench2.json 2> gpurun_out/r2c_bench2.err); echo "bench2 rc=$?"
tail -5 gpurun_out/r2c_bench2.err
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 tools/multi_gpu_check.py --full --cols 16 > gpurun_out/r2c_check2.txt 2>&1); echo "check rc=$?"; tail -5 gpurun_out/r2c_check2.txt
