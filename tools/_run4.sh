mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2d_env.txt
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2d_bench8.json 2> gpurun_out/r2d_bench8.err); echo "bench8 rc=$?"
tail -5 gpurun_out/r2d_bench8.err
