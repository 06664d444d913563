#!/bin/bash
# builds and runs the pipe micro-benchmarks on the GPU box (nvcc is there): tools/ubench/run.sh > gpurun_out/ubench.txt
cd "$(dirname "$0")" && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o pipes pipes.cu && ./pipes
