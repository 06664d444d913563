#!/bin/bash
# NVLink byte counters + durations of the two exchange kernels of the peer-memory four-step LDE on 2 GPUs:
# rank 0 under ncu, rank 1 plain; rendezvous through files (tools/prof_fourstep.py).
box=$(mktemp -d /tmp/fsprof.XXXXXX)
python tools/prof_fourstep.py 1 2 "$box" "${1:-26}" > gpurun_out/r2_nvlink_rank1.log 2>&1 &
ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -k regex:fourstep_ --csv --log-file gpurun_out/r2_nvlink.csv python tools/prof_fourstep.py 0 2 "$box" "${1:-26}"
wait
rm -rf "$box"
