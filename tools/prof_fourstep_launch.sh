#!/bin/bash
# per-rank launcher for torchrun: rank 0 under ncu (NVLink byte counters + durations of the two exchange kernels), the rest plain
if [ "$RANK" = "0" ]; then
  exec ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
       --clock-control none -k regex:fourstep_ --csv --log-file gpurun_out/r2_nvlink.csv python tools/prof_fourstep.py "$@"
else
  exec python tools/prof_fourstep.py "$@"
fi
