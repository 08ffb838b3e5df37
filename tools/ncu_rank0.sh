#!/bin/bash
# Launched by torch.distributed.run --no-python: rank 0 runs bench.py under ncu (hardware counters of the likelihood
# kernel only, a handful of metrics = one or two replays of a stand-alone kernel), the other ranks run it plainly.
# Only with MCD_COLLECTIVE=nccl: the fused kernel waits for its peers inside the launch and must not be replayed.
OUT=${NCU_OUT:-gpurun_out/r02_lnlike_n${WORLD_SIZE}_ncu.csv}
if [ "$RANK" = "0" ]; then
  exec ncu --metrics sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,nvlrx__bytes.sum,nvltx__bytes.sum \
    --clock-control none -k regex:lnlike_kernel -s 12 -c 6 --csv --log-file "$OUT" python bench.py "$@"
else
  exec python bench.py "$@"
fi
