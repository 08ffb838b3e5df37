#!/bin/bash
# FP64-pipe counters of the likelihood kernel under the conditions of an N-GPU run, without a collective inside the
# profiled process: GPU 0 runs one star shard of C5 (1.25e6 stars x 512 walkers, what each rank runs at N = 8)
# under ncu, GPUs 1..N-1 run the same shard workload back to back meanwhile (power / thermal state of a busy node).
#   bash tools/ncu_shard_under_load.sh N OUT.csv
N=${1:-8}
OUT=${2:-gpurun_out/r02_lnlike_shard_n${N}_ncu.csv}
PIDS=""
for r in $(seq 1 $((N - 1))); do
  CUDA_VISIBLE_DEVICES=$r timeout 120 python tools/probe/ncu_targets.py c5s 100000 > /dev/null 2>&1 &
  PIDS="$PIDS $!"
done
sleep 12      # the background ranks have built their catalogue and are launching
CUDA_VISIBLE_DEVICES=0 timeout 120 ncu --metrics sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg.per_second \
  --clock-control none -k regex:lnlike_kernel -s 4 -c 6 --csv --log-file "$OUT" python tools/probe/ncu_targets.py c5s 12
echo "ncu rc=$?"
nvidia-smi --query-gpu=index,utilization.gpu,clocks.sm,power.draw --format=csv,noheader > "${OUT%.csv}_smi.txt" 2>&1
for p in $PIDS; do kill $p 2>/dev/null; done
wait
