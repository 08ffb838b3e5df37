# round 2, multi-GPU call: N = number of GPUs of the box (argument 1)
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== check_fused_allreduce ($N ranks)"
timeout 300 $TR --master-port 29511 tools/check_fused_allreduce.py > gpurun_out/r2_check_fused_n$N.log 2>&1; echo "rc=$?"; grep -v "Missing units" gpurun_out/r2_check_fused_n$N.log | tail -16
echo "== bench fused"
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 40 --warmup 3 > gpurun_out/r2_bench_n${N}_fused.json 2> gpurun_out/r2_bench_n${N}_fused.err; echo "rc=$?"; tail -3 gpurun_out/r2_bench_n${N}_fused.err; head -c 1800 gpurun_out/r2_bench_n${N}_fused.json; echo
echo "== bench nccl"
MCD_COLLECTIVE=nccl timeout 600 $TR --master-port 29513 bench.py --gpus $N --steps 40 --warmup 3 > gpurun_out/r2_bench_n${N}_nccl.json 2> gpurun_out/r2_bench_n${N}_nccl.err; echo "rc=$?"; tail -3 gpurun_out/r2_bench_n${N}_nccl.err; head -c 600 gpurun_out/r2_bench_n${N}_nccl.json; echo
echo "== exchange timeline"
MCD_B200_LIB=scratch_ab/profile/libmcd_b200.so timeout 300 $TR --master-port 29514 tools/probe/exchange_timeline.py > gpurun_out/r2_exchange_timeline_n$N.log 2>&1; echo "rc=$?"; grep -A1 "^launch" gpurun_out/r2_exchange_timeline_n$N.log | tail -24
echo "== ncu on rank 0 (nccl mode; after the same command ran plainly above)"
MCD_COLLECTIVE=nccl NCU_OUT=gpurun_out/r02_lnlike_n${N}_ncu.csv timeout 600 $TR --master-port 29515 --no-python bash tools/ncu_rank0.sh --gpus $N --steps 6 --warmup 3 --no-samplers > gpurun_out/r2_ncu_n${N}.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2_ncu_n${N}.log; tail -8 gpurun_out/r02_lnlike_n${N}_ncu.csv | cut -c1-300
