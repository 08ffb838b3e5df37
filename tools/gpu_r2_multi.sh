# round 2, multi-GPU call: N = number of GPUs of the box (argument 1); "quick" as argument 2 = one bench line only
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== bench fused (tagged-word exchange)"
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 60 --warmup 3 > gpurun_out/r2_bench_n${N}_fused.json 2> gpurun_out/r2_bench_n${N}_fused.err; echo "rc=$?"; tail -2 gpurun_out/r2_bench_n${N}_fused.err; python tools/bench_digest.py gpurun_out/r2_bench_n${N}_fused.json
if [ "$2" = "quick" ]; then exit 0; fi
echo "== check_fused_allreduce ($N ranks, tagged-word exchange)"
timeout 300 $TR --master-port 29511 tools/check_fused_allreduce.py > gpurun_out/r2_check_fused_n$N.log 2>&1; echo "rc=$?"; grep -v "Missing units\|OMP_NUM\|\*\*\*" gpurun_out/r2_check_fused_n$N.log | tail -15
echo "== check_fused_allreduce ($N ranks, flag + fence exchange)"
MCD_XCHG=flags timeout 300 $TR --master-port 29516 tools/check_fused_allreduce.py > gpurun_out/r2_check_fused_flags_n$N.log 2>&1; echo "rc=$?"; grep -v "Missing units\|OMP_NUM\|\*\*\*" gpurun_out/r2_check_fused_flags_n$N.log | tail -5
echo "== bench fused (flag + fence exchange of round 1)"
MCD_XCHG=flags timeout 600 $TR --master-port 29517 bench.py --gpus $N --steps 60 --warmup 3 --no-samplers > gpurun_out/r2_bench_n${N}_fused_flags.json 2> gpurun_out/r2_bench_n${N}_fused_flags.err; echo "rc=$?"; python tools/bench_digest.py gpurun_out/r2_bench_n${N}_fused_flags.json
echo "== bench nccl"
MCD_COLLECTIVE=nccl timeout 600 $TR --master-port 29513 bench.py --gpus $N --steps 60 --warmup 3 --no-samplers > gpurun_out/r2_bench_n${N}_nccl.json 2> gpurun_out/r2_bench_n${N}_nccl.err; echo "rc=$?"; python tools/bench_digest.py gpurun_out/r2_bench_n${N}_nccl.json
echo "== exchange timeline (tagged, then flags)"
MCD_B200_LIB=scratch_ab/profile/libmcd_b200.so timeout 300 $TR --master-port 29514 tools/probe/exchange_timeline.py > gpurun_out/r2_exchange_timeline_n$N.log 2>&1; echo "rc=$?"
MCD_XCHG=flags MCD_B200_LIB=scratch_ab/profile/libmcd_b200.so timeout 300 $TR --master-port 29518 tools/probe/exchange_timeline.py > gpurun_out/r2_exchange_timeline_flags_n$N.log 2>&1; echo "rc=$?"
grep -c "^launch" gpurun_out/r2_exchange_timeline_n$N.log gpurun_out/r2_exchange_timeline_flags_n$N.log
