# round 2, final multi-GPU record: N = number of GPUs of the box (argument 1); "check" as argument 2 adds the fused-exchange checker
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== bench fused"
timeout 300 $TR --master-port 29512 bench.py --gpus $N --steps 60 --warmup 3 > gpurun_out/r02f_bench_n${N}.json 2> gpurun_out/r02f_bench_n${N}.err; echo "rc=$?"; python tools/bench_digest.py gpurun_out/r02f_bench_n${N}.json
if [ "$2" = "check" ]; then
echo "== check_fused_allreduce ($N ranks)"
timeout 200 $TR --master-port 29511 tools/check_fused_allreduce.py > gpurun_out/r02f_check_fused_n$N.log 2>&1; echo "rc=$?"; grep -v "Missing units\|OMP_NUM\|\*\*\*" gpurun_out/r02f_check_fused_n$N.log | tail -14
echo "== bench nccl"
MCD_COLLECTIVE=nccl timeout 300 $TR --master-port 29513 bench.py --gpus $N --steps 60 --warmup 3 --no-samplers > gpurun_out/r02f_bench_n${N}_nccl.json 2> gpurun_out/r02f_bench_n${N}_nccl.err; echo "rc=$?"; python tools/bench_digest.py gpurun_out/r02f_bench_n${N}_nccl.json
fi
