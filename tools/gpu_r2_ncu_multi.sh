# round 2: counters at N GPUs (argument 1).  (A) rank 0 of the bench under ncu, NCCL collective; (B) shard under load.
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== bench fused"
timeout 300 $TR --master-port 29512 bench.py --gpus $N --steps 60 --warmup 3 > gpurun_out/r2q_bench_n${N}_fused.json 2> gpurun_out/r2q_bench_n${N}_fused.err; echo "rc=$?"; tail -2 gpurun_out/r2q_bench_n${N}_fused.err; python tools/bench_digest.py gpurun_out/r2q_bench_n${N}_fused.json
echo "== (B) shard under load"
bash tools/ncu_shard_under_load.sh $N gpurun_out/r02_lnlike_shard_n${N}_ncu.csv 2>&1 | tail -3
grep -c lnlike_kernel gpurun_out/r02_lnlike_shard_n${N}_ncu.csv
echo "== (A) rank 0 of the bench under ncu (NCCL collective)"
MCD_COLLECTIVE=nccl NCU_OUT=gpurun_out/r02_lnlike_n${N}_ncu.csv timeout -k 10 150 $TR --master-port 29519 --no-python bash tools/ncu_rank0.sh --gpus $N --steps 12 --warmup 3 --no-samplers --no-configs --no-cpu-baseline > gpurun_out/r2q_ncu_n${N}.log 2>&1; echo "rc=$?"
grep -c lnlike_kernel gpurun_out/r02_lnlike_n${N}_ncu.csv; tail -3 gpurun_out/r2q_ncu_n${N}.log | cut -c1-200
