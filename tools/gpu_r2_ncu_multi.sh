# round 2: bench line and hardware counters at N GPUs (argument 1): (B) one star shard under ncu while the other GPUs run the same shard workload.
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== bench fused"
timeout 300 $TR --master-port 29512 bench.py --gpus $N --steps 60 --warmup 3 > gpurun_out/r2q_bench_n${N}_fused.json 2> gpurun_out/r2q_bench_n${N}_fused.err; echo "rc=$?"; tail -2 gpurun_out/r2q_bench_n${N}_fused.err; python tools/bench_digest.py gpurun_out/r2q_bench_n${N}_fused.json
echo "== (B) shard under load"
bash tools/ncu_shard_under_load.sh $N gpurun_out/r02_lnlike_shard_n${N}_ncu.csv 2>&1 | tail -3
grep -c lnlike_kernel gpurun_out/r02_lnlike_shard_n${N}_ncu.csv
# (A) rank 0 of the bench under ncu with the NCCL collective (tools/ncu_rank0.sh) stalls after NCCL's start-up at
# N = 2 (twice, killed by timeout: gpurun_out/r2_ncu_n2.log, r2q_ncu_n2.log), so the counters come from (B).
