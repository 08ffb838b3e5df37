N=2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== check_fused_allreduce (tagged)"; timeout 300 $TR --master-port 29511 tools/check_fused_allreduce.py > gpurun_out/r2_check_fused_n$N.log 2>&1; echo "rc=$?"; grep -v "Missing units\|OMP_NUM\|\*\*\*" gpurun_out/r2_check_fused_n$N.log | tail -15
echo "== pytest two-GPU test"; python -m pytest tests/test_sharded_gloo.py -m gpu -x -q 2>&1 | tail -3
echo "== bench fused quick"; timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 30 --warmup 3 > gpurun_out/r2_bench_n${N}_fused.json 2> gpurun_out/r2_bench_n${N}_fused.err; echo "rc=$?"; tail -2 gpurun_out/r2_bench_n${N}_fused.err; python tools/bench_digest.py gpurun_out/r2_bench_n${N}_fused.json
echo "== exchange timeline (tagged)"; MCD_B200_LIB=scratch_ab/profile/libmcd_b200.so timeout 300 $TR --master-port 29514 tools/probe/exchange_timeline.py > gpurun_out/r2_exchange_timeline_n$N.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2_exchange_timeline_n$N.log | cut -c1-300
