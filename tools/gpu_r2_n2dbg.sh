mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
N_STARS=100000 timeout 300 $TR --master-port 29511 tools/check_fused_allreduce.py > gpurun_out/r2_check_fused_small.log 2>&1; echo "rc=$?"; grep -v "Missing units\|OMP_NUM\|\*\*\*" gpurun_out/r2_check_fused_small.log | tail -40
python -m pytest tests/test_sharded_gloo.py -m gpu -x -q 2>&1 | tail -40
