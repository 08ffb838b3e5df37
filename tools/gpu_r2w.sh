mkdir -p gpurun_out
python -m pytest tests/test_sharded_gloo.py -m gpu -x -q 2>&1 | grep -v "Missing units" | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29512 bench.py --gpus 2 --steps 60 --warmup 3 > gpurun_out/r02f_bench_n2.json 2> gpurun_out/r02f_bench_n2.err; echo "rc=$?"; python tools/bench_digest.py gpurun_out/r02f_bench_n2.json 2>/dev/null | cut -c1-330
