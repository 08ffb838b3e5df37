mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2j_pytest.log
grep -v "Missing units" gpurun_out/r2j_pytest.log | tail -12
python tools/probe/host_call_split.py 2>&1 | grep -v "Missing units" > gpurun_out/r2j_host_split.log; cat gpurun_out/r2j_host_split.log
MCD_HOST_CALL=graph python tools/probe/host_call_split.py 2>&1 | grep -v "Missing units" > gpurun_out/r2j_host_split_graph.log; cat gpurun_out/r2j_host_split_graph.log
MCD_HOST_CALL=sync python tools/probe/host_call_split.py 2>&1 | grep -v "Missing units" > gpurun_out/r2j_host_split_sync.log; cat gpurun_out/r2j_host_split_sync.log
python bench.py --steps 30 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2j_bench.err
python tools/bench_digest.py gpurun_out/r2j_bench.json
