# round 2, GPU call C: tests with the new defaults, configs, host-call split, bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2c_pytest.log
grep -v "Missing units" gpurun_out/r2c_pytest.log | tail -30
python tools/ab_configs.py c1 c2 c3 c3b c4 mix mixgb c5 2>&1 | grep -v "Missing units" > gpurun_out/r2c_ab.log; cat gpurun_out/r2c_ab.log
python tools/probe/host_call_split.py 2>&1 | grep -v "Missing units" > gpurun_out/r2c_host_split.log; cat gpurun_out/r2c_host_split.log
python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2c_bench.err; head -c 1500 gpurun_out/r2c_bench.json
