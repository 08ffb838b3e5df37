mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2o_pytest.log
grep -v "Missing units" gpurun_out/r2o_pytest.log | tail -30
( echo "== shipped"; python tools/ab_configs.py c5 c4 c3 c3b mix mixgb c1 c2
) 2>&1 | grep -v "Missing units" | cut -c1-150 > gpurun_out/r2o_ab.log; cat gpurun_out/r2o_ab.log
python bench.py --steps 30 > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2o_bench.err
python tools/bench_digest.py gpurun_out/r2o_bench.json
