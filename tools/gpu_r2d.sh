mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2d_pytest.log
grep -v "Missing units" gpurun_out/r2d_pytest.log | tail -30
python tools/probe/sampler_c5.py 2>&1 | grep -v "Missing units" > gpurun_out/r2d_sampler_c5.log; cat gpurun_out/r2d_sampler_c5.log
