mkdir -p gpurun_out
python -m pytest tests/test_gpu_curves.py tests/test_gpu_golden.py -m gpu -q 2>&1 | grep -v "Missing units" | tail -25 | tee gpurun_out/r2z2_pytest.log
