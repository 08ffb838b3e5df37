mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2k_pytest.log
grep -v "Missing units" gpurun_out/r2k_pytest.log | tail -12
( echo "== shipped (lean2: 1024-entry table, quadratic)"; python tools/ab_configs.py c3 c3b mix mixgb
for v in lean1; do echo "== $v"; MCD_B200_LIB=scratch_ab/$v/libmcd_b200.so python tools/ab_configs.py c3 c3b mix mixgb; done ) 2>&1 | grep -v "Missing units" | cut -c1-210 > gpurun_out/r2k_ab.log; cat gpurun_out/r2k_ab.log
