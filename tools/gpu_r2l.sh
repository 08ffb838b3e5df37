mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2l_pytest.log
grep -v "Missing units" gpurun_out/r2l_pytest.log | tail -12
( echo "== shipped (rsqrt_twice for D2, late flag check)"; python tools/ab_configs.py c5 c4 c3 c3b mix mixgb
echo "== a0b0 (neither)"; MCD_B200_LIB=scratch_ab/a0b0/libmcd_b200.so python tools/ab_configs.py c5 c4 c3 c3b mix mixgb
echo "== a1b0 (rsqrt_twice only)"; MCD_B200_LIB=scratch_ab/a1b0/libmcd_b200.so python tools/ab_configs.py c5 c3 mix
echo "== a0b1 (late flag only)"; MCD_B200_LIB=scratch_ab/a0b1/libmcd_b200.so python tools/ab_configs.py c3 mix
) 2>&1 | grep -v "Missing units" | cut -c1-150 > gpurun_out/r2l_ab.log; cat gpurun_out/r2l_ab.log
