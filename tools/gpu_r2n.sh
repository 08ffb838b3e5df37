mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2n_pytest.log
grep -v "Missing units" gpurun_out/r2n_pytest.log | tail -12
( echo "== shipped (light renormalisation, exponent folded every four stars)"; python tools/ab_configs.py c5 c4 c3 c3b mix mixgb c1 c2
echo "== g2 (light renormalisation, every two stars in the no-background variants)"; MCD_B200_LIB=scratch_ab/g2/libmcd_b200.so python tools/ab_configs.py c5 c4 c2
) 2>&1 | grep -v "Missing units" | cut -c1-150 > gpurun_out/r2n_ab.log; cat gpurun_out/r2n_ab.log
