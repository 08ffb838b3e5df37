mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2m_pytest.log
grep -v "Missing units" gpurun_out/r2m_pytest.log | tail -12
( echo "== shipped (variance scale 4: rsqrt_twice everywhere in the mixture term)"; python tools/ab_configs.py c5 c4 c3 c3b mix mixgb c1 c2
) 2>&1 | grep -v "Missing units" | cut -c1-150 > gpurun_out/r2m_ab.log; cat gpurun_out/r2m_ab.log
