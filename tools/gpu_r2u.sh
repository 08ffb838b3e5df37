mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
( echo "== shipped (polynomial coefficients as literals)"; python tools/ab_configs.py c3 c3b mix mixgb ) 2>&1 | grep -v "Missing units" | cut -c1-150 > gpurun_out/r2u_ab.log; cat gpurun_out/r2u_ab.log
