mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | grep -v "Missing units" | tail -3
python tools/ab_configs.py c3 c3b c4 c5s c2 c1 2>&1 | grep -v "Missing units" | cut -c1-150 | tee gpurun_out/r2x_ab.log
