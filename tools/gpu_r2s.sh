mkdir -p gpurun_out
( echo "== shipped (geometry model: remainder rounds, imbalance 2 % / waves^1.5)"; python tools/ab_configs.py c5 c5h c5q c5s mix mixgb c3 c3b c4 c2 c1
echo "-- mix one wave (208,65)"; MCD_GEOMETRY=208,65 python tools/ab_configs.py mix mixgb
echo "-- mix 256,14"; MCD_GEOMETRY=256,14 python tools/ab_configs.py mix mixgb
) 2>&1 | grep -v "Missing units" | cut -c1-20,70-140,180-215 > gpurun_out/r2s_geometry.log; cat gpurun_out/r2s_geometry.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
