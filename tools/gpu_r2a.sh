# round 2, GPU call A: tests, baseline timings of the mid-size configurations, launch timelines
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
python tools/probe/single_stars_timing.py > gpurun_out/r2a_single_stars.log 2>&1; cat gpurun_out/r2a_single_stars.log
python tools/ab_configs.py c1 c2 c3 c3b c4 mix mixgb > gpurun_out/r2a_ab_base.log 2>&1; cat gpurun_out/r2a_ab_base.log
bash tools/probe/launch_timeline.sh c2 c3 c4 > gpurun_out/r2a_timeline.log 2>&1; cat gpurun_out/r2a_timeline.log
for c in c3 c4; do for g in 96,1 128,1 176,1 256,1 96,2 128,2 176,2 256,2 128,3 256,3 128,4 256,4; do
  echo -n "$c $g : "; MCD_GEOMETRY=$g python tools/ab_configs.py $c --calls 300 2>/dev/null | sed 's/.*| device *\([0-9.]* us\/call\).*\(grid [0-9x ]*\),.*/\1 \2/'
done; done > gpurun_out/r2a_geometry.log 2>&1; cat gpurun_out/r2a_geometry.log
