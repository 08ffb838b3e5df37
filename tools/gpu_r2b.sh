# round 2, GPU call B: tests with the new mixture fast path, A/B of the mixture builds, geometry check, bench smoke
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2b_pytest.log
grep -v "Missing units" gpurun_out/r2b_pytest.log | tail -30
( echo "== shipped (fast path, full precision, 1 pair)"; python tools/ab_configs.py c3 c3b c4 mix mixgb
for v in pairs2 lean leanp2 leanp2b3; do echo "== $v"; MCD_B200_LIB=scratch_ab/$v/libmcd_b200.so python tools/ab_configs.py c3 c3b mix mixgb; done
echo "== newton2 (headline)"; python tools/ab_configs.py c5 c2; MCD_B200_LIB=scratch_ab/newton2/libmcd_b200.so python tools/ab_configs.py c5 c2 ) 2>&1 | grep -v "Missing units" > gpurun_out/r2b_ab.log; cat gpurun_out/r2b_ab.log
for c in c3 c4; do for g in 176,1 256,1 176,2 256,2 128,3 176,3 256,3 128,4 256,4; do
  echo -n "$c $g : "; MCD_GEOMETRY=$g python tools/ab_configs.py $c --calls 300 2>/dev/null | sed 's/.*| device *\([0-9.]* us\/call\).*\(grid [0-9x ]*\),.*/\1 \2/'
done; done > gpurun_out/r2b_geometry.log 2>&1; cat gpurun_out/r2b_geometry.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2b_bench.err; head -c 3000 gpurun_out/r2b_bench.json
