mkdir -p gpurun_out
echo "== sampler timing C5"; MCD_TIMING=1 python tools/probe/sampler_c5.py 2>&1 | grep -v "Missing units" > gpurun_out/r2f_sampler_c5.log; cat gpurun_out/r2f_sampler_c5.log
echo "== bench (samplers, timing)"; MCD_TIMING=1 python bench.py --steps 60 --no-configs --no-cpu-baseline > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; grep mcd_ensemble_run gpurun_out/r2f_bench.err; python -c "
import json; d=json.loads(open('gpurun_out/r2f_bench.json').read()); print(d['value'], d['e2e']['value'], d['steps_per_s'])"
echo "== headline tuning variants"
( python tools/ab_configs.py c5 c4 c2; for v in mb4 mb2 p1mb4; do echo "-- $v"; MCD_B200_LIB=scratch_ab/$v/libmcd_b200.so python tools/ab_configs.py c5 c4 c2; done; echo "-- bgp2"; MCD_B200_LIB=scratch_ab/bgp2/libmcd_b200.so python tools/ab_configs.py c3 c3b mix mixgb ) 2>&1 | grep -v "Missing units" > gpurun_out/r2f_ab.log; cat gpurun_out/r2f_ab.log | cut -c1-190
echo "== shard-sized workload (what each of 8 ranks runs)"; python tools/ab_configs.py c5 2>/dev/null | cut -c1-190
