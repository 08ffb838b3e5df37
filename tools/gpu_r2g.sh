mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2g_pytest.log
grep -v "Missing units" gpurun_out/r2g_pytest.log | tail -12
echo "== sampler timing C5"; MCD_TIMING=1 python tools/probe/sampler_c5.py 2>&1 | grep -v "Missing units" | grep -v "^mcd_ensemble_run" > gpurun_out/r2g_sampler_c5.log; cat gpurun_out/r2g_sampler_c5.log
echo "== shard-sized workload geometry"; ( python tools/ab_configs.py c5s; for g in 256,1 256,2 256,3 256,4 256,6 256,8 256,11 ; do echo -n "$g : "; MCD_GEOMETRY=$g python tools/ab_configs.py c5s 2>/dev/null | sed 's/.*| device *\([0-9.]* us\/call\).*\(grid [0-9x ]*\),.*/\1 \2/'; done ) 2>&1 | grep -v "Missing units" | cut -c1-200 > gpurun_out/r2g_c5s_geometry.log; cat gpurun_out/r2g_c5s_geometry.log
echo "== bench"; python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2g_bench.json').read()); print(d['value'], d['e2e']['value'], d['steps_per_s'], d['roofline']['frac'], d['roofline']['fp64_pipe_frac'])
for k,v in d['configs'].items(): print(k, '%.2e'%v['max_rel_err_vs_oracle'], '%.3e'%v['terms_per_s'], '%.1f us'%v['us_per_call'], '%.3e'%v['e2e_terms_per_s'], '%.1f us'%v['e2e_us_per_call'], v['steps_per_s'])
"
echo "== ncu launch list of the bench command"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 20 --no-configs --no-cpu-baseline --no-samplers > gpurun_out/r2g_ncu_list.log 2>&1; echo "rc=$?"
echo "== ncu full: headline"
ncu --set full --clock-control none --import-source on -k regex:lnlike_kernel -s 20 -c 2 -o gpurun_out/r02_prof_lnlike -f python bench.py --steps 20 --no-configs --no-cpu-baseline --no-samplers > gpurun_out/r2g_ncu_full.log 2>&1; echo "rc=$?"
for t in mix mixgb c5s; do echo "== ncu full: $t"; python tools/probe/ncu_targets.py $t > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:lnlike_kernel -s 3 -c 2 -o gpurun_out/r02_prof_$t -f python tools/probe/ncu_targets.py $t > gpurun_out/r2g_ncu_$t.log 2>&1; echo "rc=$?"; done
echo "== ncu full: single_stars"; python tools/probe/ncu_targets.py single_stars > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:single_stars_kernel -s 2 -c 2 -o gpurun_out/r02_prof_single_stars -f python tools/probe/ncu_targets.py single_stars > gpurun_out/r2g_ncu_ss.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep
