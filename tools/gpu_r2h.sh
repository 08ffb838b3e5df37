mkdir -p gpurun_out
echo "== shard-sized workload with the refitted geometry model"; python tools/ab_configs.py c5s c5 c4 c3 mix 2>&1 | grep -v "Missing units" | cut -c1-200 | tee gpurun_out/r2h_ab.log
python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"
echo "== ncu launch list of the bench command"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 20 --no-configs --no-cpu-baseline --no-samplers > gpurun_out/r2h_ncu_list.log 2>&1; echo "rc=$?"
echo "== ncu full: headline"
ncu --set full --clock-control none -k regex:lnlike_kernel -s 20 -c 2 -o /tmp/r02_prof_lnlike -f python bench.py --steps 20 --no-configs --no-cpu-baseline --no-samplers > gpurun_out/r2h_ncu_full.log 2>&1; echo "rc=$?"
python tools/ncu_summary.py /tmp/r02_prof_lnlike.ncu-rep gpurun_out/r02_lnlike_kernel_ncu_metrics.csv --traffic 'lnlike<RADIAL,FIXED,BG_NONE,FAST>' 10000000 512
cp profiles/r02_ncu_traffic.json gpurun_out/ 2>/dev/null
for t in mix mixgb c5s c3; do echo "== ncu full: $t"; python tools/probe/ncu_targets.py $t > /dev/null 2>&1 && ncu --set full --clock-control none -k regex:lnlike_kernel -s 3 -c 2 -o /tmp/r02_prof_$t -f python tools/probe/ncu_targets.py $t > gpurun_out/r2h_ncu_$t.log 2>&1; echo "rc=$?"; python tools/ncu_summary.py /tmp/r02_prof_$t.ncu-rep gpurun_out/r02_${t}_kernel_ncu_metrics.csv; done
echo "== ncu full: single_stars"; python tools/probe/ncu_targets.py single_stars > /dev/null 2>&1 && ncu --set full --clock-control none -k regex:single_stars_kernel -s 2 -c 2 -o /tmp/r02_prof_single_stars -f python tools/probe/ncu_targets.py single_stars > gpurun_out/r2h_ncu_ss.log 2>&1; echo "rc=$?"; python tools/ncu_summary.py /tmp/r02_prof_single_stars.ncu-rep gpurun_out/r02_single_stars_kernel_ncu_metrics.csv --kernel single_stars_kernel
cp /tmp/r02_prof_lnlike.ncu-rep gpurun_out/ ; ls -la gpurun_out; du -sh gpurun_out
