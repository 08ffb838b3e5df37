# round 2, final single-GPU record: tests, the default bench line, ncu launch list of the bench command, ncu --set full
# of every kernel family (summaries only go to profiles/), SASS counts are made on the build box.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02f_pytest.log; grep -v "Missing units" gpurun_out/r02f_pytest.log | tail -4
python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench rc=$?"; python tools/bench_digest.py gpurun_out/r02f_bench.json
python bench.py --impl reference > gpurun_out/r02f_bench_reference.json 2> gpurun_out/r02f_bench_reference.err; echo "reference arm rc=$?"; cut -c1-400 gpurun_out/r02f_bench_reference.json
echo "== A/B table of the shipped build"; python tools/ab_configs.py c5 c5free c5s c4 c3 c3b mix mixgb c2 c1 2>&1 | grep -v "Missing units" | cut -c1-230 | tee gpurun_out/r02f_ab.log
echo "== ncu launch list of the bench command"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 20 --no-configs --no-cpu-baseline --no-samplers > gpurun_out/r02f_ncu_list.log 2>&1; echo "rc=$?"
echo "== ncu full: headline"
ncu --set full --clock-control none --import-source on -k regex:lnlike_kernel -s 20 -c 2 -o /tmp/r02_prof_lnlike -f python bench.py --steps 20 --no-configs --no-cpu-baseline --no-samplers > gpurun_out/r02f_ncu_full.log 2>&1; echo "rc=$?"
python tools/ncu_summary.py /tmp/r02_prof_lnlike.ncu-rep gpurun_out/r02_lnlike_kernel_ncu_metrics.csv --traffic 'lnlike<RADIAL,FIXED,BG_NONE,FAST>' 10000000 512
cp profiles/r02_ncu_traffic.json gpurun_out/ 2>/dev/null
for t in mix mixgb c5s c3 c4; do echo "== ncu full: $t"; python tools/probe/ncu_targets.py $t > /dev/null 2>&1 && ncu --set full --clock-control none -k regex:lnlike_kernel -s 3 -c 2 -o /tmp/r02_prof_$t -f python tools/probe/ncu_targets.py $t > gpurun_out/r02f_ncu_$t.log 2>&1; echo "rc=$?"; python tools/ncu_summary.py /tmp/r02_prof_$t.ncu-rep gpurun_out/r02_${t}_kernel_ncu_metrics.csv; done
echo "== ncu full: single_stars"; python tools/probe/ncu_targets.py single_stars > /dev/null 2>&1 && ncu --set full --clock-control none -k regex:single_stars_kernel -s 2 -c 2 -o /tmp/r02_prof_single_stars -f python tools/probe/ncu_targets.py single_stars > gpurun_out/r02f_ncu_ss.log 2>&1; echo "rc=$?"; python tools/ncu_summary.py /tmp/r02_prof_single_stars.ncu-rep gpurun_out/r02_single_stars_kernel_ncu_metrics.csv --kernel single_stars_kernel
cp /tmp/r02_prof_lnlike.ncu-rep gpurun_out/r02_prof_lnlike.ncu-rep; ls -la gpurun_out/*.ncu-rep | tail -3
