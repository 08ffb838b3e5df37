mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2e_pytest.log
grep -v "Missing units" gpurun_out/r2e_pytest.log | tail -20
python tools/probe/host_call_split.py 2>&1 | grep -v "Missing units" > gpurun_out/r2e_host_split.log; cat gpurun_out/r2e_host_split.log
MCD_HOST_CALL=graph python tools/probe/host_call_split.py 2>&1 | grep -v "Missing units" > gpurun_out/r2e_host_split_graph.log; cat gpurun_out/r2e_host_split_graph.log
python bench.py --steps 30 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2e_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r2e_bench.json').read()); print(d['value'], d['e2e']['value'], d['steps_per_s'])
for k,v in d['configs'].items(): print(k, '%.2e'%v['max_rel_err_vs_oracle'], '%.3e'%v['terms_per_s'], '%.1f us'%v['us_per_call'], '%.3e'%v['e2e_terms_per_s'], '%.1f us'%v['e2e_us_per_call'], v['steps_per_s'])
"
