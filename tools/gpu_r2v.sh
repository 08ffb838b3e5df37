mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | grep -v "Missing units" | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | grep -v "Missing units" | tail -4
python tools/probe/host_call_split.py 2>&1 | grep -v "Missing units" > gpurun_out/r2v_host_split.log; cat gpurun_out/r2v_host_split.log
MCD_HOST_CALL=graph python tools/probe/host_call_split.py 2>&1 | grep -v "Missing units" > gpurun_out/r2v_host_split_graph.log; cat gpurun_out/r2v_host_split_graph.log
python tools/ab_configs.py c5 c3 2>&1 | grep -v "Missing units" | cut -c1-200
