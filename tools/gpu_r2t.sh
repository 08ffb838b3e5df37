mkdir -p gpurun_out
python tools/probe/ncu_targets.py mix > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:lnlike_kernel -s 3 -c 1 -o gpurun_out/r02b_prof_mix -f python tools/probe/ncu_targets.py mix > gpurun_out/r2t_ncu_mix.log 2>&1; echo "rc=$?"
ls -la gpurun_out/r02b_prof_mix.ncu-rep
