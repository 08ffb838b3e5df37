mkdir -p gpurun_out
( echo "== shipped"; python tools/ab_configs.py c5 c4 c2 c1
for v in ptr1 ptr2; do echo "== $v"; MCD_B200_LIB=scratch_ab/$v/libmcd_b200.so python tools/ab_configs.py c5 c4 c2 c1; done
echo "== b3p1 (mixtures: one pair per iteration, three CTAs per SM)"; MCD_B200_LIB=scratch_ab/b3p1/libmcd_b200.so python tools/ab_configs.py c3 c3b mix mixgb
) 2>&1 | grep -v "Missing units" | cut -c1-150 > gpurun_out/r2p_ab.log; cat gpurun_out/r2p_ab.log
