mkdir -p gpurun_out
( echo "== shipped (3 CTAs/SM, two pairs)"; python tools/ab_configs.py c5 c5s
for v in mb2 mb4 p1mb4; do echo "== $v"; MCD_B200_LIB=scratch_ab/$v/libmcd_b200.so python tools/ab_configs.py c5 c5s; done ) 2>&1 | grep -v "Missing units" | cut -c1-130 | tee gpurun_out/r2y_ab.log
