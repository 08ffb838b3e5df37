mkdir -p gpurun_out
( for cfg in c5s c5h; do
echo "== $cfg default geometry"; python tools/ab_configs.py $cfg
for g in 256,44 256,22 256,11 256,8 256,6 256,4 256,3 256,2 192,8 128,8; do echo "-- MCD_GEOMETRY=$g"; MCD_GEOMETRY=$g python tools/ab_configs.py $cfg; done
done ) 2>&1 | grep -v "Missing units" | cut -c1-20,70-140,210-260 > gpurun_out/r2r_geometry.log; cat gpurun_out/r2r_geometry.log
