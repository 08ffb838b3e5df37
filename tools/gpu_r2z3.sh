mkdir -p gpurun_out
timeout 26 python tools/probe/curve_kernels.py 2>&1 | grep -v "Missing units" | tee gpurun_out/r2z_curve_kernels.log
