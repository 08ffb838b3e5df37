# exchange timeline of the fused cross-GPU call with per-rank logs (argument 1: number of GPUs)
N=${1:-4}
mkdir -p gpurun_out/xtl_n$N
MCD_B200_LIB=scratch_ab/profile/libmcd_b200.so timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 \
  --redirects 3 --log-dir gpurun_out/xtl_n$N tools/probe/exchange_timeline.py > gpurun_out/xtl_n$N/driver.log 2>&1; echo "rc=$?"
find gpurun_out/xtl_n$N -name "stdout.log" | sort | while read f; do r=$(echo $f | grep -o "[0-9]*/stdout.log" | cut -d/ -f1); grep -v "Missing units" $f | tail -6 > gpurun_out/xtl_n$N/rank$r.log; done
ls gpurun_out/xtl_n$N; tail -4 gpurun_out/xtl_n$N/rank0.log | cut -c1-600
