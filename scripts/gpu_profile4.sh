CMD="python scripts/gemm_sweep.py 32768"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 277 -c 1 -o gpurun_out/prof_hidden $CMD > gpurun_out/ncu4.log 2>&1
echo full_exit=$?
