CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 46 -c 8 -o gpurun_out/prof_gemm_v2 $CMD > gpurun_out/ncu2.log 2>&1
echo full_exit=$?
