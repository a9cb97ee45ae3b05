# launch list (per-kernel device time) of one K1 rollout + AC step at the sweep size, then one
# full ncu capture of the GRU contraction kernel.  Plain run first (exit code gates the ncu runs).
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
echo launches_exit=$?
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 40 -c 12 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu2.log 2>&1
echo full_exit=$?
ls -la gpurun_out | tail
