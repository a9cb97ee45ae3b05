CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain_k4b.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"gemm_kernel<\(int\)3>" -s 4 -c 2 -o gpurun_out/prof_epibwd $CMD > gpurun_out/ncu_epibwd.log 2>&1
echo epibwd_exit=$?
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"gemm_kernel<\(int\)4>" -s 4 -c 2 -o gpurun_out/prof_fwdsave $CMD > gpurun_out/ncu_fwdsave.log 2>&1
echo fwdsave_exit=$?
for f in epibwd fwdsave; do ncu -i gpurun_out/prof_$f.ncu-rep --page raw --csv > gpurun_out/raw_$f.csv 2>/dev/null; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_v5.csv $CMD > gpurun_out/ncu_list_v5.log 2>&1
echo launches_exit=$?
ls -la gpurun_out | grep -E "raw_|launches_v5" | tail
