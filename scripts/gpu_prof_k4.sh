# full ncu captures of the K4 kernels at the sweep shape: weight gradient (layer 0 + hidden) and the dX GEMM with the LN-backward epilogue
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain_k4.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 5 -c 5 -o gpurun_out/prof_wgrad $CMD > gpurun_out/ncu_wgrad.log 2>&1
echo wgrad_exit=$?
ncu --set full --clock-control none --import-source on -k regex:"gemm_kernel<3>|gemm_kernel<\(int\)3>" -s 4 -c 2 -o gpurun_out/prof_epibwd $CMD > gpurun_out/ncu_epibwd.log 2>&1
echo epibwd_exit=$?
ncu --set full --clock-control none --import-source on -k regex:"gemm_kernel<4>|gemm_kernel<\(int\)4>" -s 4 -c 2 -o gpurun_out/prof_fwdsave $CMD > gpurun_out/ncu_fwdsave.log 2>&1
echo fwdsave_exit=$?
for f in wgrad epibwd fwdsave; do ncu -i gpurun_out/prof_$f.ncu-rep --page raw --csv > gpurun_out/raw_$f.csv 2>/dev/null; done
ls -la gpurun_out | grep -E "prof_|raw_" | tail
