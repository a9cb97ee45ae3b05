# launch list of the default bench command + one ncu --set full capture of the roofline kernel (GRU contraction)
TAG=${1:-v8}
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
echo launches_exit=$?
# the roofline kernel: GRU contraction timed alone at the end of bench.py (the last gemm_kernel<1> launches)
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 1292 -c 2 -o gpurun_out/prof_gru_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo full_exit=$?
ncu -i gpurun_out/prof_gru_$TAG.ncu-rep --page raw --csv > gpurun_out/raw_gru_$TAG.csv 2>/dev/null
rm -f gpurun_out/prof_gru_$TAG.ncu-rep
