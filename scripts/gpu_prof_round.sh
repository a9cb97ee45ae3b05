# bench (default sweep, with cpu baseline), reference arm, launch list, one full capture of the GRU contraction
TAG=${1:-v3}
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_sweep_$TAG.json 2> gpurun_out/bench_sweep_$TAG.err; echo bench_exit=$?
tail -c 2500 gpurun_out/bench_sweep_$TAG.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo ref_exit=$?
tail -c 600 gpurun_out/bench_ref_$TAG.json
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
echo launches_exit=$?
# the roofline kernel: GRU contraction timed alone at the end of bench.py (last gemm_kernel<1> launches)
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 600 -c 3 -o gpurun_out/prof_gru_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo full_exit=$?
ncu -i gpurun_out/prof_gru_$TAG.ncu-rep --page raw --csv > gpurun_out/raw_gru_$TAG.csv 2>/dev/null
ls -la gpurun_out | tail -12
