TAG=${1:-v4}
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
echo launches_exit=$?
