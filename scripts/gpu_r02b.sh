TAG=${1:-r02b}
timeout 900 python -m pytest tests/test_gpu_agent.py tests/test_gpu_baseline_shapes.py tests/test_gpu_configs.py tests/test_gpu_slotted.py -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo pytest_exit=$?
tail -12 gpurun_out/pytest_gpu_$TAG.log
for w in crafter slotted; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_${w}_$TAG.json 2> gpurun_out/bench_${w}_$TAG.err; echo bench_${w}_exit=$?
  cut -c1-210 gpurun_out/bench_${w}_$TAG.json; tail -2 gpurun_out/bench_${w}_$TAG.err
done
for w in config1 dino; do
CMD="python bench.py --workload $w --steps 1 --warmup 3 --no-cpu-baseline --no-extras"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_${w}_$TAG.csv $CMD > gpurun_out/ncu_list_${w}_$TAG.log 2>&1
echo launches_exit=$?
done
