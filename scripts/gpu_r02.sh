# round-2 batch: tests, smoke, every bench workload, reference arm, launch list
TAG=${1:-r02a}
make -C oracle >/dev/null 2>&1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo pytest_exit=$?
tail -15 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_sweep_$TAG.json 2> gpurun_out/bench_sweep_$TAG.err; echo bench_exit=$?
cut -c1-400 gpurun_out/bench_sweep_$TAG.json; tail -2 gpurun_out/bench_sweep_$TAG.err
for w in config1 dino crafter slotted; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_${w}_$TAG.json 2> gpurun_out/bench_${w}_$TAG.err; echo bench_${w}_exit=$?
  cut -c1-210 gpurun_out/bench_${w}_$TAG.json; tail -2 gpurun_out/bench_${w}_$TAG.err
done
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_$TAG.json 2> gpurun_out/bench_reference_$TAG.err; echo ref_exit=$?; cut -c1-300 gpurun_out/bench_reference_$TAG.json
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
echo launches_exit=$?
python scripts/launch_summary.py gpurun_out/launches_$TAG.csv 2>&1 | head -40
