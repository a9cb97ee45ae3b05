# round-end measurement batch: tests, smoke, every bench workload, reference arm
TAG=${1:-v8}
make -C oracle >/dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo pytest_exit=$?
tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_sweep_$TAG.json 2> gpurun_out/bench_sweep_$TAG.err; echo bench_exit=$?
tail -c 2800 gpurun_out/bench_sweep_$TAG.json; tail -3 gpurun_out/bench_sweep_$TAG.err
for w in config1 dino crafter slotted; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${w}_$TAG.json 2> gpurun_out/bench_${w}_$TAG.err; echo bench_${w}_exit=$?
  cut -c1-210 gpurun_out/bench_${w}_$TAG.json; tail -2 gpurun_out/bench_${w}_$TAG.err
done
timeout 600 python bench.py --workload dino --rows 32768 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dino32k_$TAG.json 2> gpurun_out/bench_dino32k_$TAG.err; cut -c1-210 gpurun_out/bench_dino32k_$TAG.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_$TAG.json 2> gpurun_out/bench_reference_$TAG.err; echo ref_exit=$?; cut -c1-400 gpurun_out/bench_reference_$TAG.json
