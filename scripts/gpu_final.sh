# round-end batch: GPU tests (with their printed parity figures), smoke, every bench workload, the reference arm
TAG=${1:-r02}
make -C oracle >/dev/null 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo pytest_exit=$?
tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_sweep_$TAG.json 2> gpurun_out/bench_sweep_$TAG.err; echo bench_exit=$?
cut -c1-330 gpurun_out/bench_sweep_$TAG.json; tail -2 gpurun_out/bench_sweep_$TAG.err
for w in config1 dino crafter slotted; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_${w}_$TAG.json 2> gpurun_out/bench_${w}_$TAG.err; echo bench_${w}_exit=$?
  cut -c1-210 gpurun_out/bench_${w}_$TAG.json
done
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_$TAG.json 2> gpurun_out/bench_reference_$TAG.err; echo ref_exit=$?; cut -c1-300 gpurun_out/bench_reference_$TAG.json
