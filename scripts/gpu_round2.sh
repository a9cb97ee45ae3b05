# tests (all gpu), smoke, bench sweep (fused K4 path), config1
make -C oracle >/dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$?
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_sweep_v4.json 2> gpurun_out/bench_sweep_v4.err; echo bench_exit=$?
tail -c 2500 gpurun_out/bench_sweep_v4.json; tail -5 gpurun_out/bench_sweep_v4.err
timeout 600 python bench.py --workload config1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_config1_v4.json 2> gpurun_out/bench_config1_v4.err; echo bench_c1_exit=$?
timeout 600 python bench.py --workload dino --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dino_v4.json 2> gpurun_out/bench_dino_v4.err; echo bench_dino_exit=$?; tail -c 1200 gpurun_out/bench_dino_v4.json; tail -5 gpurun_out/bench_dino_v4.err
timeout 600 python bench.py --workload dino --rows 32768 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dino32k_v4.json 2> gpurun_out/bench_dino32k_v4.err; echo bench_dino32k_exit=$?; tail -c 1200 gpurun_out/bench_dino32k_v4.json; tail -5 gpurun_out/bench_dino32k_v4.err
tail -c 1500 gpurun_out/bench_config1_v4.json; tail -5 gpurun_out/bench_config1_v4.err
