# one GPU call: tests, smoke, bench (small + default)
make -C oracle >/dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$?
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --workload config1 --steps 5 --warmup 3 > gpurun_out/bench_config1.json 2> gpurun_out/bench_config1.err; echo bench_c1_exit=$?
tail -c 3000 gpurun_out/bench_config1.json; tail -5 gpurun_out/bench_config1.err
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_sweep.json 2> gpurun_out/bench_sweep.err; echo bench_exit=$?
tail -c 3000 gpurun_out/bench_sweep.json; tail -5 gpurun_out/bench_sweep.err
