# multi-GPU batch (gpurun --gpus G): 2-rank NCCL tests, weak-scaling sweep line, Crafter-shaped full train step (configs[4])
G=${1:-2}; TAG=${2:-r02}
if [ "$G" = "2" ]; then
  timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/pytest_multi_$TAG.log 2>&1; echo pytest_multi_exit=$?; tail -3 gpurun_out/pytest_multi_$TAG.log
fi
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $RUN bench.py --gpus $G --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sweep_g${G}_$TAG.json 2> gpurun_out/bench_sweep_g${G}_$TAG.err; echo sweep_exit=$?
cut -c1-260 gpurun_out/bench_sweep_g${G}_$TAG.json
timeout 600 $RUN bench.py --gpus $G --workload crafter --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_crafter_g${G}_$TAG.json 2> gpurun_out/bench_crafter_g${G}_$TAG.err; echo crafter_exit=$?
cut -c1-260 gpurun_out/bench_crafter_g${G}_$TAG.json; tail -3 gpurun_out/bench_crafter_g${G}_$TAG.err
