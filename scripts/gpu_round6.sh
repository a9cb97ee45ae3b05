# pipelined EPI_BWD: parity tests, sweep bench, K4 phase timing via launch list
TAG=${1:-v13}
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo pytest_exit=$?
tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sweep_$TAG.json 2> gpurun_out/bench_sweep_$TAG.err; echo bench_exit=$?
python - <<PY
import json
d=json.load(open("gpurun_out/bench_sweep_$TAG.json"))
print("sweep", d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"], d.get("imagination_only"), d["roofline"]["achieved"])
PY
RLSB_GRAPH_MAX_ROWS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list_$TAG.log 2>&1; echo launches_exit=$?
python scripts/launch_summary.py gpurun_out/launches_$TAG.csv 2>&1 | head -30
