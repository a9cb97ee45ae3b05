# quick validation of the current build: GPU tests, smoke, default bench line
TAG=${1:-check}
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo pytest_exit=$?
tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_sweep_$TAG.json 2> gpurun_out/bench_sweep_$TAG.err; echo bench_exit=$?
python - <<PY
import json
d=json.load(open("gpurun_out/bench_sweep_$TAG.json"))
print("sweep", d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"], d.get("imagination_only"), d["roofline"]["achieved"])
PY
