# quick A/B of the current build: rollout trace, GPU tests, default + config1 + dino bench lines
TAG=${1:-q}
timeout 60 python scripts/rollout_trace.py config1 800 > gpurun_out/trace_config1_$TAG.md 2>&1; tail -17 gpurun_out/trace_config1_$TAG.md
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo pytest_exit=$?
tail -4 gpurun_out/pytest_gpu_$TAG.log
for w in sweep config1 dino; do
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_${w}_$TAG.json 2> gpurun_out/bench_${w}_$TAG.err; echo bench_${w}_exit=$?
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_${w}_$TAG.json"))
print("$w", round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), d["gpu_launches"], d.get("imagination_only"), d["roofline"]["achieved"], d["roofline"]["ms_per_launch"])
PY
done
