# A/B: batched pack launches (config1 / dino), graph replay at the sweep shape; full GPU test suite first
TAG=${1:-v11}
make -C oracle >/dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo pytest_exit=$?
tail -3 gpurun_out/pytest_gpu_$TAG.log
for w in config1 dino; do
  for b in 0 1; do
    RLSB_PACK_BATCH=$b timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${w}_batch${b}_$TAG.json 2> gpurun_out/bench_${w}_batch${b}_$TAG.err; echo bench_${w}_batch${b}_exit=$?
    python - <<PY
import json
d=json.load(open("gpurun_out/bench_${w}_batch${b}_$TAG.json"))
print("$w batch=$b", d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"], d.get("imagination_only"))
PY
  done
done
for g in 8192 32768; do
  RLSB_GRAPH_MAX_ROWS=$g timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sweep_graph${g}_$TAG.json 2> gpurun_out/bench_sweep_graph${g}_$TAG.err; echo bench_sweep_graph${g}_exit=$?
  tail -3 gpurun_out/bench_sweep_graph${g}_$TAG.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_sweep_graph${g}_$TAG.json"))
print("sweep graph_max_rows=$g", d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"], d.get("imagination_only"))
PY
done
