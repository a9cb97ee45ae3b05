TAG=${1:-v15}
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo pytest_exit=$?
tail -5 gpurun_out/pytest_gpu_$TAG.log
for r in 0 1; do
RLSB_ACTOR_REUSE=$r timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sweep_reuse${r}_$TAG.json 2> gpurun_out/bench_sweep_reuse${r}_$TAG.err; echo bench_exit=$?
tail -2 gpurun_out/bench_sweep_reuse${r}_$TAG.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_sweep_reuse${r}_$TAG.json"))
print("sweep reuse=$r", d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"], d.get("imagination_only"), d["roofline"]["achieved"])
PY
done
for w in config1 crafter; do
RLSB_ACTOR_REUSE=1 timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${w}_$TAG.json 2> gpurun_out/bench_${w}_$TAG.err; echo bench_${w}_exit=$?
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${w}_$TAG.json"))
print("$w", d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"])
PY
done
