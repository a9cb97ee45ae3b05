# A/B of an environment switch on the bench workloads: gpu_ab.sh <ENV_NAME> <tag>
V=$1; TAG=${2:-ab}
timeout 900 python -m pytest tests/test_gpu_ac_update.py tests/test_gpu_observe.py tests/test_gpu_slot_attention.py tests/test_gpu_agent.py tests/test_gpu_baseline_shapes.py -m gpu -q -x 2>&1 | tail -3
for w in sweep dino crafter; do for x in 0 1; do
  env $V=$x timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_${w}_${x}_$TAG.json 2> gpurun_out/bench_${w}_${x}_$TAG.err; 
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_${w}_${x}_$TAG.json"))
print("$w $V=$x", round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), d.get("imagination_only"))
PY
done; done
