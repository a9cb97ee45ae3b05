# latency-bound sizes: the engine's own choice of rollout kernel vs the chained rollout (RLSB_PERSISTENT=0) vs the persistent
# kernels forced (RLSB_PERSISTENT=1), sweep dims (D = 1024, discrete) and dino dims (D = 200, continuous, with backward)
TAG=${1:-sm}
run() {
  env $1 timeout 300 python bench.py --workload $3 --rows $2 --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_small_$TAG.json 2> gpurun_out/bench_small_$TAG.err;
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_small_$TAG.json"))
print("$3 rows $2 $1", round(d["ms_per_step"],3), d["gpu_launches"], round(d.get("imagination_only",{}).get("ms",0),3), d["config"].get("rollout_kernel","")[:12])
PY
}
for r in ${ROWS:-800 1024 1536 1920 2048}; do
  for w in sweep dino; do
    run AUTO=1 $r $w
    run RLSB_PERSISTENT=0 $r $w
    run RLSB_PERSISTENT=1 $r $w
  done
done
