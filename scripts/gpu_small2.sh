TAG=${1:-sm}
run() {
  env $1 timeout 300 python bench.py --workload $3 --rows $2 --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_small_$TAG.json 2> gpurun_out/bench_small_$TAG.err; 
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_small_$TAG.json"))
print("$3 rows $2 $1", round(d["ms_per_step"],3), d["gpu_launches"], round(d.get("imagination_only",{}).get("ms",0),3), d["config"].get("rollout_kernel","")[:12])
PY
}
for r in 800 1024 1280 1536 2048; do
run RLSB_PERSISTENT=1 $r sweep
run RLSB_PERSISTENT=0 $r sweep
done
for r in 800 1024 2048; do
run RLSB_PERSISTENT=1 $r dino
run RLSB_PERSISTENT=0 $r dino
done
