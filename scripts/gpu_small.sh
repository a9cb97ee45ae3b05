# latency-bound sizes of the sweep dims: 2048 / 4096 rows per GPU, fused / unfused / persistent switches
TAG=${1:-sm}
run() {
  env $1 timeout 300 python bench.py --workload sweep --rows $2 --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_small_$TAG.json 2> gpurun_out/bench_small_$TAG.err; 
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_small_$TAG.json"))
print("rows $2 $1", round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), d["gpu_launches"], d.get("imagination_only"), d["config"].get("rollout_kernel","")[:40])
PY
}
run RLSB_FUSED_RSSM=1 2048
run RLSB_FUSED_RSSM=0 2048
run RLSB_PERSISTENT=0 2048
run RLSB_FUSED_RSSM=1 4096
run RLSB_FUSED_RSSM=0 4096
run RLSB_FUSED_RSSM=1 8192
run RLSB_FUSED_RSSM=0 8192
