TAG=${1:-fz}
timeout 300 python -m pytest tests/test_gpu_fused_rssm.py -m gpu -q -x 2>&1 | tail -3
run() {
  env $1 timeout 300 python bench.py --workload sweep --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_sweep_$2_$TAG.json 2> gpurun_out/bench_sweep_$2_$TAG.err; echo exit=$?
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_sweep_$2_$TAG.json"))
print("sweep $1", round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), d["gpu_launches"], d.get("imagination_only"), d["roofline"]["achieved"], d["roofline"]["ms_per_launch"])
PY
}
run RLSB_FUSED_RSSM=1 f1
run RLSB_FUSED_RSSM=0 f0
run RLSB_FUSED_RSSM=1 f1b
