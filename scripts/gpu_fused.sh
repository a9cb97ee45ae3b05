# fused RSSM epilogues of the chained rollout: parity tests, the GRU cell's phase trace, A/B of the sweep bench with the switch
# off (0) / LayerNorm layers only (2) / everything (1, default)
TAG=${1:-fz}
timeout 300 python -m pytest tests/test_gpu_fused_rssm.py -m gpu -q -x -s 2>&1 | grep "fused rssm\|parity\|passed\|failed\|Error\|error" | tail -40
timeout 120 python scripts/gru_cell_trace.py 2>&1 | tail -18
for x in 0 2 1; do
  RLSB_FUSED_RSSM=$x timeout 300 python bench.py --workload sweep --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_sweep_${x}_$TAG.json 2> gpurun_out/bench_sweep_${x}_$TAG.err; echo exit=$?
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_sweep_${x}_$TAG.json"))
print("sweep RLSB_FUSED_RSSM=$x", round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), d["gpu_launches"], d.get("imagination_only"), d["roofline"]["achieved"], d["roofline"]["ms_per_launch"])
PY
done
