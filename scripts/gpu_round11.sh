TAG=${1:-v19}
timeout 120 python scripts/hidden_variants.py 2>&1 | tail -7; echo hidden_exit=$?
RLSB_STAGED=0 timeout 120 python scripts/hidden_variants.py 2>&1 | head -2
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_imagine.py -m gpu -q -x > gpurun_out/pytest_a_$TAG.log 2>&1; echo pytest_a_exit=$?
tail -3 gpurun_out/pytest_a_$TAG.log
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo pytest_exit=$?
tail -3 gpurun_out/pytest_gpu_$TAG.log
for sp in 0 1; do
RLSB_STAGED=$sp timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sweep_staged${sp}_$TAG.json 2> gpurun_out/bench_sweep_staged${sp}_$TAG.err; echo bench_exit=$?
python - <<PY
import json
d=json.load(open("gpurun_out/bench_sweep_staged${sp}_$TAG.json"))
print("sweep staged=$sp", d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"], d.get("imagination_only"), d["roofline"]["achieved"])
PY
done
