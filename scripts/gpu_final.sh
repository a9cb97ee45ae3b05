# round-end batch: tests, smoke, every bench workload, reference arm, launch list, full capture of the roofline kernel
TAG=${1:-v25}
make -C oracle >/dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo pytest_exit=$?
tail -2 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_sweep_$TAG.json 2> gpurun_out/bench_sweep_$TAG.err; echo bench_exit=$?
cut -c1-330 gpurun_out/bench_sweep_$TAG.json; tail -2 gpurun_out/bench_sweep_$TAG.err
for w in config1 dino crafter slotted; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${w}_$TAG.json 2> gpurun_out/bench_${w}_$TAG.err; echo bench_${w}_exit=$?
  cut -c1-210 gpurun_out/bench_${w}_$TAG.json; tail -2 gpurun_out/bench_${w}_$TAG.err
done
RLSB_ACTOR_REUSE=0 timeout 600 python bench.py --workload config1 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-210
timeout 600 python bench.py --workload dino --rows 32768 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dino32k_$TAG.json 2> gpurun_out/bench_dino32k_$TAG.err; cut -c1-210 gpurun_out/bench_dino32k_$TAG.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_$TAG.json 2> gpurun_out/bench_reference_$TAG.err; echo ref_exit=$?; cut -c1-300 gpurun_out/bench_reference_$TAG.json
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
echo launches_exit=$?
SKIP=$(python - <<PY
import csv
n=0
for row in csv.DictReader(l for l in open("gpurun_out/launches_$TAG.csv") if not l.startswith("==")):
    if "gemm_kernel" in row["Kernel Name"]: n+=1
print(max(0,n-10))
PY
)
echo skip=$SKIP
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s $SKIP -c 2 -f -o gpurun_out/prof_gru_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo full_exit=$?
ncu -i gpurun_out/prof_gru_$TAG.ncu-rep --page raw --csv > gpurun_out/raw_gru_$TAG.csv 2>/dev/null
rm -f gpurun_out/prof_gru_$TAG.ncu-rep
python scripts/ncu_raw_summary.py gpurun_out/raw_gru_$TAG.csv | head -34
