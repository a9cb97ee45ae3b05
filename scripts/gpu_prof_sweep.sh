# sweep only: plain run, launch list, full ncu capture of the GRU kernel bench.py times alone (its last gemm_kernel launches)
TAG=${1:-r02}
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/plain_sweep_$TAG.json 2> gpurun_out/plain_sweep_$TAG.err || { echo plain sweep failed; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_sweep_$TAG.csv $CMD > gpurun_out/ncu_list_sweep_$TAG.log 2>&1
echo launches_sweep_exit=$?
python scripts/launch_summary.py gpurun_out/launches_sweep_$TAG.csv > gpurun_out/launch_summary_sweep_$TAG.md; head -16 gpurun_out/launch_summary_sweep_$TAG.md
SKIP=$(python - <<PY
import csv
n=0
for row in csv.DictReader(l for l in open("gpurun_out/launches_sweep_$TAG.csv") if not l.startswith("==")):
    if "gemm_kernel" in row["Kernel Name"]: n+=1
print(max(0,n-10))
PY
)
echo skip=$SKIP
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s $SKIP -c 2 -f -o gpurun_out/prof_gru_$TAG $CMD > gpurun_out/ncu_full_gru_$TAG.log 2>&1
echo full_gru_exit=$?
ncu -i gpurun_out/prof_gru_$TAG.ncu-rep --page raw --csv > gpurun_out/raw_gru_$TAG.csv 2>/dev/null
python scripts/ncu_raw_summary.py gpurun_out/raw_gru_$TAG.csv | head -40
ncu -i gpurun_out/prof_gru_$TAG.ncu-rep --page details > gpurun_out/details_gru_$TAG.txt 2>/dev/null
rm -f gpurun_out/prof_gru_$TAG.ncu-rep
