# profile capture of the current build: launch lists (sweep, config1, dino) + full ncu captures of the GRU contraction and
# of the persistent rollout kernel.  Every profiled command first exits 0 without ncu.
TAG=${1:-r02}
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/plain_sweep_$TAG.json 2> gpurun_out/plain_sweep_$TAG.err || { echo plain sweep failed; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_sweep_$TAG.csv $CMD > gpurun_out/ncu_list_sweep_$TAG.log 2>&1
echo launches_sweep_exit=$?
python scripts/launch_summary.py gpurun_out/launches_sweep_$TAG.csv > gpurun_out/launch_summary_sweep_$TAG.md; head -16 gpurun_out/launch_summary_sweep_$TAG.md
# the last gemm_kernel launches of the command are bench.py's roofline repetitions of the GRU contraction
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
rm -f gpurun_out/prof_gru_$TAG.ncu-rep
for w in config1 dino; do
  CMD2="python bench.py --workload $w --steps 1 --warmup 3 --no-cpu-baseline --no-extras"
  $CMD2 > gpurun_out/plain_${w}_$TAG.json 2> gpurun_out/plain_${w}_$TAG.err || { echo plain $w failed; continue; }
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_${w}_$TAG.csv $CMD2 > gpurun_out/ncu_list_${w}_$TAG.log 2>&1
  python scripts/launch_summary.py gpurun_out/launches_${w}_$TAG.csv > gpurun_out/launch_summary_${w}_$TAG.md; head -12 gpurun_out/launch_summary_${w}_$TAG.md
done
CMD2="python bench.py --workload config1 --steps 1 --warmup 3 --no-cpu-baseline --no-extras"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 3 -c 1 -f -o gpurun_out/prof_rollout_$TAG $CMD2 > gpurun_out/ncu_full_rollout_$TAG.log 2>&1
echo full_rollout_exit=$?
ncu -i gpurun_out/prof_rollout_$TAG.ncu-rep --page raw --csv > gpurun_out/raw_rollout_$TAG.csv 2>/dev/null
python scripts/ncu_raw_summary.py gpurun_out/raw_rollout_$TAG.csv | head -24
rm -f gpurun_out/prof_rollout_$TAG.ncu-rep
