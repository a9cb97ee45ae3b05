TAG=r02f
for w in config1 dino; do
  CMD2="python bench.py --workload $w --steps 1 --warmup 3 --no-cpu-baseline --no-extras"
  $CMD2 > gpurun_out/plain_${w}_$TAG.json 2> gpurun_out/plain_${w}_$TAG.err || { echo plain $w failed; continue; }
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_${w}_$TAG.csv $CMD2 > gpurun_out/ncu_list_${w}_$TAG.log 2>&1
  python scripts/launch_summary.py gpurun_out/launches_${w}_$TAG.csv > gpurun_out/launch_summary_${w}_$TAG.md; head -14 gpurun_out/launch_summary_${w}_$TAG.md
done
