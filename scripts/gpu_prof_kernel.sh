# one full ncu capture of a kernel of the default bench, with the per-instruction source page
# usage: [WORKLOAD=config1 WARMUP=3] gpu_prof_kernel.sh <tag> <kernel regex> <launches to skip>
TAG=$1; RE=$2; SKIP=${3:-0}
CMD="python bench.py ${WORKLOAD:+--workload $WORKLOAD} --steps 1 --warmup ${WARMUP:-1} --no-cpu-baseline --no-extras"
$CMD > /dev/null 2>&1 || { echo plain run failed; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c 1 -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo full_exit=$?
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/raw_$TAG.csv 2>/dev/null
ncu -i gpurun_out/prof_$TAG.ncu-rep --page source --csv > gpurun_out/source_$TAG.csv 2>/dev/null
ncu -i gpurun_out/prof_$TAG.ncu-rep --page details > gpurun_out/details_$TAG.txt 2>/dev/null
python scripts/ncu_raw_summary.py gpurun_out/raw_$TAG.csv | head -20
ls -la gpurun_out/source_$TAG.csv
rm -f gpurun_out/prof_$TAG.ncu-rep
