export RLSB_SPLIT2=0
CMD="python scripts/hidden_biasonly.py"
$CMD > gpurun_out/biasonly_plain.log 2>&1 || { echo plain_failed; tail -5 gpurun_out/biasonly_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 3 -c 1 -f -o gpurun_out/prof_biasonly $CMD > gpurun_out/ncu_biasonly.log 2>&1
echo full_exit=$?
ncu -i gpurun_out/prof_biasonly.ncu-rep --page raw --csv > gpurun_out/raw_biasonly.csv 2>/dev/null
ncu -i gpurun_out/prof_biasonly.ncu-rep --page source --csv > gpurun_out/src_biasonly.csv 2>/dev/null
rm -f gpurun_out/prof_biasonly.ncu-rep
ls -la gpurun_out/*biasonly*.csv
