# ncu full + source capture of one hidden head layer (EPI_LN_ACT, K=448, N=400) launched alone
CMD="python scripts/hidden_only.py"
$CMD > gpurun_out/hidden_plain.log 2>&1 || { echo plain_failed; tail -5 gpurun_out/hidden_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 3 -c 1 -f -o gpurun_out/prof_hidden_v25 $CMD > gpurun_out/ncu_hidden.log 2>&1
echo full_exit=$?
ncu -i gpurun_out/prof_hidden_v25.ncu-rep --page raw --csv > gpurun_out/raw_hidden_v25.csv 2>/dev/null
ncu -i gpurun_out/prof_hidden_v25.ncu-rep --page source --csv > gpurun_out/src_hidden_v25.csv 2>/dev/null
rm -f gpurun_out/prof_hidden_v25.ncu-rep; ls -la gpurun_out/raw_hidden_v25.csv gpurun_out/src_hidden_v25.csv
