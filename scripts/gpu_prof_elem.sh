# ncu --set full of the HBM-bound K1 kernels (one launch each, mid-rollout)
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
for k in ${KERNELS:-gru_gate_kernel ln_act_kernel sample_latent_kernel}; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 20 -c 1 -o gpurun_out/prof_$k $CMD > gpurun_out/ncu_$k.log 2>&1
  ncu -i gpurun_out/prof_$k.ncu-rep --page raw --csv > gpurun_out/raw_$k.csv 2>/dev/null
  ncu -i gpurun_out/prof_$k.ncu-rep --page source --csv > gpurun_out/src_$k.csv 2>/dev/null
  ncu -i gpurun_out/prof_$k.ncu-rep --page details > gpurun_out/det_$k.txt 2>/dev/null
  rm -f gpurun_out/prof_$k.ncu-rep
done
echo done
