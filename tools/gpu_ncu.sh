# launch list + full captures of the MSM kernels, each after the same command exits 0 without ncu.
# The .ncu-rep files stay on the box (/tmp): only the CSV pages come back (gpurun_out/ is limited to 64 MiB).
set -x
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-proofs"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv $CMD > gpurun_out/ncu_launch.log 2>&1
for K in ${KERNELS:-k_chunk_acc k_reduce_l1 k_scatter}; do
  $CMD > gpurun_out/ncu_plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o /tmp/prof_${K}_r01 $CMD > gpurun_out/ncu_full_$K.log 2>&1
  ncu -i /tmp/prof_${K}_r01.ncu-rep --page raw --csv > gpurun_out/${K}_r01_raw.csv 2>/dev/null
done
ls -la gpurun_out/ | tail -12
