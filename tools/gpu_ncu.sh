# launch list + one full capture of the dominant kernel, each after the same command exits 0 without ncu
set -x
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-proofs"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv $CMD > gpurun_out/ncu_launch.log 2>&1
$CMD > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_chunk_acc -s 3 -c 1 -o gpurun_out/prof_chunk_acc_r01 $CMD > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/
tail -3 gpurun_out/ncu_full.log | cut -c1-300
