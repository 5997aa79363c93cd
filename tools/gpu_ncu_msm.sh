# MSM profiles of the round: launch list + full captures of k_chunk_acc and k_reduce_l1 (after the plain command exits 0)
set -x
R=r02
MSM="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-proofs --no-sweep"
$MSM > gpurun_out/ncu_plain_msm.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv $MSM > gpurun_out/ncu_launch_msm.log 2>&1
cap() {  # kernel-regex  skip  tag  command...
  K=$1; S=$2; T=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -f -o /tmp/prof_${T}_$R "$@" > gpurun_out/ncu_full_$T.log 2>&1
  ncu -i /tmp/prof_${T}_$R.ncu-rep --page raw --csv > gpurun_out/${T}_${R}_raw.csv 2>/dev/null
}
cap k_chunk_acc 3 k_chunk_acc $MSM
cap k_reduce_l1 3 k_reduce_l1 $MSM
