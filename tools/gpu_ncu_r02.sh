# Round-2 profiles: launch lists + full captures, each after the same command exits 0 without ncu.
# The .ncu-rep files stay on the box (/tmp): only the CSV pages come back (gpurun_out/ is limited to 64 MiB).
set -x
R=r02
MSM="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-proofs --no-sweep"
VB="python tools/verify_batch_bench.py 4096"
PB="python tools/prove_batch_bench.py 2048 1 64 bls 2 0"
export VB_MODES=0 BPH_VB_DRIVERS=1
$MSM > gpurun_out/ncu_plain_msm.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv $MSM > gpurun_out/ncu_launch_msm.log 2>&1
$VB > gpurun_out/ncu_plain_vb.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_verify_batch_$R.csv $VB > gpurun_out/ncu_launch_vb.log 2>&1
$PB > gpurun_out/ncu_plain_pb.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_prove_batch_$R.csv $PB > gpurun_out/ncu_launch_pb.log 2>&1
cap() {  # kernel-regex  skip  tag  command...
  K=$1; S=$2; T=$3; shift 3
  "$@" > gpurun_out/ncu_plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -f -o /tmp/prof_${T}_$R "$@" > gpurun_out/ncu_full_$T.log 2>&1
  ncu -i /tmp/prof_${T}_$R.ncu-rep --page raw --csv > gpurun_out/${T}_${R}_raw.csv 2>/dev/null
}
cap k_chunk_acc 3 k_chunk_acc $MSM
cap k_reduce_l1 3 k_reduce_l1 $MSM
cap k_batch_fixed_warp 2 k_batch_fixed_warp_verify $VB
cap k_vb_scalars 2 k_vb_scalars $VB
cap k_vb_transcript 2 k_vb_transcript $VB
cap k_batch_horner 2 k_batch_horner $VB
cap k_batch_windows 2 k_batch_windows $VB
cap k_batch_fixed_warp 8 k_batch_fixed_warp_ipp $PB
cap k_pd_normalise 5 k_pd_normalise $PB
ls -la gpurun_out/ | tail -20
