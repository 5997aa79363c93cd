# k_msm_small (one-launch MSM for <= 4096 terms): full capture inside a config-3 proof (n = 2049 terms, BN254) + refreshed launch list
set -x
CMD="python tools/proof_trace.py 256 bn 1"
$CMD > gpurun_out/ncu_plain_small.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_msm_small -s 12 -c 1 -f -o /tmp/prof_small $CMD > gpurun_out/ncu_full_small.log 2>&1
ncu -i /tmp/prof_small.ncu-rep --page raw --csv > gpurun_out/k_msm_small_r02_raw.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_config3_r02.csv $CMD > gpurun_out/ncu_launch_c3.log 2>&1
