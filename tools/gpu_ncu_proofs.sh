# launch list (gpu__time_duration.sum) of the lock-step batch prover on 1024 64-bit range
# proofs; the ncu pass runs only after the same command has exited 0 without ncu
set -x
P="python tools/batch_prove_trace.py 1 1024"
$P > gpurun_out/pp_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_prove_batch_r01.csv $P > gpurun_out/pp_ncu.log 2>&1
ls -la gpurun_out | tail -8
