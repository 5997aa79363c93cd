# config 3 (n = 2^14, BN254, window tables, hybrid IPP, recorded circuit): launch list of one proof + one verification and a
# full capture of the table-sum kernel.  CSV pages only come back.
set -x
CMD="python tools/proof_trace.py 256 bn 1"
$CMD > gpurun_out/ncu_plain_tbl.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_table_sum -s 12 -c 1 -f -o /tmp/prof_tbl $CMD > gpurun_out/ncu_full_tbl.log 2>&1
ncu -i /tmp/prof_tbl.ncu-rep --page raw --csv > gpurun_out/k_table_sum_n16384_r02_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:k_ipp_materialise -s 1 -c 1 -f -o /tmp/prof_mat $CMD > gpurun_out/ncu_full_mat.log 2>&1
ncu -i /tmp/prof_mat.ncu-rep --page raw --csv > gpurun_out/k_ipp_materialise_r02_raw.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_config3_r02.csv $CMD > gpurun_out/ncu_launch_c3.log 2>&1
