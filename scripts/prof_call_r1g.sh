set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --emd-genes 0 --no-mds --s10-genes 0 --pairs-genes 500"
$B > gpurun_out/r1g_plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1g_launches_bench.csv $B > gpurun_out/r1g_ncu_a.log 2>&1
$B > gpurun_out/r1g_plain_bench2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gmm_em_kernel -s 6 -c 2 -f -o gpurun_out/r1g_bench_step $B > gpurun_out/r1g_ncu_b.log 2>&1
S="python scripts/prof_s10.py 296"
$S > gpurun_out/r1g_plain_s10.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gmm_em_kernel -s 0 -c 7 -f -o gpurun_out/r1g_s10 $S > gpurun_out/r1g_ncu_c.log 2>&1
M="python scripts/prof_mds.py 4000 6"
$M > gpurun_out/r1g_plain_mds.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mds_rows -s 2 -c 1 -f -o gpurun_out/r1g_mds $M > gpurun_out/r1g_ncu_d.log 2>&1
cat gpurun_out/r1g_plain_s10.log gpurun_out/r1g_plain_mds.log
ls -la gpurun_out/*.ncu-rep
