# Round-2 captures: launch list + full capture of the bench step, the distance kernel, the EMD kernel (Visium-size, multiscale
# start) and the staging kernels.  Every command runs once WITHOUT ncu first (exit 0), then under ncu.
set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --emd-genes 0 --no-mds --s10-genes 0 --pairs-genes 500 --no-adata"
$B > gpurun_out/r2_plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv $B > gpurun_out/r2_ncu_a.log 2>&1
$B > gpurun_out/r2_plain_bench2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gmm_em_kernel -s 6 -c 2 -f -o gpurun_out/r2_bench_step $B > gpurun_out/r2_ncu_b.log 2>&1
D="python scripts/prof_dist.py 1500"
$D > gpurun_out/r2_plain_dist.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gmm_dist_kernel -s 1 -c 1 -f -o gpurun_out/r2_dist $D > gpurun_out/r2_ncu_c.log 2>&1
E="python scripts/prof_emd.py 148"
$E > gpurun_out/r2_plain_emd.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:emd_kernel -s 1 -c 1 -f -o gpurun_out/r2_emd_visium $E > gpurun_out/r2_ncu_d.log 2>&1
cat gpurun_out/r2_plain_dist.log gpurun_out/r2_plain_emd.log | grep -v Warn
ls -la gpurun_out/r2_*.ncu-rep
