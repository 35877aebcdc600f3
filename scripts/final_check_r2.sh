# Round-2 closing run on one B200: GPU test suite, smoke, both bench arms, and a fresh full capture of the EMD kernel.
set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2f_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2f_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2f_smoke.log 2>&1; tail -2 gpurun_out/r2f_smoke.log
timeout 600 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -c 3000 gpurun_out/r2f_bench.json
timeout 600 python bench.py --impl reference > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err; tail -c 600 gpurun_out/r2f_bench_ref.json
E="python scripts/prof_emd.py 148"
$E > gpurun_out/r2f_plain_emd.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:emd_kernel -s 1 -c 1 -f -o gpurun_out/r2f_emd_visium $E > gpurun_out/r2f_ncu_emd.log 2>&1
tail -2 gpurun_out/r2f_plain_emd.log; ls -la gpurun_out/r2f_*.ncu-rep
