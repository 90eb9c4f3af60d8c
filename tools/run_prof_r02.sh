set -x
mkdir -p gpurun_out
PLSPM_TRACE=1 timeout 300 python bench.py --workload c3 --steps 5 --warmup 3 --no-parity > gpurun_out/bench_c3_r02_v5.json 2> gpurun_out/bench_c3_r02_v5.err
tail -20 gpurun_out/bench_c3_r02_v5.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_c3.csv python tools/profile_step.py c3 3 1536 > gpurun_out/launches_r02.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:gram_mma_kernel|vote_mma_kernel|gram_finalize_kernel|solve_kernel|counts8_image_kernel|vote_c8_image_kernel|counts_kernel' -s 7 -c 7 -f -o gpurun_out/r02_c3_full python tools/profile_step.py c3 2 1536 > gpurun_out/ncu_full_r02.log 2>&1
tail -3 gpurun_out/ncu_full_r02.log
ls -la gpurun_out/*.ncu-rep
