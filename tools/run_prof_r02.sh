# ncu captures of one c3 bootstrap step (1536 replicates): launch list + --set full of every kernel of the second step
set -x
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_c3.csv python tools/profile_step.py c3 3 1536 > gpurun_out/launches_r02.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:resample_images_kernel|gram_mma_kernel|vote_mma_kernel|gram_finalize_kernel|solve_kernel' -s 6 -c 6 -f -o gpurun_out/r02_c3_full python tools/profile_step.py c3 2 1536 > gpurun_out/ncu_full_r02.log 2>&1
tail -3 gpurun_out/ncu_full_r02.log
ls -la gpurun_out/*.ncu-rep
