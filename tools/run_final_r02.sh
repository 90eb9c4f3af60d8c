# final single-GPU measurements of round 2 (bench lines, reference arm, ncu captures); results under gpurun_out/
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 500 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench_r02_c3_final.json 2> gpurun_out/bench_r02_c3_final.err; tail -c 300 gpurun_out/bench_r02_c3_final.json
timeout 400 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_r02_c4_final.json 2> gpurun_out/bench_r02_c4_final.err; tail -c 200 gpurun_out/bench_r02_c4_final.json
timeout 600 python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_r02_c5_final.json 2> gpurun_out/bench_r02_c5_final.err; tail -c 200 gpurun_out/bench_r02_c5_final.json
timeout 200 python bench.py --workload c3s --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_r02_c3s_final.json
bash tools/run_prof_r02.sh > /dev/null 2>&1
