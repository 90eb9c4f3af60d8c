set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 400 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench_r02_c3_final.json 2> gpurun_out/bench_r02_c3_final.err; tail -c 400 gpurun_out/bench_r02_c3_final.json
timeout 400 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/bench_r02_c3_reference.json 2> gpurun_out/bench_r02_c3_reference.err; tail -c 1500 gpurun_out/bench_r02_c3_reference.json; tail -3 gpurun_out/bench_r02_c3_reference.err
bash tools/run_prof_r02.sh > /dev/null 2>&1
timeout 400 python bench.py --workload c4 --steps 5 --warmup 3 > gpurun_out/bench_r02_c4_final.json 2> gpurun_out/bench_r02_c4_final.err; tail -c 300 gpurun_out/bench_r02_c4_final.json
timeout 600 python bench.py --workload c5 --steps 3 --warmup 3 > gpurun_out/bench_r02_c5_final.json 2> gpurun_out/bench_r02_c5_final.err; tail -c 300 gpurun_out/bench_r02_c5_final.json
