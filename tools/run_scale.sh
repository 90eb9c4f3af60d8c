# bench.py on N GPUs of one box, launched the way the driver does; result under gpurun_out/
# usage: bash tools/run_scale.sh <n_gpus> <workload> <steps> <tag>
N=$1; W=$2; K=$3; TAG=$4
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --gpus 1 --workload $W --steps $K --warmup 3 > gpurun_out/bench_${TAG}_${W}_${N}gpu.json 2> gpurun_out/bench_${TAG}_${W}_${N}gpu.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload $W --steps $K --warmup 3 > gpurun_out/bench_${TAG}_${W}_${N}gpu.json 2> gpurun_out/bench_${TAG}_${W}_${N}gpu.err
fi
tail -c 600 gpurun_out/bench_${TAG}_${W}_${N}gpu.json; tail -2 gpurun_out/bench_${TAG}_${W}_${N}gpu.err
