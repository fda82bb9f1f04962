#!/bin/bash
# bench.py at N GPUs exactly as the driver launches it, with the NVLink byte counters read before and after.
# usage: scripts/bench_n.sh N [tag] [extra bench args...]
N=$1; TAG=${2:-r02}; shift 2
mkdir -p gpurun_out
nvidia-smi nvlink -gt d > gpurun_out/nvlink_before_${TAG}_n$N.txt 2>&1
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 "$@" > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N "$@" > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err
fi
echo "rc=$?"
nvidia-smi nvlink -gt d > gpurun_out/nvlink_after_${TAG}_n$N.txt 2>&1
tail -c 600 gpurun_out/bench_${TAG}_n$N.err
