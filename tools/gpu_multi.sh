#!/bin/bash
# multi-GPU visit: tools/gpu_multi.sh TAG N [particles]  -> bench line at N GPUs (with the ring-vs-single verification)
set -u
mkdir -p gpurun_out
TAG=${1:-m2}; N=${2:-2}; P=${3:-1.0e7}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 20 --warmup 3 --particles $P ${BENCH_ARGS:-} > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
echo "bench rc=$?"
cat gpurun_out/${TAG}_bench_${N}gpu.json | cut -c1-3000; tail -5 gpurun_out/${TAG}_bench_${N}gpu.err
