#!/bin/bash
# multi-GPU visit: data-parallel bench at N = $1 (peer-memory SyncBN), short
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?"
cut -c1-400 gpurun_out/bench_n$N.json; grep -v "Warning\|warnings.warn\|^\*\*\*\|OMP_NUM" gpurun_out/bench_n$N.err | tail -8
