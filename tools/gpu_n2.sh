#!/bin/bash
# 2-GPU visit: data-parallel bench (NCCL) next to the 1-GPU bench on the same box.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
cat gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_same_box.json 2> gpurun_out/bench_n1.err; echo "bench n1 rc=$?"
cat gpurun_out/bench_n1_same_box.json
