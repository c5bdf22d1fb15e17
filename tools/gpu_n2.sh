#!/bin/bash
# 2-GPU visit: peer-memory SyncBN test, data-parallel bench (peer vs NCCL exchanges)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ddp.py -x -q > gpurun_out/pytest_ddp.log 2>&1; echo "pytest ddp rc=$?"
tail -4 gpurun_out/pytest_ddp.log
for mode in peer nccl; do
MMH_SYNCBN=$mode timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2_$mode.json 2> gpurun_out/bench_n2_$mode.err; echo "bench n2 $mode rc=$?"
cat gpurun_out/bench_n2_$mode.json | cut -c1-330; grep -v Warning gpurun_out/bench_n2_$mode.err | grep -v "^\*\*\*\|OMP_NUM" | tail -5
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_same_box.json 2> gpurun_out/bench_n1.err; echo "bench n1 rc=$?"
cat gpurun_out/bench_n1_same_box.json | cut -c1-330
