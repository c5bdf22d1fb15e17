#!/bin/bash
# 2-GPU visit: peer-memory SyncBN test, data-parallel bench (peer vs NCCL exchanges), rasteriser bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ddp.py -x -q > gpurun_out/pytest_ddp.log 2>&1; echo "pytest ddp rc=$?"
tail -15 gpurun_out/pytest_ddp.log
for mode in peer nccl; do
MMH_SYNCBN=$mode timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2_$mode.json 2> gpurun_out/bench_n2_$mode.err; echo "bench n2 $mode rc=$?"
cat gpurun_out/bench_n2_$mode.json | cut -c1-400; grep -v Warning gpurun_out/bench_n2_$mode.err | tail -5
done
timeout 300 python bench.py --workload raster --no-cpu-baseline > gpurun_out/bench_raster.json 2> gpurun_out/bench_raster.err; echo "raster rc=$?"
cat gpurun_out/bench_raster.json | cut -c1-300; tail -3 gpurun_out/bench_raster.err
timeout 200 python -m pytest tests/test_gpu_model.py -x -q -k "raster or known" 2>&1 | tail -3
