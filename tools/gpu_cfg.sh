#!/bin/bash
# GPU visit: secondary workloads (configs[1] inference, configs[3] rasteriser)
mkdir -p gpurun_out
timeout 300 python bench.py --workload infer > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err; echo "infer rc=$?"
cat gpurun_out/bench_infer.json; tail -3 gpurun_out/bench_infer.err
timeout 300 python bench.py --workload raster > gpurun_out/bench_raster.json 2> gpurun_out/bench_raster.err; echo "raster rc=$?"
cat gpurun_out/bench_raster.json; tail -3 gpurun_out/bench_raster.err
