#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -x -q -k "jointsmap or raster" > gpurun_out/pytest_jm.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_jm.log
timeout 300 python bench.py --workload jointsmap > gpurun_out/bench_jointsmap.json 2> gpurun_out/bench_jointsmap.err; echo "jm rc=$?"
cat gpurun_out/bench_jointsmap.json; tail -3 gpurun_out/bench_jointsmap.err
