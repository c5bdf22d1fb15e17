#!/bin/bash
# GPU-box visit: parity tests, bench, per-entry-point times
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
timeout 300 python tools/class_times.py > gpurun_out/class_times.txt 2>&1; grep -v Warn gpurun_out/class_times.txt | tail -30
