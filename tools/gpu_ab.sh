#!/bin/bash
# A/B of an environment switch on the training bench + parity tests with the default setting
mkdir -p gpurun_out
for v in 1 0; do
MMH_WGRAD_STREAM=$v timeout 300 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/bench_ws$v.json 2> gpurun_out/bench_ws$v.err; echo "bench WGRAD_STREAM=$v rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_ws$v.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['step_tensor_util'])"
grep -v Warn gpurun_out/bench_ws$v.err | tail -3
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
