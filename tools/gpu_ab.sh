#!/bin/bash
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/bench_ab_$tag.json 2> gpurun_out/bench_ab_$tag.err; echo "bench $tag ($*) rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_ab_$tag.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['step_tensor_util'], d['roofline']['achieved'])"
  grep -v Warn gpurun_out/bench_ab_$tag.err | tail -3
}
run pdl1 MMH_PDL=1
run pdl0 MMH_PDL=0
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --workload infer --no-cpu-baseline > gpurun_out/bench_infer.json 2>/dev/null; cut -c1-200 gpurun_out/bench_infer.json
