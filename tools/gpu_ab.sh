#!/bin/bash
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/bench_ab_$tag.json 2> gpurun_out/bench_ab_$tag.err; echo "bench $tag ($*) rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_ab_$tag.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['step_tensor_util'])"
  grep -v Warn gpurun_out/bench_ab_$tag.err | tail -3
}
run base MMH_EW_THREADS=128
run prio MMH_EW_THREADS=128 MMH_MAIN_PRIORITY=1
run prio_w2 MMH_EW_THREADS=128 MMH_MAIN_PRIORITY=1 MMH_WGRAD_WAVES=2
run prio_w4 MMH_EW_THREADS=128 MMH_MAIN_PRIORITY=1 MMH_WGRAD_WAVES=4
run w2 MMH_EW_THREADS=128 MMH_WGRAD_WAVES=2
