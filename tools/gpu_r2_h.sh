#!/bin/bash
# Round 2, visit H (1 GPU): two epilogue warp groups (setmaxnreg) -- conv parity, then A/B against the one-group build,
# fused statistics everywhere / fused BN-backward on top, elementwise occupancy variant.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -q > gpurun_out/pytest_conv.log 2>&1; echo "conv pytest rc=$?"
grep -E "passed|failed" gpurun_out/pytest_conv.log | tail -2
grep -E "^FAILED|AssertionError: \{" gpurun_out/pytest_conv.log | head -20
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --no-secondary > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; rc=$?
  python -c "
import json; d=json.load(open('gpurun_out/bench_$name.json')); print('%-22s rc=$rc value %.1f ms %.2f e2e %.1f conv %.3f (%.2f ms) wgrad %.3f (%.2f ms)' % ('$name', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'], d['rooflines_other'][0]['frac'], d['rooflines_other'][0]['kernel_ms_per_step']))" || tail -3 gpurun_out/bench_$name.err
}
run epi1 MMH_LIB_PATH=$PWD/mmhand_b200/libmmhand_sm100_epi1.so
run epi2
run epi2_stats1 MMH_CONV_STATS=1
run epi2_fuse MMH_FUSE_BN_BWD=1
run epi2_stats1_fuse MMH_CONV_STATS=1 MMH_FUSE_BN_BWD=1
run epi2_mb3 MMH_LIB_PATH=$PWD/mmhand_b200/libmmhand_sm100_mb3.so
run epi2_mb3_stats1_fuse MMH_LIB_PATH=$PWD/mmhand_b200/libmmhand_sm100_mb3.so MMH_CONV_STATS=1 MMH_FUSE_BN_BWD=1
timeout 200 python tools/exp/bs_bench.py > gpurun_out/bs_bench_epi2.log 2>&1; grep -v Warn gpurun_out/bs_bench_epi2.log
PYTHONPATH=$PWD:$PWD/tests timeout 200 python -c "
import conv_cases as c
for n in ('perf','perf512','perf_stem','perf_stem42','perf_out','perf_d1'):
    r=c.CASES[n](); print(n, {k: round(v,1) for k,v in r.items() if k.endswith('tflops')})
" 2>&1 | grep -v Warn | tee gpurun_out/perf_cases_epi2.log
MMH_LIB_PATH=$PWD/mmhand_b200/libmmhand_sm100_epi1.so PYTHONPATH=$PWD:$PWD/tests timeout 200 python -c "
import conv_cases as c
for n in ('perf','perf512','perf_stem','perf_stem42','perf_out','perf_d1'):
    r=c.CASES[n](); print(n, {k: round(v,1) for k,v in r.items() if k.endswith('tflops')})
" 2>&1 | grep -v Warn | tee gpurun_out/perf_cases_epi1.log
