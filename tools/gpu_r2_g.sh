#!/bin/bash
# Round 2, visit G (1 GPU): whole GPU suite, elementwise occupancy A/B (register cap 80 vs 128), stem wgrad profile.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|error" gpurun_out/pytest_gpu.log | tail -3
grep -E "^E  " gpurun_out/pytest_gpu.log | head -10
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --no-secondary > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; rc=$?
  python -c "
import json; d=json.load(open('gpurun_out/bench_$name.json')); print('%-22s rc=$rc value %.1f ms %.2f e2e %.1f conv %.3f (%.2f ms) wgrad %.3f (%.2f ms)' % ('$name', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'], d['rooflines_other'][0]['frac'], d['rooflines_other'][0]['kernel_ms_per_step']))" || tail -3 gpurun_out/bench_$name.err
}
run base
run mb3 MMH_LIB_PATH=$PWD/mmhand_b200/libmmhand_sm100_mb3.so
run mb3_wave6 MMH_LIB_PATH=$PWD/mmhand_b200/libmmhand_sm100_mb3.so MMH_REDUCE_WAVE=6
run mb3_nochain MMH_LIB_PATH=$PWD/mmhand_b200/libmmhand_sm100_mb3.so MMH_PAT_STREAMS=0
MMH_PERF_ITERS=1 PYTHONPATH=$PWD:$PWD/tests timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad2_kernel -c 2 -f -o gpurun_out/stem_wgrad \
  python -c "import conv_cases as c; print(c.CASES['perf_stem']()); print(c.CASES['perf_stem42']())" > gpurun_out/ncu_stem.log 2>&1; echo "ncu stem rc=$?"
tail -3 gpurun_out/ncu_stem.log
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "full bench rc=$?"
cut -c1-1500 gpurun_out/bench_full.json
