#!/bin/bash
# Round 2, visit I (1 GPU): lean row kernels for BN backward / norm_act -- parity, isolated timing, step A/B;
# stem weight-gradient ablation (MMH_W2_DEBUG); full GPU suite.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_bn_lean.py -q > gpurun_out/pytest_lean.log 2>&1; echo "lean pytest rc=$?"; tail -2 gpurun_out/pytest_lean.log
timeout 300 python tools/exp/ew_bench.py > gpurun_out/ew_bench.log 2>&1; echo "ew_bench rc=$?"; grep -v Warn gpurun_out/ew_bench.log | tail -30
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --no-secondary > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; rc=$?
  python -c "
import json; d=json.load(open('gpurun_out/bench_$name.json')); print('%-22s rc=$rc value %.1f ms %.2f e2e %.1f conv %.3f (%.2f ms) wgrad %.3f (%.2f ms)' % ('$name', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'], d['rooflines_other'][0]['frac'], d['rooflines_other'][0]['kernel_ms_per_step']))" || tail -3 gpurun_out/bench_$name.err
}
run general MMH_EW_LEAN=0
run lean MMH_ROWS_PF=0 MMH_EW_REVERSE=0
run lean_rev MMH_ROWS_PF=0 MMH_EW_REVERSE=1
run lean_pf MMH_ROWS_PF=1 MMH_EW_REVERSE=0
run lean_pf_rev MMH_ROWS_PF=1 MMH_EW_REVERSE=1
for d in 0 1 2 4 6; do
  echo -n "w2dbg=$d "; MMH_W2_DEBUG=$d PYTHONPATH=$PWD:$PWD/tests timeout 120 python -c "
import conv_cases as c
for n in ('perf_stem','perf_stem42','perf_out'):
    r=c.CASES[n](); print(n, {k: round(v,1) for k,v in r.items() if 'wgrad' in k}, end=' ')
print()
" 2>&1 | grep -v Warn
done | tee gpurun_out/w2dbg.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_i.log 2>&1; echo "gpu pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_i.log
