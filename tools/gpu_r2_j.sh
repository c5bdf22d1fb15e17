#!/bin/bash
# Round 2, visit J (1 GPU): stem weight gradients with the kernel rows on M (wgrad2 mode 2) -- parity, timing against
# mode 1, step A/B; then the round's ncu evidence (launch list of one step, full captures of the top kernels).
mkdir -p gpurun_out
cases="wgrad_s1_7x7_c3 wgrad_s1_7x7_c42 wgrad_s1_7x7_out3 wgrad_s1_c32 big_wgrad_stem_c3 big_wgrad_stem_c42 big_wgrad_stem_c24 big_wgrad_out3"
for m in 2; do
  for c in $cases; do
    echo -n "mode=$m $c: "; MMH_WGRAD_MODE=$m timeout 120 python tools/bringup.py --one $c 2>&1 | grep -E "RESULT|rror" | cut -c1-260 | tail -1
  done
done | tee gpurun_out/w2_mode2_parity.log
for cfg in "1 1" "2 1" "2 0" "2 2"; do
  set -- $cfg
  echo -n "mode=$1 skew=$2 "; MMH_WGRAD_MODE=$1 MMH_WGRAD_SKEW=$2 PYTHONPATH=$PWD:$PWD/tests timeout 120 python -c "
import conv_cases as c
for n in ('perf_stem','perf_stem42','perf_out'):
    r=c.CASES[n](); print(n, {k: round(v,3) for k,v in r.items() if 'wgrad' in k}, end=' ')
print()
" 2>&1 | grep -v Warn
done | tee gpurun_out/w2_mode2_perf.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --no-secondary > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; rc=$?
  python -c "
import json; d=json.load(open('gpurun_out/bench_$name.json')); print('%-22s rc=$rc value %.1f ms %.2f e2e %.1f conv %.3f (%.2f ms) wgrad %.3f (%.2f ms)' % ('$name', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'], d['rooflines_other'][0]['frac'], d['rooflines_other'][0]['kernel_ms_per_step']))" || tail -3 gpurun_out/bench_$name.err
}
run wmode1 MMH_WGRAD_MODE=1
run wmode2 MMH_WGRAD_MODE=2
run wmode2_nochain MMH_WGRAD_MODE=2 MMH_PAT_STREAMS=0
run wmode2_nowstream MMH_WGRAD_MODE=2 MMH_WGRAD_STREAM=0
# ---- ncu evidence of the default configuration
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/r02_launches.csv python tools/profile_step.py > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'conv2_kernel' -s 40 -c 6 -f -o gpurun_out/r02_conv2_full python tools/profile_step.py > gpurun_out/ncu_conv2.log 2>&1; echo "ncu conv2 rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'rows_pg_kernel|rows_reduce_fin_kernel|pg_kernel' -s 200 -c 12 -f -o gpurun_out/r02_ew_full python tools/profile_step.py > gpurun_out/ncu_ew.log 2>&1; echo "ncu ew rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'wgrad2_kernel' -s 90 -c 18 -f -o gpurun_out/r02_wgrad_full python tools/profile_step.py > gpurun_out/ncu_wgrad.log 2>&1; echo "ncu wgrad rc=$?"
ls -la gpurun_out/*.ncu-rep
