#!/bin/bash
# Round 2, visit L (1 GPU): norm_act launch-shape knobs (vectors in flight, blocks per SM) isolated and in the step;
# ncu full capture of the lean BatchNorm-backward kernels; infer bench after the roofline-timing fix.
mkdir -p gpurun_out
for cfg in "4 8" "8 8" "2 8" "4 16" "4 4" "8 16"; do
  set -- $cfg
  echo "== unroll=$1 per_sm=$2"; MMH_NORM_UNROLL=$1 MMH_PG_PER_SM=$2 timeout 200 python tools/exp/ew_bench.py --iters 20 2>&1 | grep " general " | sed 's/.*| pair/pair/'
done | tee gpurun_out/norm_knobs.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --no-secondary > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; rc=$?
  python -c "
import json; d=json.load(open('gpurun_out/bench_$name.json')); print('%-22s rc=$rc value %.1f ms %.2f e2e %.1f conv %.3f (%.2f ms) wgrad %.3f (%.2f ms)' % ('$name', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'], d['rooflines_other'][0]['frac'], d['rooflines_other'][0]['kernel_ms_per_step']))" || tail -3 gpurun_out/bench_$name.err
}
run base
run norm_u8 MMH_NORM_UNROLL=8
run pg16 MMH_PG_PER_SM=16
run norm_u8_pg16 MMH_NORM_UNROLL=8 MMH_PG_PER_SM=16
run base2
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'rows_pg_kernel|rows_reduce_fin_kernel' -s 20 -c 8 -f -o /tmp/r02_bn_lean_ncu_full python tools/profile_step.py > gpurun_out/ncu_lean.log 2>&1; echo "ncu lean rc=$?"
python tools/ncu_summary.py /tmp/r02_bn_lean_ncu_full.ncu-rep > gpurun_out/r02_bn_lean_ncu_full.txt 2>gpurun_out/r02_bn_lean.err; wc -l gpurun_out/r02_bn_lean_ncu_full.txt
timeout 300 python bench.py --workload infer --no-cpu-baseline > gpurun_out/bench_l_infer.json 2> gpurun_out/bench_l_infer.err; echo "infer rc=$?"; cut -c1-200 gpurun_out/bench_l_infer.json
