#!/bin/bash
# Round 2, visit K (1 GPU): the round's ncu evidence of the default configuration -- launch list of one training step,
# full captures of the top kernels summarised ON THE BOX (the reports exceed what comes back), GPU suite, default bench.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/r02_launches.csv python tools/profile_step.py > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
python tools/summarize_launches.py gpurun_out/r02_launches.csv > gpurun_out/r02_step_launch_summary.txt; head -30 gpurun_out/r02_step_launch_summary.txt
cap() { # name, kernel regex, skip, count
  timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"$2" -s $3 -c $4 -f -o /tmp/$1 python tools/profile_step.py > gpurun_out/ncu_$1.log 2>&1; echo "ncu $1 rc=$?"
  python tools/ncu_summary.py /tmp/$1.ncu-rep > gpurun_out/$1.txt 2>gpurun_out/$1.err; wc -l gpurun_out/$1.txt
}
cap r02_conv2_ncu_full 'conv2_kernel' 40 6
cap r02_ew_ncu_full 'rows_pg_kernel|rows_reduce_fin_kernel|pg_kernel' 200 12
cap r02_wgrad2_ncu_full 'wgrad2_kernel' 90 18
cp /tmp/r02_conv2_ncu_full.ncu-rep gpurun_out/ 2>/dev/null
du -sh gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_k.log 2>&1; echo "gpu pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_k.log
timeout 600 python bench.py > gpurun_out/bench_k_default.json 2> gpurun_out/bench_k_default.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_k_default.json
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --no-secondary > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; rc=$?
  python -c "
import json; d=json.load(open('gpurun_out/bench_$name.json')); print('%-22s rc=$rc value %.1f ms %.2f e2e %.1f conv %.3f (%.2f ms) wgrad %.3f (%.2f ms)' % ('$name', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'], d['rooflines_other'][0]['frac'], d['rooflines_other'][0]['kernel_ms_per_step']))" || tail -3 gpurun_out/bench_$name.err
}
# what-if runs (timing only, wrong results): where does the step time go?
run whatif_no_wgrad_mma MMH_W2_DEBUG=2
run whatif_no_conv_mma MMH_C2_DEBUG=4
run whatif_no_mma MMH_W2_DEBUG=2 MMH_C2_DEBUG=4
run whatif_no_mma_no_loads MMH_W2_DEBUG=6 MMH_C2_DEBUG=7
timeout 300 python bench.py --workload infer --no-cpu-baseline > gpurun_out/bench_k_infer.json 2> gpurun_out/bench_k_infer.err; echo "infer rc=$?"; cut -c1-400 gpurun_out/bench_k_infer.json
