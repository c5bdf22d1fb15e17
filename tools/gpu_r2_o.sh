#!/bin/bash
# Round 2, visit O (1 GPU): the default bench line with the bandwidth-bound rooflines (bn_bwd, norm_act), contract check.
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_o_default.json 2> gpurun_out/bench_o_default.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_o_default.json').read().strip().splitlines()[-1])
print('value %.1f ms %.2f e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']))
for r in [d['roofline']] + d['rooflines_other']: print(' ', r['kernel'][:50], round(r['achieved'],1), r['unit'], 'frac', round(r['frac'],3), 'ms', round(r['kernel_ms_per_step'],2), 'launches', r['launches_per_step'])
print(' secondary infer e2e', d['secondary']['infer_configs1']['e2e']['value'])
" || tail -5 gpurun_out/bench_o_default.err
