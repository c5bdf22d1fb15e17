#!/bin/bash
# Round 2, visit P (1 GPU): final sanity of the shipped tree -- GPU suite, smoke, default bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_p.log 2>&1; echo "gpu pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_p.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_p_default.json 2> gpurun_out/bench_p_default.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_p_default.json').read().strip().splitlines()[-1])
print('value %.1f ms %.2f e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']))
for r in [d['roofline']] + d['rooflines_other']: print(' ', r['kernel'][:50], round(r['achieved'],1), r['unit'], 'frac', round(r['frac'],3))
print(' secondary infer', d['secondary']['infer_configs1']['value'], d['secondary']['infer_configs1']['e2e']['value'])
" || tail -5 gpurun_out/bench_p_default.err
