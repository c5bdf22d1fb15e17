#!/bin/bash
# Round 2, visit Q (1 GPU): the configs[4] shape (batch 64 per GPU) on one GPU.
mkdir -p gpurun_out
timeout 500 python bench.py --batch 64 --steps 6 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/bench_q_b64.json 2> gpurun_out/bench_q_b64.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_q_b64.json').read().strip().splitlines()[-1])
print('B=64 value %.1f ms %.2f e2e %.1f util %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['step_tensor_util']))
for r in [d['roofline']] + d['rooflines_other']: print(' ', r['kernel'][:50], round(r['achieved'],1), r['unit'], 'frac', round(r['frac'],3), 'ms', round(r['kernel_ms_per_step'],2))
" || tail -5 gpurun_out/bench_q_b64.err
