#!/bin/bash
# Round 2: configs[4] -- batch 64 per GPU, data parallel over N GPUs (one run, final defaults).
N=${1:-8}
mkdir -p gpurun_out
env MMH_BENCH_WATCHDOG_S=150 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29400 + RANDOM % 200)) \
  bench.py --gpus $N --batch 64 --steps 6 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/bench_n${N}_b64.json 2> gpurun_out/bench_n${N}_b64.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_n${N}_b64.json').read().strip().splitlines()[-1]); c=d.get('config',{})
print('N=$N B=64 value %.1f ms %.2f e2e %.1f util %.3f syncbn=%s pdl=%s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['step_tensor_util'], c.get('syncbn'), c.get('pdl')))" || tail -5 gpurun_out/bench_n${N}_b64.err
