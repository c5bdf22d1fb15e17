#!/bin/bash
# Round 2, N-GPU bench with the final defaults (one run).
N=${1:-4}
mkdir -p gpurun_out
env MMH_BENCH_WATCHDOG_S=100 timeout 160 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29400 + RANDOM % 200)) \
  bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/bench_n${N}_final.json 2> gpurun_out/bench_n${N}_final.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_n${N}_final.json').read().strip().splitlines()[-1]); c=d.get('config',{})
print('N=$N value %.1f ms %.2f e2e %.1f syncbn=%s pdl=%s gup=%s grads=%s' % (d['value'], d['ms_per_step'], d['e2e']['value'], c.get('syncbn'), c.get('pdl'), c.get('g_update_stream'), c.get('grad_allreduce')))" || tail -5 gpurun_out/bench_n${N}_final.err
