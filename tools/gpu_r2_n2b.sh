#!/bin/bash
# Round 2, 2-GPU visit: data-parallel tests against the oracle on the joint batch with the final defaults (peer SyncBN,
# no dependent launches, generator update on the launch stream -- set through mmh_set_pdl, not the environment), bench.
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ddp.py -m gpu -q -rP -x > gpurun_out/pytest_ddp_n${N}b.log 2>&1; echo "ddp pytest rc=$?"
grep -E "passed|failed|skipped|joint-batch" gpurun_out/pytest_ddp_n${N}b.log | tail -8
grep -E "^E  " gpurun_out/pytest_ddp_n${N}b.log | head -10
env MMH_BENCH_WATCHDOG_S=100 timeout 160 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29400 + RANDOM % 200)) \
  bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/bench_n${N}_final.json 2> gpurun_out/bench_n${N}_final.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_n${N}_final.json').read().strip().splitlines()[-1]); c=d.get('config',{})
print('N=$N value %.1f ms %.2f e2e %.1f syncbn=%s pdl=%s gup=%s grads=%s' % (d['value'], d['ms_per_step'], d['e2e']['value'], c.get('syncbn'), c.get('pdl'), c.get('g_update_stream'), c.get('grad_allreduce')))" || tail -5 gpurun_out/bench_n${N}_final.err
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py --workload infer --no-cpu-baseline > gpurun_out/bench_n2b_infer.json 2> gpurun_out/bench_n2b_infer.err; echo "infer rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_n2b_infer.json')); print('infer value %.1f (%.2f ms) e2e %.1f (%.2f ms) conv frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac']))"
