#!/bin/bash
# Round 2, 8-GPU visit: the default multi-GPU configuration (peer-memory SyncBN, PDL, generator update stream, bucketed
# gradient all-reduce) with a short watchdog; the conservative ladder only if it fails; NCCL SyncBN for comparison.
N=${1:-8}
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env MMH_BENCH_WATCHDOG_S=70 "$@" timeout 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29400 + RANDOM % 200)) \
    bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/bench_n${N}_$name.json 2> gpurun_out/bench_n${N}_$name.err; rc=$?
  python -c "
import json
d=json.loads(open('gpurun_out/bench_n${N}_$name.json').read().strip().splitlines()[-1]); c=d.get('config',{})
print('%-26s rc=$rc value %.1f ms %.2f e2e %.1f syncbn=%s pdl=%s grads=%s' % ('$name', d['value'], d['ms_per_step'], d['e2e']['value'], c.get('syncbn'), c.get('pdl'), c.get('grad_allreduce')))" 2>/dev/null || { echo "$name rc=$rc FAILED"; tail -4 gpurun_out/bench_n${N}_$name.err; return 1; }
}
if run default; then
  run nccl_syncbn MMH_SYNCBN=nccl
else
  run peer_pdl0 MMH_SYNCBN=peer MMH_PDL=0 MMH_G_UPDATE_STREAM=0 || run nccl_base MMH_SYNCBN=nccl MMH_PDL=0 MMH_G_UPDATE_STREAM=0 MMH_GRAD_BUCKETS=0
fi
