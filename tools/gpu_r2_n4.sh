#!/bin/bash
# Round 2, multi-GPU visit (N = number of GPUs of the box, 4 or 8): joint-batch-vs-oracle tests, then the bring-up
# ladder of the peer-memory SyncBN exchange and the bucketed gradient all-reduce. Every rung has its own watchdog.
N=${1:-4}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_ddp.py -m gpu -q -rP -x > gpurun_out/pytest_ddp_n$N.log 2>&1; echo "ddp pytest rc=$?"
grep -E "passed|failed|skipped|joint-batch" gpurun_out/pytest_ddp_n$N.log | tail -8
grep -E "^E  " gpurun_out/pytest_ddp_n$N.log | head -10
run() { # name, env...
  name=$1; shift
  env MMH_BENCH_WATCHDOG_S=100 "$@" timeout 160 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29400 + RANDOM % 200)) \
    bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n${N}_$name.json 2> gpurun_out/bench_n${N}_$name.err; rc=$?
  python -c "
import json
d=json.loads(open('gpurun_out/bench_n${N}_$name.json').read().strip().splitlines()[-1]); c=d.get('config',{})
print('%-26s rc=$rc value %.1f ms %.2f e2e %.1f syncbn=%s pdl=%s grads=%s' % ('$name', d['value'], d['ms_per_step'], d['e2e']['value'], c.get('syncbn'), c.get('pdl'), c.get('grad_allreduce')))" 2>/dev/null || { echo "$name rc=$rc FAILED"; tail -4 gpurun_out/bench_n${N}_$name.err; }
}
run nccl_base     MMH_SYNCBN=nccl MMH_PDL=0 MMH_G_UPDATE_STREAM=0 MMH_GRAD_BUCKETS=0
run nccl_buckets  MMH_SYNCBN=nccl MMH_PDL=0 MMH_G_UPDATE_STREAM=0 MMH_GRAD_BUCKETS=1
run peer_pdl0     MMH_SYNCBN=peer MMH_PDL=0 MMH_G_UPDATE_STREAM=0 MMH_GRAD_BUCKETS=1
run peer_pdl1     MMH_SYNCBN=peer MMH_PDL=1 MMH_G_UPDATE_STREAM=0 MMH_GRAD_BUCKETS=1
run peer_pdl1_gup MMH_SYNCBN=peer MMH_PDL=1 MMH_G_UPDATE_STREAM=1 MMH_GRAD_BUCKETS=1
run default
