#!/bin/bash
# Round 2, 2-GPU visit: which of this round's changes stalls the data-parallel step? Small rungs, 75 s watchdogs with a
# Python stack dump of every rank.
N=2
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env MMH_BENCH_WATCHDOG_S=${WD:-75} "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29400 + RANDOM % 200)) \
    bench.py --gpus $N --steps ${STEPS:-6} --warmup 3 --no-cpu-baseline > gpurun_out/bench_n${N}_$name.json 2> gpurun_out/bench_n${N}_$name.err; rc=$?
  python -c "
import json
d=json.loads(open('gpurun_out/bench_n${N}_$name.json').read().strip().splitlines()[-1]); c=d.get('config',{})
print('%-26s rc=$rc value %.1f ms %.2f e2e %.1f syncbn=%s pdl=%s chains=%s grads=%s' % ('$name', d['value'], d['ms_per_step'], d['e2e']['value'], c.get('syncbn'), c.get('pdl'), c.get('layer_chain_streams'), c.get('grad_allreduce')))" 2>/dev/null || { echo "$name rc=$rc FAILED"; grep -E "File \"/tmp/code|watchdog|Error" gpurun_out/bench_n${N}_$name.err | head -24; }
}
run nccl_nochain_nobucket MMH_SYNCBN=nccl MMH_PDL=0 MMH_G_UPDATE_STREAM=0 MMH_GRAD_BUCKETS=0
run nccl_nochain_bucket   MMH_SYNCBN=nccl MMH_PDL=0 MMH_G_UPDATE_STREAM=0 MMH_GRAD_BUCKETS=1
run peer_nochain          MMH_SYNCBN=peer MMH_PDL=0 MMH_GRAD_BUCKETS=1
run peer_nochain_pdl      MMH_SYNCBN=peer MMH_PDL=1 MMH_GRAD_BUCKETS=1
run nccl_chain            MMH_SYNCBN=nccl MMH_PDL=0 MMH_GRAD_BUCKETS=0 MMH_PAT_STREAMS_DP=1
timeout 600 python -m pytest tests/test_gpu_ddp.py -m gpu -q -rP -x > gpurun_out/pytest_ddp_n$N.log 2>&1; echo "ddp pytest rc=$?"
grep -E "passed|failed|skipped|joint-batch" gpurun_out/pytest_ddp_n$N.log | tail -8
grep -E "^E  " gpurun_out/pytest_ddp_n$N.log | head -10
