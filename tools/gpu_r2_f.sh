#!/bin/bash
# Round 2, visit F (1 GPU): Adam-moment race fix (regression test + model file), layer chains on CUDA streams A/B.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -rP > gpurun_out/pytest_model_full.log 2>&1; echo "model file rc=$?"
grep -E "passed|failed|worst relative loss|^FAILED" gpurun_out/pytest_model_full.log | tail -8
grep -E "^E  " gpurun_out/pytest_model_full.log | head
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --no-secondary > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; rc=$?
  python -c "
import json; d=json.load(open('gpurun_out/bench_$name.json')); print('%-28s rc=$rc value %.1f ms %.2f e2e %.1f conv %.3f (%.2f ms) wgrad %.3f (%.2f ms)' % ('$name', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'], d['rooflines_other'][0]['frac'], d['rooflines_other'][0]['kernel_ms_per_step']))" || tail -3 gpurun_out/bench_$name.err
}
run chains0_fuse0 MMH_PAT_STREAMS=0 MMH_FUSE_BN_BWD=0
run chains1_fuse0 MMH_PAT_STREAMS=1 MMH_FUSE_BN_BWD=0
run chains1_fuse1 MMH_PAT_STREAMS=1 MMH_FUSE_BN_BWD=1
run chains1_fuse0_smem208 MMH_PAT_STREAMS=1 MMH_FUSE_BN_BWD=0 MMH_CONV_SMEM_KB=208
run chains1_fuse1_smem208 MMH_PAT_STREAMS=1 MMH_FUSE_BN_BWD=1 MMH_CONV_SMEM_KB=208
run chains1_fuse0_nopdl MMH_PAT_STREAMS=1 MMH_FUSE_BN_BWD=0 MMH_PDL=0
MMH_PAT_STREAMS=1 timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -k "train_losses or generator_train" > gpurun_out/pytest_chains.log 2>&1; echo "chains pytest rc=$?"
grep -E "passed|failed" gpurun_out/pytest_chains.log | tail -2
