#!/bin/bash
# Round 2, visit C (1 GPU): batch-16 loss test in isolation and after the batch-1 case, fused epilogue with deeper
# prefetch (A/B), compact-input e2e, whole GPU suite.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -rP -k "12-16-256" > gpurun_out/pytest_b16_alone.log 2>&1; echo "b16 alone rc=$?"
grep -E "passed|failed|worst relative|^step [0-2] mine" gpurun_out/pytest_b16_alone.log | tail -8
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -rP -k "train_losses" > gpurun_out/pytest_b16_after_b1.log 2>&1; echo "b1 then b16 rc=$?"
grep -E "passed|failed|worst relative|^step [0-2] mine" gpurun_out/pytest_b16_after_b1.log | tail -14
for f in 0 1; do
  MMH_FUSE_BN_BWD=$f timeout 300 python bench.py --no-cpu-baseline --no-secondary > gpurun_out/bench_fuse$f.json 2> gpurun_out/bench_fuse$f.err; echo "bench fuse=$f rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_fuse$f.json')); print('fuse=$f', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['h2d_bytes_per_step'], 'conv', d['roofline']['frac'], d['roofline']['kernel_ms_per_step'], 'wgrad', d['rooflines_other'][0]['frac'], d['rooflines_other'][0]['kernel_ms_per_step'])"
  tail -3 gpurun_out/bench_fuse$f.err
done
timeout 1500 python -m pytest tests -m gpu -q -x --deselect "tests/test_gpu_model.py::test_train_losses_match_oracle" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|error" gpurun_out/pytest_gpu.log | tail -3
grep -E "^E " gpurun_out/pytest_gpu.log | head -20
