#!/bin/bash
# Round 2, visit B (1 GPU): full-size conv cases, batch-16 diagnosis, fused BN-backward epilogue A/B, async-input probe.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -q -rf > gpurun_out/pytest_conv.log 2>&1; echo "conv pytest rc=$?"
grep -E "passed|failed" gpurun_out/pytest_conv.log | tail -2
grep -E "^FAILED|AssertionError: \{" gpurun_out/pytest_conv.log | head -40
MMH_FUSE_BN_BWD=0 timeout 600 python tests/diag_b16.py grads steps > gpurun_out/diag_b16_plain.log 2>&1; echo "diag plain rc=$?"
grep -v Warning gpurun_out/diag_b16_plain.log | tail -60
timeout 400 python tests/diag_b16.py steps > gpurun_out/diag_b16_fused.log 2>&1; echo "diag fused rc=$?"
grep -E "^step|weight" gpurun_out/diag_b16_fused.log | tail -12
for f in 0 1; do
  MMH_FUSE_BN_BWD=$f timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_fuse$f.json 2> gpurun_out/bench_fuse$f.err; echo "bench fuse=$f rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_fuse$f.json')); print('fuse=$f', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'])"
done
MMH_ASYNC_INPUT=1 MMH_PDL=0 timeout 120 python bench.py --no-cpu-baseline --steps 5 > gpurun_out/bench_async_pdl0.json 2> gpurun_out/bench_async_pdl0.err; echo "async pdl0 rc=$?"
cat gpurun_out/bench_async_pdl0.json | cut -c1-400
