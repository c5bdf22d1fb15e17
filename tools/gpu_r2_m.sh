#!/bin/bash
# Round 2, visit M (1 GPU): taped inference (parity + infer bench), the GPU suite on the final library, default bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_m.log 2>&1; echo "gpu pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_m.log
for t in 1 0; do
  MMH_INFER_TAPE=$t timeout 300 python bench.py --workload infer --no-cpu-baseline > gpurun_out/bench_m_infer_tape$t.json 2> gpurun_out/bench_m_infer_tape$t.err; echo "infer tape=$t rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_m_infer_tape$t.json')); print('tape=$t value %.1f (%.2f ms) e2e %.1f (%.2f ms) conv frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac']))"
done
timeout 600 python bench.py > gpurun_out/bench_m_default.json 2> gpurun_out/bench_m_default.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_m_default.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_m_reference.json 2> gpurun_out/bench_m_reference.err; echo "reference rc=$?"; cut -c1-400 gpurun_out/bench_m_reference.json
