#!/bin/bash
# Round 2, visit A (1 GPU): new parity tests with printed figures, bench default vs async-input, smoke launch list.
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.json
timeout 1500 python -m pytest tests -m gpu -x -q -rP > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|error" gpurun_out/pytest_gpu.log | tail -3
grep -E "max-abs|worst relative|D train|cosine" gpurun_out/pytest_gpu.log | tail -20
cat gpurun_out/parity_metrics.json 2>/dev/null | head -40
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
MMH_ASYNC_INPUT=1 timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench_async.json 2> gpurun_out/bench_async.err; echo "bench async rc=$?"
cat gpurun_out/bench_async.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/smoke_launches.csv \
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_ncu.log 2>&1; echo "smoke ncu rc=$?"
tail -2 gpurun_out/smoke_ncu.log
python tools/summarize_launches.py gpurun_out/smoke_launches.csv 2>/dev/null | head -30
