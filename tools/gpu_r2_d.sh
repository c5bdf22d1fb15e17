#!/bin/bash
# Round 2, visit D (1 GPU): the whole model test file in order, the compact-input test alone and under memcheck,
# isolated timing + source-level profile of the fused BN-backward epilogue.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -rP > gpurun_out/pytest_model_full.log 2>&1; echo "model file rc=$?"
grep -E "passed|failed|worst relative loss|^FAILED" gpurun_out/pytest_model_full.log | tail -8
timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -q -rP -k compact_forms > gpurun_out/pytest_compact_alone.log 2>&1; echo "compact alone rc=$?"
grep -E "passed|failed|^E  " gpurun_out/pytest_compact_alone.log | tail -6
timeout 900 compute-sanitizer --tool memcheck --print-limit 30 --error-exitcode 9 python -m pytest tests/test_gpu_model.py -m gpu -q -x -k compact_forms > gpurun_out/memcheck_compact.log 2>&1; echo "memcheck rc=$?"
grep -E "Invalid|ERROR SUMMARY|at .*\(|by thread|Address" gpurun_out/memcheck_compact.log | head -40
timeout 200 python tools/exp/bs_bench.py > gpurun_out/bs_bench.log 2>&1; echo "bs_bench rc=$?"
cat gpurun_out/bs_bench.log | grep -v Warn
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv2_kernel -s 6 -c 2 -f -o gpurun_out/bs_fused \
  python tools/exp/bs_bench.py --iters 1 --only "dgrad fused" > gpurun_out/ncu_bs.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep 2>/dev/null
