#!/bin/bash
# Round 2, visit E (1 GPU): guard-gap hunt for the intermittent NaN / deviation, fused-epilogue experiments + source profile.
mkdir -p gpurun_out
timeout 600 python tests/diag_guard.py > gpurun_out/diag_guard.log 2>&1; echo "guard rc=$?"
grep -E "^model|^    |non-finite" gpurun_out/diag_guard.log | head -60
MMH_ARENA=0 MMH_ARENA_GUARD=0 DIAG_MODELS=6 timeout 600 python tests/diag_guard.py > gpurun_out/diag_noarena.log 2>&1; echo "noarena rc=$?"
grep -E "^model" gpurun_out/diag_noarena.log | grep -v "finite=True" | head; grep -c "finite=True" gpurun_out/diag_noarena.log
MMH_FUSE_BN_BWD=0 MMH_ARENA_GUARD=0 DIAG_MODELS=6 timeout 600 python tests/diag_guard.py > gpurun_out/diag_nofuse.log 2>&1; echo "nofuse rc=$?"
grep -E "^model" gpurun_out/diag_nofuse.log | grep -v "finite=True" | head; grep -c "finite=True" gpurun_out/diag_nofuse.log
MMH_PDL=0 MMH_ARENA_GUARD=0 DIAG_MODELS=6 timeout 600 python tests/diag_guard.py > gpurun_out/diag_nopdl.log 2>&1; echo "nopdl rc=$?"
grep -E "^model" gpurun_out/diag_nopdl.log | grep -v "finite=True" | head; grep -c "finite=True" gpurun_out/diag_nopdl.log
MMH_WGRAD_STREAM=0 MMH_ARENA_GUARD=0 DIAG_MODELS=6 timeout 600 python tests/diag_guard.py > gpurun_out/diag_nostream.log 2>&1; echo "nostream rc=$?"
grep -E "^model" gpurun_out/diag_nostream.log | grep -v "finite=True" | head; grep -c "finite=True" gpurun_out/diag_nostream.log
MMH_C2_DEBUG=16 timeout 200 python tools/exp/bs_bench.py --only "dgrad" > gpurun_out/bs_bench_nox.log 2>&1; echo "bs_bench nox rc=$?"
grep -v Warn gpurun_out/bs_bench_nox.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv2_kernel -s 2 -c 1 -f -o gpurun_out/bs_fused \
  python tools/exp/bs_bench.py --iters 1 --only "dgrad fused" > gpurun_out/ncu_bs.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep 2>/dev/null; true
