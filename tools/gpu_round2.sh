#!/bin/bash
# GPU-box visit (lean): parity tests, bench, ncu launch list of one step, full captures of the top bandwidth kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'pg_kernel|reduce_ch_kernel' -s 300 -c 24 -f -o gpurun_out/ew_full python tools/profile_step.py > gpurun_out/ncu_ew.log 2>&1; echo "ncu ew rc=$?"
timeout 200 python tools/layer_times.py > gpurun_out/layer_times.txt 2>&1; head -60 gpurun_out/layer_times.txt
