#!/bin/bash
# One GPU-box visit: parity tests, bench, ncu launch list of one step, full captures of the two tensor-core kernels.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:conv2_kernel -s 40 -c 4 -f -o gpurun_out/conv2_full python tools/profile_step.py > gpurun_out/ncu_conv2.log 2>&1; echo "ncu conv2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:wgrad2_kernel -s 20 -c 3 -f -o gpurun_out/wgrad2_full python tools/profile_step.py > gpurun_out/ncu_wgrad2.log 2>&1; echo "ncu wgrad2 rc=$?"
timeout 300 python tools/layer_times.py > gpurun_out/layer_times.txt 2>&1; head -40 gpurun_out/layer_times.txt
