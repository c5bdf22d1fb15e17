#!/bin/bash
mkdir -p gpurun_out
timeout 120 tools/exp/ew_bw > gpurun_out/ew_bw.txt 2>&1; cat gpurun_out/ew_bw.txt
timeout 300 python tools/class_times.py > gpurun_out/class_times.txt 2>&1; grep -v Warn gpurun_out/class_times.txt | tail -40
