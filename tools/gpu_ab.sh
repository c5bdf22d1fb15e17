#!/bin/bash
mkdir -p gpurun_out
for v in 1 0; do
echo "=== MMH_CONV_STATS=$v (serial: MMH_WGRAD_STREAM=0)"
MMH_WGRAD_STREAM=0 MMH_CONV_STATS=$v timeout 300 python tools/layer_times.py 2>/dev/null | grep " fwd " | head -24
MMH_WGRAD_STREAM=0 MMH_CONV_STATS=$v timeout 300 python tools/class_times.py 2>&1 | grep -v Warn | head -12
done
