#!/bin/bash
# First single-GPU visit of the next round: A/B of the switches that were written after round 1's GPU budget ended
# (each line: value img/s, ms/step, e2e img/s, step tensor util), then the parity tests with the winners.
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/bench_ab_$tag.json 2> gpurun_out/bench_ab_$tag.err; echo "bench $tag ($*) rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_ab_$tag.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['step_tensor_util'])"
  grep -v Warn gpurun_out/bench_ab_$tag.err | tail -3
}
run base
run async_input MMH_ASYNC_INPUT=1
run reduce_wave3 MMH_REDUCE_WAVE=3
run async_wave3 MMH_ASYNC_INPUT=1 MMH_REDUCE_WAVE=3
MMH_ASYNC_INPUT=1 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_async.log 2>&1; echo "pytest (async input) rc=$?"
tail -3 gpurun_out/pytest_gpu_async.log
timeout 300 python - <<'PY'
# aug.py write-out on the device vs the host chain
import time, numpy as np, torch, cv2
from mmhand_b200.augment import images_to_bgr8
x = torch.tanh(torch.randn(32, 3, 256, 256, device="cuda"))
got = images_to_bgr8(x).cpu().numpy()
ref = np.stack([cv2.cvtColor(((x[i].permute(1, 2, 0).cpu().numpy() * 0.5 + 0.5) * 255.), cv2.COLOR_RGB2BGR) for i in range(32)])
print("bgr8 mismatches vs saturate(round):", int((got != np.clip(np.rint(ref), 0, 255).astype(np.uint8)).sum()))
PY
