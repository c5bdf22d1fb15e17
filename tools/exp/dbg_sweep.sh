#!/bin/bash
# Timing-only ablations of conv2_kernel (MMH_C2_DEBUG: 1 skip A loads, 2 skip B loads, 4 skip MMAs, 8 skip stores)
for c in ${CASES:-perf perf512 perf_stem perf_d1}; do
  for d in ${DBGS:-0 8 4 3 12 7}; do
    echo -n "case=$c dbg=$d " ; MMH_C2_DEBUG=$d python tools/bringup.py --one $c 2>/dev/null | grep RESULT | python -c "
import sys, json
r = json.loads(sys.stdin.read()[7:])
print(' '.join('%s=%.3f' % (k, v) for k, v in r.items() if k.endswith('_ms') and 'wgrad' not in k))"
  done
done
