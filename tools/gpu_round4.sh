#!/bin/bash
# GPU-box visit: parity tests, benches of the three workloads, ncu launch list of one step, full captures of the
# conv / elementwise / rasteriser kernels, batch-64 footprint run (configs[4] shape on one GPU).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
timeout 300 python bench.py --workload infer > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err; echo "infer rc=$?"
timeout 300 python bench.py --workload raster > gpurun_out/bench_raster.json 2> gpurun_out/bench_raster.err; echo "raster rc=$?"
timeout 400 python bench.py --batch 64 --steps 5 --no-cpu-baseline > gpurun_out/bench_b64.json 2> gpurun_out/bench_b64.err; echo "b64 rc=$?"
python -c "
import json,torch
for f in ('bench_infer','bench_raster','bench_b64'):
    try:
        d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['value'], d['unit'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])
    except Exception as e: print(f, 'failed', e)
"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'conv2_kernel' -s 40 -c 8 -f -o gpurun_out/conv2_full python tools/profile_step.py > gpurun_out/ncu_conv2.log 2>&1; echo "ncu conv2 rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'pg_kernel|reduce_ch_fin_kernel' -s 200 -c 16 -f -o gpurun_out/ew_full python tools/profile_step.py > gpurun_out/ncu_ew.log 2>&1; echo "ncu ew rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'raster_kernel' -s 3 -c 2 -f -o gpurun_out/raster_full \
  python bench.py --workload raster --steps 3 --no-cpu-baseline > gpurun_out/ncu_raster.log 2>&1; echo "ncu raster rc=$?"
ls -la gpurun_out/*.ncu-rep
