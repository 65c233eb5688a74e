#!/bin/bash
# Round-1 GPU session A: parity of the new code, tile-shape A/B, ncu captures, then the rest of the GPU suite.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout -s INT 200 python -m pytest tests/test_gpu_atrous.py tests/test_gpu_async.py -m gpu -q --durations=8 > gpurun_out/pytest_new.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_new.log
timeout 200 python tools/ab_atrous.py --workload c2 --frames 30 --promo 2 --extra "SVGF_ATROUS_VARIANT=3" > gpurun_out/ab_c2.jsonl 2> gpurun_out/ab_c2.err
timeout 120 python tools/ab_atrous.py --workload c4 --frames 15 --shapes 0,2,3,5 > gpurun_out/ab_c4.jsonl 2> gpurun_out/ab_c4.err
timeout 100 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c2_default.json 2> gpurun_out/bench_c2_default.err
for s in 0 2; do
  SVGF_ATROUS_SHAPE=$s timeout 150 ncu --set full --import-source on --clock-control none -k regex:atrous_tiled --launch-skip 20 -c 5 \
     -o gpurun_out/ncu_atrous_shape$s -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_shape$s.log 2>&1
  ncu -i gpurun_out/ncu_atrous_shape$s.ncu-rep --page raw --csv > gpurun_out/ncu_atrous_shape${s}_raw.csv 2>/dev/null
  ncu -i gpurun_out/ncu_atrous_shape$s.ncu-rep --page source --csv --print-source sass > gpurun_out/ncu_atrous_shape${s}_sass.csv 2>/dev/null
done
du -sm gpurun_out/* > gpurun_out/sizes.txt
# keep the merged directory under the 64 MiB limit: the CSV exports carry what is read offline
if [ $(du -sm gpurun_out | cut -f1) -gt 55 ]; then rm -f gpurun_out/ncu_atrous_shape2.ncu-rep; fi
if [ $(du -sm gpurun_out | cut -f1) -gt 55 ]; then rm -f gpurun_out/ncu_atrous_shape0.ncu-rep; fi
timeout -s INT 300 python -m pytest tests -m gpu -x -q --durations=12 --deselect tests/test_gpu_atrous.py --deselect tests/test_gpu_async.py > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_new.log gpurun_out/pytest_gpu.log
cat gpurun_out/ab_c2.jsonl
