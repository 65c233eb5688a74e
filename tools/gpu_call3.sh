#!/bin/bash
mkdir -p gpurun_out
timeout 200 ncu --set full --import-source on --clock-control none -k regex:rt_kernel --launch-skip 8 -c 1 -o gpurun_out/ncu_rt_c2 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_rt_c2.log 2>&1
ncu -i gpurun_out/ncu_rt_c2.ncu-rep --page raw --csv > gpurun_out/ncu_rt_c2_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_rt_c2.ncu-rep --page source --csv --print-source sass > gpurun_out/ncu_rt_c2_sass.csv 2>/dev/null
ncu -i gpurun_out/ncu_rt_c2.ncu-rep --page source --csv > gpurun_out/ncu_rt_c2_src.csv 2>/dev/null
ls -la gpurun_out/ | tail -8
