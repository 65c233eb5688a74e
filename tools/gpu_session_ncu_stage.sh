#!/bin/bash
# One full ncu capture of the one-launch a-trous stage kernel at C2 (source-level stall samples included).
mkdir -p gpurun_out
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"atrous_stage" --launch-skip 8 -c 1 \
   -o gpurun_out/ncu_stage -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_stage.log 2>&1
ncu -i gpurun_out/ncu_stage.ncu-rep --page raw --csv > gpurun_out/ncu_stage_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_stage.ncu-rep --page source --csv > gpurun_out/ncu_stage_src.csv 2>/dev/null
python tools/ncu_source_summary.py < gpurun_out/ncu_stage_src.csv | cut -c1-400
tail -3 gpurun_out/ncu_stage.log
