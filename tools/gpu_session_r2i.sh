#!/bin/bash
mkdir -p gpurun_out
timeout -s INT 600 python -m pytest tests -m gpu -q -x -k "stage_in_one_launch" 2>&1 | tail -3
SVGF_LIB_PATH=$PWD/cuda-path-tracer-denoising_b200/ab/libsvgf_timers.so timeout 200 python tools/stage_timers.py c2 2>&1 | tail -3
for w in c2 c4; do timeout 200 python tools/ab_atrous.py --workload $w --frames 20 --shapes "" --extra "SVGF_ATROUS_FUSED=0" 2>&1 | cut -c1-300; done
timeout 200 python tools/ab_atrous.py --workload c4 --frames 20 --shapes "" --strip 945,1215 --extra "SVGF_ATROUS_FUSED=0" 2>&1 | cut -c1-300
