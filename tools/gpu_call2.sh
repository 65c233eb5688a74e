#!/bin/bash
mkdir -p gpurun_out
timeout 250 python tools/ab_atrous.py --workload c2 --frames 30 --shapes "" --extra "SVGF_ATROUS_SHAPES=2,2,2,2,6;SVGF_ATROUS_PROBE=1+SVGF_ATROUS_SHAPE=0;SVGF_ATROUS_PROBE=2+SVGF_ATROUS_SHAPE=0;SVGF_ATROUS_PROBE=1+SVGF_ATROUS_SHAPE=2;SVGF_ATROUS_PROBE=2+SVGF_ATROUS_SHAPE=2;SVGF_ATROUS_PROBE=1+SVGF_ATROUS_SHAPE=6;SVGF_ATROUS_PROBE=2+SVGF_ATROUS_SHAPE=6;SVGF_ATROUS_PROBE=2+SVGF_ATROUS_SHAPE=2+SVGF_ATROUS_VARIANT=3" > gpurun_out/ab2_c2.jsonl 2> gpurun_out/ab2_c2.err
cat gpurun_out/ab2_c2.jsonl; tail -3 gpurun_out/ab2_c2.err
