#!/bin/bash
# temporal_kernel at 6 / 8 blocks per SM (40 / 32 registers) against the default (48 registers, 5 blocks)
for w in c2 c5; do
  timeout 200 python tools/ab_atrous.py --workload $w --frames 20 --shapes "" 2>&1 | cut -c1-150
  for v in t6 t8; do echo "$v: $(SVGF_LIB_PATH=$PWD/cuda-path-tracer-denoising_b200/ab/libsvgf_$v.so timeout 200 python tools/ab_atrous.py --workload $w --frames 20 --shapes "" 2>&1 | cut -c1-150)"; done
done
