#!/bin/bash
# Round-2 session L: rt_kernel occupancy targets 5 / 6 / 7 blocks per SM against the default 8.
for w in c2 c3; do
  timeout 200 python tools/ab_atrous.py --workload $w --frames 20 --shapes "" 2>&1 | cut -c1-140
  for v in minb5 minb6 minb7; do
    echo "$v: $(SVGF_LIB_PATH=$PWD/cuda-path-tracer-denoising_b200/ab/libsvgf_$v.so timeout 200 python tools/ab_atrous.py --workload $w --frames 20 --shapes "" 2>&1 | cut -c1-140)"
  done
done
