#!/bin/bash
# TEA hash loop unrolled (4, 16) and packed mat-vec products (A/B builds of tools/build_rt_ab.sh): stage times on C2, C3, C5
mkdir -p gpurun_out
for w in c2 c3 c5; do
  echo "default $w: $(timeout 200 python tools/ab_atrous.py --workload $w --frames 20 --shapes "" 2>&1 | cut -c1-170)"
  for t in tea4 tea16 mv mvtea4; do
    echo "$t $w: $(SVGF_LIB_PATH=$PWD/cuda-path-tracer-denoising_b200/ab/libsvgf_$t.so timeout 200 python tools/ab_atrous.py --workload $w --frames 20 --shapes "" 2>&1 | cut -c1-170)"
  done
done 2>&1 | tee gpurun_out/ab_rt_tea_mv.txt
