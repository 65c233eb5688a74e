#!/bin/bash
# Round-2 session J (1 GPU): multi-rank-on-one-GPU parity with the copy-kernel halo push, rt warp shapes.
mkdir -p gpurun_out
timeout -s INT 600 python -m pytest tests -m gpu -q -x -k "multirank" 2>&1 | tail -3
for w in c2 c3; do
  timeout 200 python tools/ab_atrous.py --workload $w --frames 20 --shapes "" 2>&1 | cut -c1-140
  for v in w4x8 w16x2 w32x1; do
    echo "$v: $(SVGF_LIB_PATH=$PWD/cuda-path-tracer-denoising_b200/ab/libsvgf_$v.so timeout 200 python tools/ab_atrous.py --workload $w --frames 20 --shapes "" 2>&1 | cut -c1-140)"
  done
done
