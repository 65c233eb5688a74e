#!/bin/bash
# Paired-children BVH traversal: parity (everything that path-traces a mesh) and timing against the round-1 loop
timeout -s INT 700 python -m pytest tests -m gpu -q -x -k "parity or wavefront or multirank" 2>&1 | tail -4
for w in c3 c5 c2; do
  timeout 200 python tools/ab_atrous.py --workload $w --frames 20 --shapes "" 2>&1 | cut -c1-150
  echo "round-1 loop: $(SVGF_LIB_PATH=$PWD/cuda-path-tracer-denoising_b200/ab/libsvgf_bvhold.so timeout 200 python tools/ab_atrous.py --workload $w --frames 20 --shapes "" 2>&1 | cut -c1-150)"
done
