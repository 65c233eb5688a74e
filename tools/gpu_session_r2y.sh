#!/bin/bash
# Temporal kernel without owner lookups / integer divisions: stage times (A/B through SVGF_TEMPORAL_SINGLE=0), then the GPU suite
mkdir -p gpurun_out
for w in c2 c5 c4; do
  echo "new $w: $(timeout 120 python tools/ab_atrous.py --workload $w --frames 30 --shapes "" 2>&1 | cut -c1-130)"
  echo "general kernel $w: $(SVGF_TEMPORAL_SINGLE=0 timeout 120 python tools/ab_atrous.py --workload $w --frames 30 --shapes "" 2>&1 | cut -c1-130)"
done 2>&1 | tee gpurun_out/ab_temporal_single.txt
timeout -s INT 500 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest_temporal_single.log
