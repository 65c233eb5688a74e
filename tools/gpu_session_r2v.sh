#!/bin/bash
# Compacting path tracer (SVGF_RT_COMPACT) against the default kernel: times, then the GPU suite under both
mkdir -p gpurun_out
for w in c2 c3 c5; do
  echo "default $w: $(timeout 120 python tools/ab_atrous.py --workload $w --frames 20 --shapes "" 2>&1 | cut -c1-170)"
  [ $w = c2 ] && echo "compact1 $w: $(SVGF_RT_COMPACT=1 timeout 120 python tools/ab_atrous.py --workload $w --frames 20 --shapes "" 2>&1 | cut -c1-170)"
  echo "compact2 $w: $(SVGF_RT_COMPACT=2 timeout 120 python tools/ab_atrous.py --workload $w --frames 20 --shapes "" 2>&1 | cut -c1-170)"
done 2>&1 | tee gpurun_out/ab_rt_compact.txt
echo "== suite, SVGF_RT_COMPACT=2"
SVGF_RT_COMPACT=2 timeout -s INT 500 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest_compact2.log
echo "== suite, default"
timeout -s INT 500 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest_default.log
