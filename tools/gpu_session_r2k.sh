#!/bin/bash
# Round-2 session K: cross-frame overlap: parity and bench A/B.
mkdir -p gpurun_out
timeout -s INT 600 python -m pytest tests -m gpu -q -x -k "overlap or async or abi or shim" 2>&1 | tail -5
for w in c2 c3 c5 c4; do
  for ov in 1 0; do
    echo "$w overlap=$ov: $(SVGF_FRAME_OVERLAP=$ov timeout 200 python bench.py --workload $w --steps 60 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fps %.1f e2e %.1f blocking %.1f ms %.4f' % (d['fps'], d['e2e']['fps'], d['e2e']['blocking']['fps'], d['ms_per_step']))")"
  done
done
