#!/bin/bash
mkdir -p gpurun_out
for v in 0 1 2; do for w in c2 c3; do
  SVGF_RT_VARIANT=$v timeout 100 python tools/ab_atrous.py --workload $w --frames 20 --shapes "" 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('rt_variant $v', d['workload'][:10], 'rt_us', d.get('rt_us'), 'frame', d.get('frame_us'), d.get('error'))"
done; done
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "ingested" 2>&1 | tail -2
./tools/pipe_probe > gpurun_out/pipe_probe.txt 2>&1; tail -3 gpurun_out/pipe_probe.txt
