#!/bin/bash
mkdir -p gpurun_out
timeout -s INT 120 python -m pytest tests/test_gpu_atrous.py -m gpu -q -k "tile_shape" > gpurun_out/pytest_shapes.log 2>&1; tail -2 gpurun_out/pytest_shapes.log
timeout 200 python tools/ab_atrous.py --workload c2 --frames 30 --shapes 2,6,8,9,10 > gpurun_out/ab4_c2.jsonl 2> gpurun_out/ab4.err
for strip in 945,1215 0,270 810,1350 0,1080; do
  timeout 200 python tools/ab_atrous.py --workload c4 --frames 15 --shapes 2,6,8,9,10 --strip $strip >> gpurun_out/ab4_c4_strips.jsonl 2>> gpurun_out/ab4.err
done
python - <<'PY'
import json
for f in ("gpurun_out/ab4_c2.jsonl", "gpurun_out/ab4_c4_strips.jsonl"):
    for l in open(f):
        d = json.loads(l)
        print(d.get("strip"), d["env"], d.get("level_us"), d.get("atrous_us"), d.get("rt_us"), d.get("error"))
PY
tail -3 gpurun_out/ab4.err
