#!/bin/bash
mkdir -p gpurun_out
timeout -s INT 200 python -m pytest tests/test_gpu_atrous.py -m gpu -q -x > gpurun_out/pytest_atrous.log 2>&1; tail -3 gpurun_out/pytest_atrous.log
timeout 200 python tools/ab_atrous.py --workload c2 --frames 30 --shapes 0,2,9 --extra "SVGF_ATROUS_PROBE=1+SVGF_ATROUS_SHAPE=2" > gpurun_out/ab5_c2.jsonl 2> gpurun_out/ab5.err
python - <<'PY'
import json
for l in open("gpurun_out/ab5_c2.jsonl"):
    d = json.loads(l); print(d["env"], d.get("level_us"), d.get("atrous_us"), d.get("frame_us"), d.get("error"))
PY
tail -3 gpurun_out/ab5.err
