#!/bin/bash
# Round-2 session G: a-trous without the NaN guard (A/B + parity), BVH rebuild timing.
mkdir -p gpurun_out
timeout -s INT 600 python -m pytest tests -m gpu -q -x -k "atrous or parity" > gpurun_out/pytest_gpu_g.log 2>&1; tail -3 gpurun_out/pytest_gpu_g.log
for w in c2 c3; do
  timeout 200 python tools/ab_atrous.py --workload $w --frames 20 --shapes "" --extra "SVGF_ATROUS_NAN_GUARD=1" > gpurun_out/ab_nan_$w.jsonl 2> gpurun_out/ab_nan_$w.err; cut -c1-330 gpurun_out/ab_nan_$w.jsonl
done
timeout 200 python tools/time_bvh.py > gpurun_out/time_bvh.jsonl 2> gpurun_out/time_bvh.err; cat gpurun_out/time_bvh.jsonl; tail -3 gpurun_out/time_bvh.err
