#!/bin/bash
# Round-2 session H: the a-trous stage as one launch: bit-equality with the per-level launches, A/B timing.
mkdir -p gpurun_out
timeout -s INT 600 python -m pytest tests -m gpu -q -x -k "stage_in_one_launch" > gpurun_out/pytest_gpu_h.log 2>&1; tail -15 gpurun_out/pytest_gpu_h.log | cut -c1-300
for w in c2 c4; do
  timeout 200 python tools/ab_atrous.py --workload $w --frames 20 --shapes "" --extra "SVGF_ATROUS_FUSED=0" > gpurun_out/ab_fused_$w.jsonl 2> gpurun_out/ab_fused_$w.err; cut -c1-330 gpurun_out/ab_fused_$w.jsonl; tail -2 gpurun_out/ab_fused_$w.err
done
timeout 200 python tools/ab_atrous.py --workload c4 --frames 20 --shapes "" --strip 945,1215 --extra "SVGF_ATROUS_FUSED=0" > gpurun_out/ab_fused_c4strip.jsonl 2> gpurun_out/ab_fused_c4strip.err; cut -c1-330 gpurun_out/ab_fused_c4strip.jsonl
