#!/bin/bash
# Round-2 session F: kl folded into the tile kernel: parity (a-trous, multirank, parity suites) and A/B against the pre-pass.
mkdir -p gpurun_out
timeout -s INT 900 python -m pytest tests -m gpu -q -x -k "atrous or multirank or parity or sliding or async" > gpurun_out/pytest_gpu_f.log 2>&1; tail -4 gpurun_out/pytest_gpu_f.log
for w in c2 c4; do
  timeout 200 python tools/ab_atrous.py --workload $w --frames 20 --shapes "" --extra "SVGF_ATROUS_KL=prepass" > gpurun_out/ab_kl_$w.jsonl 2> gpurun_out/ab_kl_$w.err; cut -c1-330 gpurun_out/ab_kl_$w.jsonl
done
timeout 200 python tools/ab_atrous.py --workload c4 --frames 20 --shapes "" --strip 945,1215 --extra "SVGF_ATROUS_KL=prepass" > gpurun_out/ab_kl_c4strip.jsonl 2> gpurun_out/ab_kl_c4strip.err; cut -c1-330 gpurun_out/ab_kl_c4strip.jsonl
