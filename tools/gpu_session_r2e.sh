#!/bin/bash
# Round-2 session E: light-first A/B, the denoise() entry diagnostic, the CUDA-graph test.
mkdir -p gpurun_out
for w in c2 c3 c5; do
  timeout 200 python tools/ab_atrous.py --workload $w --frames 20 --shapes "" > gpurun_out/ab_rt_$w.jsonl 2> gpurun_out/ab_rt_$w.err; cut -c1-130 gpurun_out/ab_rt_$w.jsonl
  for v in lightfirst; do
    SVGF_LIB_PATH=$PWD/cuda-path-tracer-denoising_b200/ab/libsvgf_$v.so timeout 200 python tools/ab_atrous.py --workload $w --frames 20 --shapes "" > gpurun_out/ab_rt_${w}_$v.jsonl 2> gpurun_out/ab_rt_${w}_$v.err
    echo "$v: $(cut -c1-130 gpurun_out/ab_rt_${w}_$v.jsonl)"
  done
done
timeout 300 python tools/diag_entry2.py > gpurun_out/diag_entry2.log 2>&1; tail -16 gpurun_out/diag_entry2.log
timeout 300 python -m pytest tests/test_gpu_async.py -m gpu -q > gpurun_out/pytest_async.log 2>&1; tail -5 gpurun_out/pytest_async.log
