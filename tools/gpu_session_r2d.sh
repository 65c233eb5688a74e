#!/bin/bash
# Round-2 session D: A/B of the light-query restructure (data-only exits), the denoise() entry diagnostics, the C2 bench.
mkdir -p gpurun_out
for w in c2 c3 c5; do
  timeout 200 python tools/ab_atrous.py --workload $w --frames 20 --shapes "" --extra "SVGF_RT_ANYHIT=0" > gpurun_out/ab_rt_$w.jsonl 2> gpurun_out/ab_rt_$w.err; cut -c1-130 gpurun_out/ab_rt_$w.jsonl
  for v in nolq nobvhany; do
    SVGF_LIB_PATH=$PWD/cuda-path-tracer-denoising_b200/ab/libsvgf_$v.so timeout 200 python tools/ab_atrous.py --workload $w --frames 20 --shapes "" > gpurun_out/ab_rt_${w}_$v.jsonl 2> gpurun_out/ab_rt_${w}_$v.err
    echo "$v: $(cut -c1-130 gpurun_out/ab_rt_${w}_$v.jsonl)"
  done
done
timeout 600 python -m pytest tests -m gpu -q -x -k "parity or denoise_entry or wavefront or async" > gpurun_out/pytest_gpu_d.log 2>&1; tail -5 gpurun_out/pytest_gpu_d.log
timeout 300 python -m pytest tests/test_gpu_denoise_entry.py -m gpu -q -k "room" > gpurun_out/pytest_entry_alone.log 2>&1; tail -3 gpurun_out/pytest_entry_alone.log
timeout 200 python tools/diag_denoise_entry.py > gpurun_out/diag_denoise_entry.log 2>&1; tail -12 gpurun_out/diag_denoise_entry.log
timeout 300 python bench.py --steps 60 --warmup 10 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; cut -c1-400 gpurun_out/bench_c2.json
