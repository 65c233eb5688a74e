#!/bin/bash
# Round-2 session A: the full GPU suite with the new parity cases (measured fractions -> gpurun_out/parity_report.json), a
# 2-rank --verify of the multi-GPU path on whatever GPUs the box has (1 GPU: both ranks share it), C2 bench for regressions.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout -s INT 900 python -m pytest tests -m gpu -q --durations=12 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 60 --warmup 10 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 1500 gpurun_out/bench_c2.json
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
