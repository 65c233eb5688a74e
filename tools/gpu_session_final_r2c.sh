#!/bin/bash
# Closing 1-GPU session of round 2 on the final code (packed slab tests in the path tracer): GPU suite, smoke, then everything
# tools/gpu_session_final_r2b.sh records (bench lines for every config, launch list, full ncu captures).
mkdir -p gpurun_out
timeout -s INT 600 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
bash tools/gpu_session_final_r2b.sh
