#!/bin/bash
K="test_against_live_reference_gpu or test_every_switch or denoise_entry or goldens"
echo "== current (pinned arithmetic, default bounds)"; timeout 600 python -m pytest tests -m gpu -q -k "$K" 2>&1 | tail -2
for v in t6 t8; do echo "== $v"; SVGF_LIB_PATH=$PWD/cuda-path-tracer-denoising_b200/ab/libsvgf_$v.so timeout 600 python -m pytest tests -m gpu -q -k "$K" 2>&1 | tail -3; done
