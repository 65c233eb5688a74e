#!/bin/bash
K="test_against_live_reference_gpu and bunny-640"
echo "== current"; timeout 300 python -m pytest tests -m gpu -q -k "$K" 2>&1 | tail -2
echo "== PDL off"; SVGF_PDL=0 timeout 300 python -m pytest tests -m gpu -q -k "$K" 2>&1 | tail -2
for v in t1 t5 t6; do echo "== $v"; SVGF_LIB_PATH=$PWD/cuda-path-tracer-denoising_b200/ab/libsvgf_$v.so timeout 300 python -m pytest tests -m gpu -q -k "$K" 2>&1 | tail -2; done
