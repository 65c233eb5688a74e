#!/bin/bash
# Diagnostic build of the one-launch a-trous stage with clock64 accumulators in thread 0 (SVGF_STAGE_TIMERS): cuda-path-tracer-denoising_b200/ab/libsvgf_timers.so
set -e
cd "$(dirname "$0")/../cuda-path-tracer-denoising_b200/csrc"
make -s
mkdir -p ../ab build_ab
FLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wall,-ffp-contract=off"
nvcc -ccbin /usr/bin/g++ $FLAGS -DSVGF_STAGE_TIMERS -c atrous.cu -o build_ab/atrous_timers.o
nvcc -ccbin /usr/bin/g++ -shared -gencode arch=compute_100a,code=sm_100a -o ../ab/libsvgf_timers.so build/api.o build/denoise.o build_ab/atrous_timers.o build/lbvh.o build/camera.o build/scene_ingest.o build/jpeg_decode.o build/pathtrace.o -Xlinker --no-undefined -lcudart
ls -la ../ab/libsvgf_timers.so
