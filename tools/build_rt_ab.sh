#!/bin/bash
# A/B builds of the path tracer (same objects, pathtrace.cu recompiled with one switch each): cuda-path-tracer-denoising_b200/ab/libsvgf_<tag>.so
set -e
cd "$(dirname "$0")/../cuda-path-tracer-denoising_b200/csrc"
make -s
mkdir -p ../ab build_ab
FLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wall,-ffp-contract=off"
for tag in ${RT_AB_TAGS:-lq:-DSVGF_RT_LIGHT_QUERY lightfirst:-DSVGF_RT_LIGHT_FIRST}; do
  name=${tag%%:*}; defs=${tag#*:}; defs=${defs//@/ }
  nvcc -ccbin /usr/bin/g++ $FLAGS $defs -c pathtrace.cu -o build_ab/pathtrace_$name.o &
done
wait
for tag in ${RT_AB_TAGS:-lq:-DSVGF_RT_LIGHT_QUERY lightfirst:-DSVGF_RT_LIGHT_FIRST}; do name=${tag%%:*}
  nvcc -ccbin /usr/bin/g++ -shared -gencode arch=compute_100a,code=sm_100a -o ../ab/libsvgf_$name.so build/api.o build/denoise.o build/atrous.o build/lbvh.o build/camera.o build/scene_ingest.o build/jpeg_decode.o build_ab/pathtrace_$name.o -Xlinker --no-undefined -lcudart
done
ls -la ../ab
