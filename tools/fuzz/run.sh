#!/bin/bash
# Mutation fuzzing of the host-side readers (JPEG decoder, scene text reader, OBJ reader) under ASan + UBSan.
#   tools/fuzz/run.sh <textures_dir> [iterations] [seeds]
# <textures_dir> holds the JPEG files to mutate (the scenes' Textures directory); the scene/OBJ seeds are tests/golden/scenes_txt.
# A clean run prints "decoded/loaded N rejected M" per seed and exits 0; any sanitizer report is a bug.
set -e
HERE=$(cd "$(dirname "$0")" && pwd); ROOT=$(cd "$HERE/../.." && pwd)
TEX=${1:?textures dir}; IT=${2:-1000}; SEEDS=${3:-4}
OUT=${TMPDIR:-/tmp}/svgf_fuzz_build; mkdir -p "$OUT"
C="$ROOT/cuda-path-tracer-denoising_b200/csrc"
FL="-O1 -g -fsanitize=address,undefined -fno-sanitize-recover=undefined -std=c++17 -I$ROOT/include"
g++ $FL "$HERE/fuzz_jpeg.cpp" "$C/jpeg_decode.cpp" -o "$OUT/fuzz_jpeg"
g++ $FL "$HERE/fuzz_scene.cpp" "$C/scene_ingest.cpp" "$C/jpeg_decode.cpp" "$C/camera.cpp" -o "$OUT/fuzz_scene"
g++ $FL "$HERE/fuzz_obj.cpp" "$C/scene_ingest.cpp" "$C/jpeg_decode.cpp" "$C/camera.cpp" -o "$OUT/fuzz_obj"
G="$ROOT/tests/golden/scenes_txt"
for s in $(seq 1 "$SEEDS"); do
  "$OUT/fuzz_jpeg" "$IT" "$s" "$TEX"/*.jpg
  "$OUT/fuzz_scene" "$IT" "$s" "$G/Models" "$G/two_meshes.txt" "$G/two_lights.txt" "$G/missing_mesh.txt"
  "$OUT/fuzz_obj" "$IT" "$s" "$G/two_meshes.txt" "$G/Models" pyramid.obj quad.obj
done
