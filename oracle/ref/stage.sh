#!/usr/bin/env bash
# oracle/ref/stage.sh -- TEST INFRASTRUCTURE (oracle/), not product code.
#
# Stages a *temporary* patched copy of the reference's hot-path sources so they can be compiled
# from where they lie under /root/reference (read-only) by nvcc 12.9 / g++ 13. Nothing staged here
# is ever written into the repository: the caller passes a mktemp directory and removes it after
# the build; only the resulting .so files land in oracle/_ref/.
#
#   stage.sh <reference_root> <stage_dir> <variant>
#
# variant:
#   gpu         reference as-is (+ portability patch)                   -> nvcc
#   gpu_jacobi  + ATrousFilter writes variance to a second buffer       -> nvcc
#   cpu         + <<<>>> launches rewritten to EMU_LAUNCH (cuda_emu/)   -> g++
#   cpu_jacobi  both
#
# Portability patch (MSVC-isms that nvcc/g++ reject):
#   src/boundingbox.h:6      `extern struct Ray {`            -> `struct Ray {`
#   src/boundingbox.h:36,52  `BoundingBox& operator||(...)`   -> returns by value, const
# Jacobi patch (deterministic oracle variant, SURVEY.md section 8(c) "protocol (ii)"):
#   src/denoise.cu:161       variance[p] = ...   ->   variance_out[p] = ...
#   with variance_out pre-filled from variance before each level and copied back after it,
#   so every read in a level sees the previous level's variance (no in-place race).
set -euo pipefail
REF="$1"; STAGE="$2"; VARIANT="$3"

mkdir -p "$STAGE/src"
cp "$REF"/src/*.h "$REF"/src/*.hpp "$REF"/src/*.cpp "$REF"/src/*.cu "$STAGE/src/"
cp -r "$REF"/src/tinyobjloader "$STAGE/src/"
chmod -R u+w "$STAGE"

sed -i \
  -e 's/^extern struct Ray {/struct Ray {/' \
  -e 's/BoundingBox& operator || (BoundingBox& b2)/BoundingBox operator || (const BoundingBox\& b2) const/' \
  -e 's/BoundingBox& operator || (const glm::vec3& p)/BoundingBox operator || (const glm::vec3\& p) const/' \
  "$STAGE/src/boundingbox.h"
grep -q '^struct Ray {' "$STAGE/src/boundingbox.h"
[ "$(grep -c 'BoundingBox operator || (const' "$STAGE/src/boundingbox.h")" = 2 ]

case "$VARIANT" in
  *jacobi)
    perl -0pi -e '
      s/(__global__ void ATrousFilter\(glm::vec3 \* colorin, glm::vec3 \* colorout, float \* variance,)/$1 float * variance_out,/ or die "sig";
      s/variance\[p\] = variance_sum \/ weights_squared_sum;/variance_out[p] = variance_sum \/ weights_squared_sum;/ or die "write";
      s/(static float \* dev_variance = NULL;)/$1\nstatic float * dev_variance_jacobi = NULL;/ or die "decl";
      s/(cudaMalloc\(&dev_variance, pixelcount \* sizeof\(float\)\);)/$1\n    cudaMalloc(&dev_variance_jacobi, pixelcount * sizeof(float));/ or die "alloc";
      s/(cudaFree\(dev_variance\);)/$1\n    cudaFree(dev_variance_jacobi);/ or die "free";
      s/(ATrousFilter<<<blocksPerGrid2d, blockSize2d>>>\(src, dst, dev_variance,)/cudaMemcpy(dev_variance_jacobi, dev_variance, pixelcount * sizeof(float), cudaMemcpyDeviceToDevice);\n                $1 dev_variance_jacobi,/ or die "launch";
      s/(if \(level == ui_history_level\) cudaMemcpy\(dev_color_history, dst, pixelcount \* sizeof\(glm::vec3\), cudaMemcpyDeviceToDevice\);)/$1\n                cudaMemcpy(dev_variance, dev_variance_jacobi, pixelcount * sizeof(float), cudaMemcpyDeviceToDevice);/ or die "copyback";
    ' "$STAGE/src/denoise.cu"
    ;;
esac

case "$VARIANT" in
  cpu*)
    for f in denoise.cu pathtrace.cu; do
      perl -0pi -e 's/(\w+)\s*<<<\s*(.*?)\s*>>>\s*(\((?:[^()]++|(?3))*\))/EMU_LAUNCH($1, ($2), $3)/gs' "$STAGE/src/$f"
      if grep -q '<<<' "$STAGE/src/$f"; then echo "launch rewrite incomplete in $f" >&2; exit 1; fi
    done
    ;;
esac
