#!/bin/bash
# compute-sanitizer over the new kernels at small sizes: memcheck and racecheck on the a-trous kernels (TMA +
# mbarrier paths, every tile shape, the sliding kernel), the multi-rank path (cross-context flags, dual stores) and the
# pipelined readback. Tails go to gpurun_out/ (copied to profiles/ by hand). The 8-contexts-on-one-GPU case is left out: under the
# tools' slowdown it runs into the library's bounded 2 s cross-rank wait (SVGF_ERR_COMM by design).
mkdir -p gpurun_out
K='test_every_tile_shape_matches_oracle or test_tiny_and_ragged_sizes or test_sharded_equals_unsharded_bitwise or test_uneven_partition or test_async_equals_blocking or test_sliding_variant_matches_oracle'
for tool in memcheck racecheck; do
  timeout 700 compute-sanitizer --tool $tool --target-processes all --print-limit 20 python -m pytest tests -m gpu -q -k "($K) and not w8" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool: exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_$tool.log | tail -5
done
