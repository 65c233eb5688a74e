#!/bin/bash
# compute-sanitizer on the round's last kernel changes: the usual selection (tools/gpu_session_sanitizer.sh: temporal kernel in both
# instantiations, packed slab tests), then racecheck + memcheck of the compacting path tracer (shared-memory exchange between barriers)
bash tools/gpu_session_sanitizer.sh
for tool in memcheck racecheck; do
  SVGF_RT_COMPACT=2 timeout 400 compute-sanitizer --tool $tool --target-processes all --print-limit 20 python -m pytest tests -m gpu -q -k "test_async_equals_blocking or test_sharded_equals_unsharded_bitwise and not w8" > gpurun_out/sanitizer_compact_$tool.log 2>&1
  echo "== compact $tool: exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_compact_$tool.log | tail -3
done
