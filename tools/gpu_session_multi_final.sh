#!/bin/bash
# Final multi-GPU session of round 2 (gpurun --gpus N): bit-equality on the real transport, then bench lines (C4 strong scaling,
# C5 and C3 sharded) with per-rank stage times.
N=${1:-8}
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) "$@"; }
for w in c4 c5; do
  run bench.py --gpus $N --verify --workload $w --steps 6 > gpurun_out/verify_${N}gpu_$w.json 2> gpurun_out/verify_${N}gpu_$w.err
  echo "verify $w: $(tail -1 gpurun_out/verify_${N}gpu_$w.json | cut -c1-120)"; grep -E "^rank|Error|error" gpurun_out/verify_${N}gpu_$w.err | head -3
done
for w in c4 c5 c3 $EXTRA_WORKLOADS; do
  run bench.py --gpus $N --workload $w --steps 60 --warmup 10 > gpurun_out/bench_${N}gpu_$w.json 2> gpurun_out/bench_${N}gpu_$w.err
  python - gpurun_out/bench_${N}gpu_$w.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "fps %.1f"%d["fps"], "e2e %.1f"%d["e2e"]["fps"], "blocking %.1f"%d["e2e"]["blocking"]["fps"], "speedup", round(d.get("speedup_vs_1gpu_same_run",0),3), d["config"]["parallelism"][:90])
    for r in d.get("stages_ms_per_rank", []): print("   ", r)
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
