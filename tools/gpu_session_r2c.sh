#!/bin/bash
# Round-2 session C (run with gpurun --gpus N, N >= 2): parity of the REAL multi-GPU path (bench.py --verify: one process per GPU,
# CUDA IPC, dual stores over NVLink, cross-device flags, uneven strips; every rank's strip bit for bit against an unsharded
# render on the same GPU), then the sharded bench lines.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi_multi.txt 2>&1
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) "$@"; }
for w in c4 c5 c3; do
  run bench.py --gpus $N --verify --workload $w --steps 6 > gpurun_out/verify_${N}gpu_$w.json 2> gpurun_out/verify_${N}gpu_$w.err
  echo "verify $w: $(tail -1 gpurun_out/verify_${N}gpu_$w.json | cut -c1-200)"; grep -E "^rank|Error|error" gpurun_out/verify_${N}gpu_$w.err | head -5
done
for w in c4 ${EXTRA_WORKLOADS}; do
  run bench.py --gpus $N --workload $w --steps 60 --warmup 10 > gpurun_out/bench_${N}gpu_$w.json 2> gpurun_out/bench_${N}gpu_$w.err
  python - gpurun_out/bench_${N}gpu_$w.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "fps %.1f"%d["fps"], "e2e %.1f"%d["e2e"]["fps"], "blocking %.1f"%d["e2e"]["blocking"]["fps"], "anchor", d.get("anchor_1gpu",{}).get("fps"), "speedup", d.get("speedup_vs_1gpu_same_run"), "stages", d["stages_ms"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
