#!/bin/bash
# north_star's config matrix on N GPUs (gpurun --gpus N): cornell and room at 720p / 1080p / 4K, sharded bench lines only
# (bit-equality on the real transport: tools/gpu_session_multi_final.sh).
N=${1:-8}
mkdir -p gpurun_out
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) "$@"; }
for w in ${WORKLOADS:-room4k c2 cornell720 room720}; do
  run bench.py --gpus $N --workload $w --steps 60 --warmup 10 > gpurun_out/bench_${N}gpu_$w.json 2> gpurun_out/bench_${N}gpu_$w.err
  python - gpurun_out/bench_${N}gpu_$w.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "fps %.1f"%d["fps"], "e2e %.1f"%d["e2e"]["fps"], "blocking %.1f"%d["e2e"]["blocking"]["fps"], "anchor %.1f"%d["anchor_1gpu"]["fps"], "speedup", round(d.get("speedup_vs_1gpu_same_run",0),3))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
