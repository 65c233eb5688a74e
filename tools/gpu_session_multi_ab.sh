#!/bin/bash
# A/B of the sharded frame's protocol switches on N GPUs: per-rank stage times of C4 for each setting.
N=${1:-2}
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) "$@"; }
i=0
for cfg in "" ${CFGS:-SVGF_RT_PUSH=0 SVGF_HALO=pull SVGF_CUDA_GRAPH=1}; do
  i=$((i+1))
  env $cfg bash -c "$(declare -f run); N=$N; run bench.py --gpus $N --workload c4 --steps 40 --warmup 5 --no-cpu-baseline" > gpurun_out/multi_ab_${N}gpu_$i.json 2> gpurun_out/multi_ab_${N}gpu_$i.err
  python - gpurun_out/multi_ab_${N}gpu_$i.json "$cfg" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("[%s]" % sys.argv[2], "fps %.1f"%d["fps"], "speedup", round(d.get("speedup_vs_1gpu_same_run",0),3), d["config"]["parallelism"][:60])
    for r in d.get("stages_ms_per_rank", []): print("   ", r)
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
