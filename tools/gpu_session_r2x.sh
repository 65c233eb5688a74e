#!/bin/bash
# 2-GPU check of the final code: bit-equality of sharded and unsharded frames on the real transport (C4, C5), then the C4 line
N=2
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) "$@"; }
for w in c4 c5; do
  run bench.py --gpus $N --verify --workload $w --steps 6 > gpurun_out/verify_${N}gpu_$w.json 2> gpurun_out/verify_${N}gpu_$w.err
  echo "verify $w: $(tail -1 gpurun_out/verify_${N}gpu_$w.json | cut -c1-120)"; grep -E "^rank|Error|error" gpurun_out/verify_${N}gpu_$w.err | head -3
done
run bench.py --gpus $N --steps 60 --warmup 10 > gpurun_out/bench_${N}gpu_c4.json 2> gpurun_out/bench_${N}gpu_c4.err
tail -1 gpurun_out/bench_${N}gpu_c4.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fps %.1f e2e %.1f blocking %.1f speedup %.3f'%(d['fps'],d['e2e']['fps'],d['e2e']['blocking']['fps'],d.get('speedup_vs_1gpu_same_run',0)))"
