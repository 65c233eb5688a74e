N=4
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) "$@"; }
run bench.py --gpus $N --verify --workload c4 --steps 6 > gpurun_out/verify_${N}gpu_c4.json 2> gpurun_out/verify_${N}gpu_c4.err; echo "verify: $(tail -1 gpurun_out/verify_${N}gpu_c4.json | cut -c1-100)"
run bench.py --gpus $N --workload c4 --steps 60 --warmup 10 > gpurun_out/bench_${N}gpu_c4.json 2> gpurun_out/bench_${N}gpu_c4.err
python - gpurun_out/bench_${N}gpu_c4.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("fps %.1f"%d["fps"], "e2e %.1f"%d["e2e"]["fps"], "blocking %.1f"%d["e2e"]["blocking"]["fps"], "speedup", round(d.get("speedup_vs_1gpu_same_run",0),3))
for r in d.get("stages_ms_per_rank", []): print("   ", r)
PY
