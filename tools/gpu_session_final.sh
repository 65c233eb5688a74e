#!/bin/bash
# End-of-round GPU session (EXTRA_WORKLOADS="c1 c3 c5 c4" EXTRA_REF_WORKLOADS="c1 c3" for every BASELINE.json config): full GPU test suite, bench lines for every BASELINE.json config (ours + reference arm),
# the ncu launch list of the bench command and one full capture of the a-trous / temporal kernels. Output: gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout -s INT 400 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 200 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
for w in $EXTRA_WORKLOADS; do
  timeout 200 python bench.py --workload $w --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
for w in c2 $EXTRA_REF_WORKLOADS; do
  timeout 200 python bench.py --impl reference --workload $w --steps 30 --warmup 5 > gpurun_out/bench_ref_$w.json 2> gpurun_out/bench_ref_$w.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"atrous_tiled|atrous_kl|temporal_kernel|pack_pbo" --launch-skip 48 -c 12 \
   -o gpurun_out/ncu_denoise -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_denoise.log 2>&1
ncu -i gpurun_out/ncu_denoise.ncu-rep --page raw --csv > gpurun_out/ncu_denoise_raw.csv 2>/dev/null
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -1 gpurun_out/smoke.log; du -sm gpurun_out
for f in gpurun_out/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d.get("impl","ours"), d["config"]["workload"][:12], "fps %.1f"%d.get("fps",0), "e2e", d.get("e2e",{}).get("fps"), "frac", d.get("roofline",{}).get("frac"))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
