#!/bin/bash
# Last 1-GPU session of round 2, on the final code: bench lines (ours) for every BASELINE.json config + 720p, the ncu launch list
# and the full denoise capture (the reference arm, sanitizer and rt captures of tools/gpu_session_final_r2.sh stand).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 300 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
for w in c1 c3 c4 c5 cornell720 room720; do
  timeout 200 python bench.py --workload $w --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
timeout 200 python bench.py --impl reference --steps 30 --warmup 5 > gpurun_out/bench_ref_c2.json 2> gpurun_out/bench_ref_c2.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"atrous_tiled|atrous_kl|temporal_kernel|pack_pbo" --launch-skip 48 -c 12 \
   -o gpurun_out/ncu_denoise -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_denoise.log 2>&1
ncu -i gpurun_out/ncu_denoise.ncu-rep --page raw --csv > gpurun_out/ncu_denoise_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_denoise.ncu-rep --page source --csv --kernel-name regex:temporal --launch-count 1 > gpurun_out/ncu_temporal_src.csv 2>/dev/null
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"rt_kernel" --launch-skip 4 -c 1 \
   -o gpurun_out/ncu_rt -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_rt.log 2>&1
ncu -i gpurun_out/ncu_rt.ncu-rep --page raw --csv > gpurun_out/ncu_rt_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_rt.ncu-rep --page source --csv > gpurun_out/ncu_rt_src.csv 2>/dev/null
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
for f in gpurun_out/bench_c*.json gpurun_out/bench_room720.json gpurun_out/bench_ref_c2.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d.get("impl","ours"), d["config"]["workload"][:14], "fps %.1f"%d.get("fps",0), "e2e", round(d.get("e2e",{}).get("fps",0),1), "blk", round(d.get("e2e",{}).get("blocking",{}).get("fps",0),1), "frac", d.get("roofline",{}).get("frac"), "traffic", d.get("roofline",{}).get("traffic"), d.get("e2e",{}).get("host_ms_per_step"))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
