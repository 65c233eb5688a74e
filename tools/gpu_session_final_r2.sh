#!/bin/bash
# End-of-round-2 GPU session (1 GPU): full GPU suite, compute-sanitizer (memcheck + racecheck) over the new kernels at small sizes,
# bench lines for every BASELINE.json config + 720p (ours + reference arm), the ncu launch list of the bench command, full ncu
# captures of the denoise kernels and of rt_kernel with source-level stall samples, BVH rebuild timing, smoke. Output: gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout -s INT 600 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
cp gpurun_out/parity_report.json gpurun_out/parity_report_full.json 2>/dev/null
timeout 300 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
for w in c1 c3 c4 c5 cornell720 room720; do
  timeout 200 python bench.py --workload $w --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
for w in c2 c1 c3 c5 cornell720; do
  timeout 200 python bench.py --impl reference --workload $w --steps 30 --warmup 5 > gpurun_out/bench_ref_$w.json 2> gpurun_out/bench_ref_$w.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"atrous_tiled|atrous_kl|temporal_kernel|pack_pbo" --launch-skip 48 -c 12 \
   -o gpurun_out/ncu_denoise -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_denoise.log 2>&1
ncu -i gpurun_out/ncu_denoise.ncu-rep --page raw --csv > gpurun_out/ncu_denoise_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_denoise.ncu-rep --page source --csv --kernel-name regex:atrous_tiled --launch-count 1 > gpurun_out/ncu_tiled_src.csv 2>/dev/null
ncu -i gpurun_out/ncu_denoise.ncu-rep --page source --csv --kernel-name regex:temporal --launch-count 1 > gpurun_out/ncu_temporal_src.csv 2>/dev/null
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"rt_kernel" --launch-skip 4 -c 1 \
   -o gpurun_out/ncu_rt -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_rt.log 2>&1
ncu -i gpurun_out/ncu_rt.ncu-rep --page raw --csv > gpurun_out/ncu_rt_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_rt.ncu-rep --page source --csv > gpurun_out/ncu_rt_src.csv 2>/dev/null
rm -f gpurun_out/ncu_slide.ncu-rep gpurun_out/ncu_stage.ncu-rep
timeout 200 python tools/time_bvh.py > gpurun_out/time_bvh.jsonl 2> gpurun_out/time_bvh.err
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
tail -1 gpurun_out/smoke.log
bash tools/gpu_session_sanitizer.sh
for f in gpurun_out/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d.get("impl","ours"), d["config"]["workload"][:14], "fps %.1f"%d.get("fps",0), "e2e", round(d.get("e2e",{}).get("fps",0),1), "blk", round(d.get("e2e",{}).get("blocking",{}).get("fps",0),1), "frac", d.get("roofline",{}).get("frac"))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
du -sm gpurun_out
