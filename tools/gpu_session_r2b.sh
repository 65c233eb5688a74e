#!/bin/bash
# Round-2 session B: full GPU suite (incl. the sliding a-trous kernel), A/B of the a-trous variants and the rt switches, C2 bench,
# one ncu capture of the sliding kernel, the uninitialised-read diagnostic.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout -s INT 1200 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/pytest_gpu.log | head -40
timeout 200 python tools/diag_denoise_entry.py > gpurun_out/diag_denoise_entry.log 2>&1; cat gpurun_out/diag_denoise_entry.log | tail -8
timeout 300 python tools/ab_atrous.py --workload c2 --frames 30 --shapes "" --extra "SVGF_ATROUS_VARIANT=5;SVGF_ATROUS_VARIANT=5+SVGF_ATROUS_BANDS=8;SVGF_ATROUS_VARIANT=5+SVGF_ATROUS_BANDS=16" > gpurun_out/ab_slide_c2.jsonl 2> gpurun_out/ab_slide_c2.err; cut -c1-300 gpurun_out/ab_slide_c2.jsonl
timeout 300 python tools/ab_atrous.py --workload c4 --frames 20 --shapes "" --strip 945,1215 --extra "SVGF_ATROUS_VARIANT=5" > gpurun_out/ab_slide_c4strip.jsonl 2> gpurun_out/ab_slide_c4strip.err; cut -c1-300 gpurun_out/ab_slide_c4strip.jsonl
for w in c2 c3 c5; do timeout 200 python tools/ab_atrous.py --workload $w --frames 20 --shapes "" --extra "SVGF_RT_ANYHIT=0" > gpurun_out/ab_rt_$w.jsonl 2> gpurun_out/ab_rt_$w.err; cut -c1-330 gpurun_out/ab_rt_$w.jsonl; done
timeout 300 python bench.py --steps 60 --warmup 10 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 1200 gpurun_out/bench_c2.json
SVGF_ATROUS_VARIANT=5 timeout 300 ncu --set full --import-source on --clock-control none -k regex:"atrous_slide" --launch-skip 15 -c 2 \
   -o gpurun_out/ncu_slide -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_slide.log 2>&1
ncu -i gpurun_out/ncu_slide.ncu-rep --page raw --csv > gpurun_out/ncu_slide_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_slide.ncu-rep --page source --csv > gpurun_out/ncu_slide_sass.csv 2>/dev/null
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
